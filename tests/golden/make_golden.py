"""Generates tests/golden/golden_cv.npz from the importable third-party library the reference calls (cv2), with the
reference's exact arguments.  Run in the build container:  python tests/golden/make_golden.py
The fixture pins the oracle (and, on the GPU box, the CUDA path) independently of the cv2 build present at test time."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from util import texture, moved_pair, stereo_pair  # noqa: E402

out = {"cv2_version": np.array(cv2.__version__)}
img = texture(96, 160, 11)
out["img"] = img
out["half"] = cv2.resize(img, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_NEAREST)                    # src/dataset.cpp:128
out["min_eig"] = cv2.cornerMinEigenVal(img, 3, ksize=3)
rng = np.random.RandomState(0)
occ = np.stack([rng.rand(6) * 160, rng.rand(6) * 96], 1).astype(np.float32)
mask = np.full(img.shape, 255, np.uint8)
for (x, y) in occ:                                                                                         # src/frontend.cpp:42-47
    cv2.rectangle(mask, (int(np.rint(x - np.float32(10))), int(np.rint(y - np.float32(10)))),
                  (int(np.rint(x + np.float32(10))), int(np.rint(y + np.float32(10)))), 0, -1)
kps = cv2.GFTTDetector_create(150, 0.01, 20).detect(img, mask)                                            # src/frontend.cpp:24,51
out["occ"], out["mask"] = occ, mask
out["gftt_xy"] = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
out["gftt_resp"] = np.array([k.response for k in kps], np.float32)
out["pyr1"] = cv2.pyrDown(img)
a, b = moved_pair(96, 160, 12)
p0 = np.array([k.pt for k in cv2.GFTTDetector_create(60, 0.01, 8).detect(a, None)], np.float32).reshape(-1, 2)
init = (p0 + np.array([2.5, -1.5], np.float32)).astype(np.float32)
p1, st, _ = cv2.calcOpticalFlowPyrLK(a, b, p0, init.copy(), winSize=(11, 11), maxLevel=3,                 # src/frontend.cpp:353-357
                                     criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                     flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
out.update(lk_a=a, lk_b=b, lk_p0=p0, lk_init=init, lk_p1=p1, lk_status=st.ravel())
l, r = stereo_pair(64, 200, 13)
out.update(bm_l=l, bm_r=r, bm_disp=cv2.StereoBM_create(128, 15).compute(l, r))                            # dense_reconstruction.cpp:89,114
bgr = rng.randint(0, 256, (20, 30, 3), np.uint8)
out.update(bgr=bgr, gray=cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
np.savez_compressed(os.path.join(HERE, "golden_cv.npz"), **out)
print("wrote golden_cv.npz", {k: getattr(v, "shape", None) for k, v in out.items()})
