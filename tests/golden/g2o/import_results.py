"""golden_g2o.txt (output of make_golden_g2o, i.e. of the REAL g2o + the reference's g2o_types.h) -> tests/golden/golden_g2o.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
tok = open(sys.argv[1]).read().split()
i, out = 0, {}
while i < len(tok):
    if tok[i] == "POSE_ONLY":
        m = int(tok[i + 1]); i += 2
        out["po_T"] = np.array(tok[i:i + 7], float); i += 7
        out["po_outlier"] = np.array(tok[i:i + m], np.uint8); i += m
    elif tok[i] == "BA":
        N, L, E = (int(x) for x in tok[i + 1:i + 4]); i += 4
        out["ba_P"] = np.array(tok[i:i + 7 * N], float).reshape(N, 7); i += 7 * N
        out["ba_L"] = np.array(tok[i:i + 3 * L], float).reshape(L, 3); i += 3 * L
        out["ba_chi2"] = np.array(tok[i:i + E], float); i += E
    elif tok[i] == "POSE_GRAPH":
        N = int(tok[i + 1]); i += 2
        out["pg_P"] = np.array(tok[i:i + 7 * N], float).reshape(N, 7); i += 7 * N
    else:
        raise SystemExit("unexpected token " + tok[i])
np.savez_compressed(os.path.join(os.path.dirname(HERE), "golden_g2o.npz"), **out)
print("wrote golden_g2o.npz", {k: v.shape for k, v in out.items()})
