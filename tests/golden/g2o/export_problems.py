"""Writes tests/golden/g2o/g2o_problems.txt: the committed pose-only and window-BA problems of golden_geom.npz plus one pose
graph, as plain text for make_golden_g2o.cpp (which needs the real g2o; see CMakeLists.txt)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))
from util import K05, EXT_L, EXT_R, pose_graph_problem  # noqa: E402

G = np.load(os.path.join(os.path.dirname(HERE), "golden_geom.npz"))
f = lambda a: " ".join(repr(float(x)) for x in np.ravel(a))
with open(os.path.join(HERE, "g2o_problems.txt"), "w") as o:
    m = len(G["po_pts"])
    o.write("POSE_ONLY %d %s %s\n" % (m, f(G["po_K"]), f(G["po_T0"])))
    for p, uv in zip(G["po_pts"], G["po_uv"]):
        o.write("%s %s\n" % (f(p), f(uv)))
    N, L, E = len(G["ba_poses"]), len(G["ba_lms"]), len(G["ba_edge_kf"])
    o.write("BA %d %d %d %s %s %s %s %r\n" % (N, L, E, f(K05), f(K05), f(EXT_L), f(EXT_R), 5.991))
    for p in G["ba_poses"]:
        o.write(f(p) + "\n")
    for p in G["ba_lms"]:
        o.write(f(p) + "\n")
    for k, l, c, uv in zip(G["ba_edge_kf"], G["ba_edge_lm"], G["ba_edge_cam"], G["ba_edge_uv"]):
        o.write("%d %d %d %s\n" % (k, l, c, f(uv)))
    pr = pose_graph_problem(40, n=40, loops=((-1, 2),))
    o.write("POSE_GRAPH %d %d\n" % (len(pr["poses"]), len(pr["edge_a"])))
    for fx, p in zip(pr["fixed"], pr["poses"]):
        o.write("%d %s\n" % (fx, f(p)))
    for a, b, mm in zip(pr["edge_a"], pr["edge_b"], pr["meas"]):
        o.write("%d %d %s\n" % (a, b, f(mm)))
print("wrote g2o_problems.txt")
