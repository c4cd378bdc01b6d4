"""Generates tests/golden/frames_n16.npz from the CPU oracle: a 16-feature 320x240 scenario, initial state
and the per-frame inputs / outputs of 4 frames (state, covariance, match / inlier / rescued sets).  The GPU
parity test compares against these committed values; the CPU suite checks the oracle still reproduces them.
Run:  python tests/golden/make_frame_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openekfmonoslam_b200.scenario import Scenario  # noqa: E402
from oracle.oracle_lib import OracleFilter  # noqa: E402

sc = Scenario(320, 240, 16, seed_offset=3)
x, P, ft, fo, desc, _ = sc.init_map()
f = OracleFilter(sc.params)
f.set_state(x, P, ft, fo, desc)
out = dict(x0=x, P0=P, ftype=ft, foff=fo, desc=desc)
for t in range(1, 5):
    kp, ds = sc.frame(t)
    info = f.step(kp, ds)
    xs, Ps = f.get_state()
    out[f"kp{t}"], out[f"ds{t}"] = kp, ds
    out[f"x{t}"], out[f"P{t}"] = xs, Ps
    out[f"matched{t}"] = f.get_match()["matched"]
    out[f"mkp{t}"] = f.get_match()["kp"]
    out[f"inlier{t}"] = f.get_ransac()["inlier"]
    out[f"rescued{t}"] = f.get_rescue()
    out[f"counts{t}"] = np.array([info[k] for k in ("n_predicted", "n_matches", "n_hypotheses", "n_inliers", "n_rescued")])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frames_n16.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path))
