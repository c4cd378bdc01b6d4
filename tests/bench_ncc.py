"""On-box timing of the NCC active search alone (BASELINE.json config 5: 1280x720 pyramid, N = 100..2000 features, every
feature searched).  Prints per-launch time, features/s, candidate evaluations/s and the shared-memory read rate they imply
(2 x 121 byte reads per candidate) beside the window bytes fetched from HBM / L2.
usage: python tests/bench_ncc.py [N ...]   (lives under tests/ because it takes its templates from the oracle's helpers)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario
from oracle import ncc_oracle
from test_ncc import feature_pixels, render, smooth_textures

W, H = 1280, 720
Ns = [int(a) for a in sys.argv[1:]] or [100, 200, 500, 1000, 1500, 2000]
for N in Ns:
    sc = Scenario(W, H, N)
    x, P, ft, fo, desc, uv0 = sc.init_map()
    rng = np.random.default_rng(5)
    tex = smooth_textures(rng, N)
    tmpl = ncc_oracle.cut_templates(ncc_oracle.pyramid(render(W, H, uv0, tex, 1)), uv0)
    gpu = EkfBatch(sc.params, 1, N, 64)
    gpu.set_state(0, x, P, ft, fo, desc)
    gpu.ncc_set_templates(0, 0, tmpl)
    gpu.ncc_set_image(0, render(W, H, feature_pixels(sc, 1), tex, 2))
    gpu.predict(); gpu.measure()
    for _ in range(3):
        gpu.match_ncc(0.8)
    gpu.sync()
    reps = 20
    gpu.timer_record(0)
    for _ in range(reps):
        gpu.match_ncc(0.8)
    gpu.timer_record(1)
    gpu.sync()
    ms = gpu.timer_elapsed_ms(0, 1) / reps
    r = gpu.feature_results(0)
    s, lv = gpu.ncc_scores(0)
    # candidate count: the gate area at the start level + 9 per refinement level (upper bound: the square window)
    cand = 0
    for j in range(N):
        if lv[j] < 0:
            continue
        cand += 625 + 9 * int(lv[j])
    print(json.dumps({"W": W, "H": H, "N": N, "us_per_launch_incl_list_kernel": round(ms * 1e3, 1), "features_per_s": round(N / ms * 1e3),
                      "matched": int(r["matched"].sum()), "start_levels": np.bincount(lv[lv >= 0], minlength=3).tolist(),
                      "candidates_upper_bound": int(cand), "smem_read_GBs_upper_bound": round(cand * 242 / ms / 1e6, 1),
                      "window_bytes_from_l2": int((lv >= 0).sum()) * 36 * 64}))
