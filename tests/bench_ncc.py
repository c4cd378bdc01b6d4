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


# ---- the whole frame with the NCC matcher (EKFB_OPT_MATCHER = 1: image upload + pyramid + ekfb_step) beside the descriptor matcher
# on the same scene: 1280x720, N features, templates + anchors as ekfb_add_features captures them, 12 frames after 6 warm-up frames
for N in [n for n in Ns if n <= 1000][:3]:
    sc = Scenario(W, H, N)
    x, P, ft, fo, desc, uv0 = sc.init_map()
    rng = np.random.default_rng(5)
    tex = smooth_textures(rng, N)
    img0 = render(W, H, uv0, tex, 1)
    tmpl = ncc_oracle.cut_templates(ncc_oracle.pyramid(img0), uv0)
    res = {}
    for name, matcher in (("descriptor", 0), ("ncc", 1)):
        gpu = EkfBatch(sc.params, 1, N, 2 * N + 256)
        gpu.set_state(0, x, P, ft, fo, desc)
        gpu.ncc_set_templates(0, 0, tmpl)
        gpu.ncc_set_anchors(0, 0, ncc_oracle.anchors(x, uv0))
        gpu.set_option(14, matcher)
        frames = [sc.frame(t) for t in range(1, 19)]
        imgs = [render(W, H, feature_pixels(sc, t), tex, 40 + t) for t in range(1, 19)]
        ms, inl = [], []
        for t in range(18):
            gpu.sync()
            gpu.timer_record(0)
            if matcher:
                gpu.ncc_set_image(0, imgs[t])           # H2D of the frame + the pyramid kernels
                gpu.set_keypoints(0, np.zeros((0, 2), np.float32), np.zeros((0, 32), np.uint8))
            else:
                gpu.set_keypoints(0, *frames[t])
            gpu.timer_record(1)
            gpu.step()
            gpu.timer_record(2)
            gpu.sync()
            if t >= 6:
                ms.append((gpu.timer_elapsed_ms(0, 1), gpu.timer_elapsed_ms(1, 2)))
                inl.append(gpu.frame_info(0)["n_inliers"])
        ms = np.array(ms)
        res[name] = {"input_ms": round(float(ms[:, 0].mean()), 4), "step_ms": round(float(ms[:, 1].mean()), 4), "inliers_mean": float(np.mean(inl))}
        gpu.close()
    print(json.dumps({"W": W, "H": H, "N": N, "frame_with_matcher": res}))
