#!/bin/bash
O=gpurun_out/r02af
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_slab or single_buffered" > $O/pytest.log 2>&1; tail -12 $O/pytest.log
timeout 200 python - > $O/quick_grid.txt 2>&1 <<'PY'
import json, sys
import numpy as np
sys.path.insert(0, '.')
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario
W, H, N, T = 640, 480, 500, 80
for opts in ([], [(3, 0)]):
    sc = Scenario(W, H, N, clutter_ratio=0.0, outlier_frac=0.0, noise_px=0.1, flip_p=0.0)
    p = sc.params
    gx, gy = np.meshgrid(np.linspace(0.16 * 640, 0.84 * 640, 25), np.linspace(0.10 * 480, 0.90 * 480, 20))
    d = np.random.default_rng(3).uniform(3.0, 6.0, 500)
    sc.points = np.stack([(gx.ravel() - p.cx) / p.fx * d, (gy.ravel() - p.cy) / p.fy * d, d], axis=1)
    gpu = EkfBatch(sc.params, 1, N, 2 * N + 256)
    for o in opts: gpu.set_option(*o)
    x, P, ft, fo, desc, _ = sc.init_map()
    gpu.set_state(0, x, P, ft, fo, desc)
    gpu.load_sequence(0, [sc.frame(t) for t in range(1, T + 1)])
    for t in range(T // 2):
        gpu.select_frame(t); gpu.step()
    gpu.sync()
    gpu.timer_record(0)
    for t in range(T // 2, T):
        gpu.select_frame(t); gpu.step()
    gpu.timer_record(1); gpu.sync()
    ms = gpu.timer_elapsed_ms(0, 1) / (T - T // 2)
    gpu.profile_enable(True)
    for t in range(T // 2, T):
        gpu.select_frame(t); gpu.step()
    pm, pl = gpu.profile_read()
    print(json.dumps({"opts": opts, "ms_per_frame": ms, "fps": 1e3 / ms, "info": gpu.frame_info(0), "group_ms_per_frame": {k: round(v / (T - T // 2), 4) for k, v in pm.items()}}))
    gpu.close()
PY
cat $O/quick_grid.txt | cut -c1-700
