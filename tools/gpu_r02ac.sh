#!/bin/bash
O=gpurun_out/r02ad
mkdir -p $O
timeout 400 python bench.py --workload c3k1000 --no-c4-leg --no-cpu-baseline > $O/bench_c3_k1000.json 2> $O/bench_c3_k1000.err
timeout 200 python - > $O/quick_c3_k1000.txt 2>&1 <<'PY'
import json, sys
sys.path.insert(0, '.')
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario
W, H, N, T = 640, 480, 500, 80
sc = Scenario(W, H, N, clutter_ratio=0.0, outlier_frac=0.0, noise_px=0.1, flip_p=0.0)
gpu = EkfBatch(sc.params, 1, N, 2 * N + 256)
x, P, ft, fo, desc, _ = sc.init_map()
gpu.set_state(0, x, P, ft, fo, desc)
gpu.load_sequence(0, [sc.frame(t) for t in range(1, T + 1)])
for t in range(T // 2):
    gpu.select_frame(t); gpu.step()
gpu.sync()
gpu.profile_enable(True)
for t in range(T // 2, T):
    gpu.select_frame(t); gpu.step()
pm, pl = gpu.profile_read()
print(json.dumps({"info": gpu.frame_info(0), "group_ms_per_frame": {k: round(v / (T - T // 2), 4) for k, v in pm.items()}, "group_launches": {k: v / (T - T // 2) for k, v in pl.items()}}))
PY
cut -c1-900 $O/bench_c3_k1000.json; tail -2 $O/bench_c3_k1000.err; cat $O/quick_c3_k1000.txt | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_buffered or schain or numeric" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
