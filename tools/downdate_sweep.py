"""On-box timing of the covariance downdate alone (P -= W W^T) over k, at a given n.  Prints per-launch time with the
L2 flushed before every launch, TFLOP/s of the symmetric form n(n+1)k and GB/s of the minimum traffic 16 n^2.
usage: downdate_sweep.py N_FEATURES [k ...]"""
import json
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario

N = int(sys.argv[1])
ks = [int(a) for a in sys.argv[2:]] or [32, 72, 160, 256, 640]
n = 13 + 6 * N
rng = np.random.default_rng(0)
gpu = EkfBatch(Scenario(640, 480, 4).params, 1, N, 64)
gpu.flush_l2()   # from here on ekfb_test_downdate evicts P and W from L2 before the timed launch
A = rng.normal(size=(n, 8))
P = A @ A.T + np.eye(n)
for k in ks:
    Wt = rng.normal(size=(k, n)) * 0.01
    res = {}
    for name, small_k, variant in (("tiles128", 0, 0), ("tiles64", 1 << 20, 0), ("tma_swizzle128", 1 << 20, 2), ("tma_dense", 1 << 20, 3)):
        gpu.set_option(4, small_k)
        gpu.set_option(2, variant)
        gpu.test_downdate(P, Wt)
        ts = []
        for _ in range(6):
            gpu.downdate_timing(True)
            gpu.test_downdate(P, Wt)
            ts.append(gpu.downdate_stats()["ms"])
        res[name] = round(min(ts) * 1e3, 1)
    best = min(res.values())
    print(json.dumps({"n": n, "k": k, "us": res, "tflops_best": round(n * (n + 1) * k / best / 1e6, 2),
                      "gbs_min_traffic_best": round(16 * n * n / best / 1e3, 1)}))
