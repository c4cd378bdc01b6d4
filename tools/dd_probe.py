"""Timing probe of the TMA-fed downdate: what bounds it at a given (n, k)?  Option 13 switches parts of the kernel off
(1 = no DMMA, 2 = no stores, 4 = no mirror store; the results are wrong then -- timing only).  usage: dd_probe.py N k [k ...]"""
import json
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario

N = int(sys.argv[1])
ks = [int(a) for a in sys.argv[2:]]
n = 13 + 6 * N
rng = np.random.default_rng(0)
for flush in (True, False):
    gpu = EkfBatch(Scenario(640, 480, 4).params, 1, N, 64)
    if flush:
        gpu.flush_l2()
    A = rng.normal(size=(n, 8))
    P = A @ A.T + np.eye(n)
    for k in ks:
        Wt = rng.normal(size=(k, n)) * 0.01
        res = {}
        for name, probe, ctas in (("full", 0, 2), ("no_dmma", 1, 2), ("no_stores", 2, 2), ("no_mirror", 4, 2), ("loads_only", 3, 2), ("full_1cta", 0, 1)):
            gpu.set_option(13, probe)
            gpu.set_option(12, ctas)
            ts = []
            for _ in range(5):
                gpu.downdate_timing(True)
                gpu.test_downdate(P, Wt)
                ts.append(gpu.downdate_stats()["ms"])
            res[name] = round(min(ts) * 1e3, 1)
        print(json.dumps({"n": n, "k": k, "l2_flushed": flush, "us": res}))
    gpu.close()
