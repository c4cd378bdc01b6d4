"""Quick on-box timing of ekfb_step for one workload; prints per-group milliseconds.  Not the bench.
usage: quick_time.py W H N F T [dbg]"""
import json
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario

W, H, N, F, T = (int(a) for a in sys.argv[1:6])
dbg = len(sys.argv) > 6
scs = [Scenario(W, H, N, seed_offset=i) for i in range(min(F, 4))]
gpu = EkfBatch(scs[0].params, F, N, 2 * N + 256)
t0 = time.time()
inits = [sc.init_map() for sc in scs]
frames = [[sc.frame(t) for t in range(1, T + 1)] for sc in scs]
for f in range(F):
    gpu.load_sequence(f, frames[f % len(scs)])
print("setup s", round(time.time() - t0, 1))
warm = T // 2


def reset_and_warm():
    for f in range(F):
        x, P, ft, fo, desc, _ = inits[f % len(scs)]
        gpu.set_state(f, x, P, ft, fo, desc)
    for t in range(warm):
        gpu.select_frame(t); gpu.step()
        if dbg:
            i = gpu.frame_info(0)
            print(t, {k: i[k] for k in ("n_predicted", "n_matches", "n_hypotheses", "n_inliers", "n_rescued", "status")})
    gpu.sync()


reset_and_warm()
gpu.timer_record(0)
l0 = gpu.kernel_launches()
tw = time.time()
for t in range(warm, T):
    gpu.select_frame(t); gpu.step()
gpu.timer_record(1)
gpu.sync()
wall = time.time() - tw
ms = gpu.timer_elapsed_ms(0, 1) / (T - warm)
print(json.dumps({"W": W, "H": H, "N": N, "F": F, "ms_per_frame": ms, "wall_ms": 1e3 * wall / (T - warm), "fps": 1e3 / ms * F,
                  "launches_per_frame": (gpu.kernel_launches() - l0) / (T - warm), "info": gpu.frame_info(0)}))
dbg = False
reset_and_warm()
gpu.profile_enable(True)
for t in range(warm, T):
    gpu.select_frame(t); gpu.step()
pm, pl = gpu.profile_read()
print(json.dumps({"group_ms_per_frame": {k: round(v / (T - warm), 4) for k, v in pm.items()},
                  "group_launches": {k: v / (T - warm) for k, v in pl.items()}}))
