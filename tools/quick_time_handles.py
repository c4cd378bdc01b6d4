"""On-box experiment: F filters of the C4 shape on ONE GPU as H handles (each with its own stream) driven by H host threads,
so that the latency-bound phases of one sub-batch overlap the throughput-bound phases of the others.  Wall clock over all
handles.  usage: quick_time_handles.py F H [T]"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario

F, H = int(sys.argv[1]), int(sys.argv[2])
T = int(sys.argv[3]) if len(sys.argv) > 3 else 24
W, Hh, N = 640, 480, 200
scs = [Scenario(W, Hh, N, seed_offset=i) for i in range(4)]
inits = [sc.init_map() for sc in scs]
frames = [[sc.frame(t) for t in range(1, T + 1)] for sc in scs]
per = F // H
handles = []
for h in range(H):
    g = EkfBatch(scs[0].params, per, N, 2 * N + 256)
    for f in range(per):
        g.load_sequence(f, frames[(h * per + f) % 4])
        x, P, ft, fo, desc, _ = inits[(h * per + f) % 4]
        g.set_state(f, x, P, ft, fo, desc)
    handles.append(g)
warm = T // 2


def run(g, a, b, barrier):
    barrier.wait()
    for t in range(a, b):
        g.select_frame(t); g.step()
    g.sync()


def phase(a, b):
    bar = threading.Barrier(H + 1)
    th = [threading.Thread(target=run, args=(g, a, b, bar)) for g in handles]
    for t in th:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    return time.perf_counter() - t0


phase(0, warm)
dt = phase(warm, T)
print(json.dumps({"filters": F, "handles": H, "ms_per_frame_all_handles": round(1e3 * dt / (T - warm), 3),
                  "filter_frames_per_s": round(F * (T - warm) / dt)}))
