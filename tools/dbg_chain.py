"""Phase clocks of one mid-chain step (I = 5) of the critical CTA of the fused chain (csrc/ekf_chain.cuh), C3 size."""
import ctypes, sys
sys.path.insert(0, ".")
import numpy as np
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario
sc = Scenario(640, 480, 500)
gpu = EkfBatch(sc.params, 1, 500, 1300)
x, P, ft, fo, desc, _ = sc.init_map()
gpu.set_state(0, x, P, ft, fo, desc)
for t in range(1, 40):
    kp, ds = sc.frame(t); gpu.set_keypoints(0, kp, ds); gpu.step()
out = np.zeros(64, np.int64)
gpu.L.ekfb_debug_read(gpu.h, out.ctypes.data_as(ctypes.c_void_p))
d = out[48:58]
names = ["wait+load", "X gemm", "store X + signal", "update", "nu/pad", "factor", "publish", "signal"]
print("critical CTA, step I=5 (prefetched=%d): " % d[9] + ", ".join(f"{n} {d[i + 1] - d[i]}" for i, n in enumerate(names)) + f", total {d[8] - d[0]} cycles")
print(gpu.frame_info(0))
