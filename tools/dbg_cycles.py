import ctypes, sys
sys.path.insert(0, ".")
import numpy as np
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario
sc = Scenario(640, 480, 500)
gpu = EkfBatch(sc.params, 1, 500, 1300)
x, P, ft, fo, desc, _ = sc.init_map()
gpu.set_state(0, x, P, ft, fo, desc)
for t in range(1, 12):
    kp, ds = sc.frame(t); gpu.set_keypoints(0, kp, ds); gpu.step()
out = np.zeros(64, np.int64)
gpu.L.ekfb_debug_read(gpu.h, out.ctypes.data_as(ctypes.c_void_p))
for base in (16, 40):
  d = out[base:base+17]
  e = d[8:]
  print("rank8 b=0 warp0: loads %d dmma %d rmw %d to-barrier %d barrier-wait %d" % (e[4]-e[3], e[5]-e[4], e[6]-e[5], e[7]-e[6], e[8]-e[7]))
  print("pad/zero %d factor-only %d" % (d[6]-d[3], d[4]-d[6]))
  print("schain step J=1 fac CTA cycles: load %d X %d update %d factor %d store %d total %d | factor phases chol %d subst %d rank8 %d" % (d[1]-d[0], d[2]-d[1], d[3]-d[2], d[4]-d[3], d[5]-d[4], d[5]-d[0], d[8], d[9], d[10]))
print("cycles: load %d loop %d post %d phase2 %d total %d" % (out[1]-out[0], out[2]-out[1], out[3]-out[2], out[4]-out[3], out[4]-out[0]))
dc, dt = out[10] - out[8], out[11] - out[9]
print("downdate tile 40 main loop: %d cycles, %d ns -> %.0f MHz, K=%d, DMMA-bound cycles %d" % (dc, dt, dc / max(dt, 1) * 1e3, out[12], out[12] // 4 * 128 * 4))
print(gpu.frame_info(0))
