"""BASELINE.json config 1 on the box: the reference's bundled experiments/s3 sequence (keypoint fixture, tests/golden/
s3_keypoints.npz) through the sample driver of the drop-in EKF class (GPU, output.yml with the seven phase timers) and
through the reference's own EKF::init / EKF::step (oracle/_ref, CPU, 1 thread).  Prints one JSON line: parity over the
run, GPU microseconds per phase and frame, CPU milliseconds per frame.  usage: python tests/report_s3.py [out.json]   (lives under tests/ because it runs the reference arm from oracle/)"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2  # noqa: E402
from openekfmonoslam_b200 import build  # noqa: E402
from openekfmonoslam_b200.params import MapPolicy, write_config  # noqa: E402
from oracle import ref_lib  # noqa: E402
from test_host_ekf import write_kpseq  # noqa: E402
from test_s3_sequence import load_s3  # noqa: E402

p, frames = load_s3()
tmp = tempfile.mkdtemp()
cfg, seq, out = os.path.join(tmp, "config.yml"), os.path.join(tmp, "frames.kpseq"), tmp + "/"
write_config(cfg, p, 60, max_map_size=240,
             extra={"MapManagementFrequency": 1, "AlwaysRemoveUnseenMapFeatures": "true", "GoodFeatureMatchingPercent": 0.5,
                    "InverseDepthLinearityIndexThreshold": 0.1, "DetectNewFeaturesImageAreasDivideTimes": 2,
                    "DetectNewFeaturesImageMaskEllipseSize": 10})
write_kpseq(seq, frames)
build.build_host()
t0 = time.time()
run = subprocess.run([build.SAMPLE_OUT, cfg, seq, out], capture_output=True, text=True, timeout=600)
gpu_wall = time.time() - t0
assert run.returncode == 0, run.stderr
rows = [l.split() for l in run.stdout.splitlines() if l.startswith("STEP")]
fs = cv2.FileStorage(out + "output.yml", cv2.FILE_STORAGE_READ)
phases = ("Prediction", "Matching", "Ransac", "UpdateLI", "RescueOutliers", "UpdateHI", "MapManagement")
us = {k: [] for k in phases}
for key in fs.root().keys():
    node = fs.root().getNode(key)
    for k in phases:
        us[k].append(node.getNode(k).real())
fs.release()
ctypes.CDLL(None).srand(1)
r = ref_lib.ReferenceFilter(p)
r.set_policy(MapPolicy(60, 0, 240, 1, 0.5, 0.1), 1)
r.full_init(*frames[0])
worst, cpu_ms, sizes_equal = 0.0, [], True
for t in range(1, len(frames)):
    t1 = time.perf_counter()
    r.step(*frames[t])
    cpu_ms.append(1e3 * (time.perf_counter() - t1))
    xr, _ = r.get_state()
    n, N = r.dims()
    row = rows[t - 1]
    sizes_equal = sizes_equal and (int(row[25]), int(row[27])) == (N, n)
    xg = np.array([float(v) for v in row[9:22]])
    worst = max(worst, float(np.abs(xg - xr[:13]).max() / np.abs(xr[:13]).max()))
line = {"config": "experiments/s3 costado_recto1 frames 00090-00210, experiments/s3/config.yml, ORB keypoint fixture",
        "frames": len(rows), "map_sizes_equal_every_frame": bool(sizes_equal), "worst_camera_state_rel_err": worst,
        "final_map_features": int(rows[-1][25]), "final_state_dim": int(rows[-1][27]),
        "gpu_us_per_phase_mean": {k: round(float(np.mean(v[5:])), 1) for k, v in us.items()},
        "gpu_us_per_frame_mean": round(float(sum(np.mean(v[5:]) for v in us.values())), 1),
        "gpu_sample_wall_ms_per_frame_incl_process_start": round(1e3 * gpu_wall / len(rows), 2),
        "reference_cpu_ms_per_frame_mean": round(float(np.mean(cpu_ms[5:])), 3), "reference_cpu_threads": 1}
print(json.dumps(line))
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(json.dumps(line) + "\n")
