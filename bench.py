#!/usr/bin/env python
"""bench.py -- EKF frames/s of the B200-native hot path (BASELINE.json metric) on synthetic sequences.

    python bench.py --gpus N --steps K --warmup W [--workload c3|c2|c4|c5] [--impl b200|reference]

A "step" is one frame of the per-frame hot path (EKF::step order: predict, measure, match, 1-point
RANSAC, low-innovation update, rescue, high-innovation update, map-feature bookkeeping) over this
rank's filter(s).  Workloads (BASELINE.json configs):
  c3 (default)  640x480, 500 inverse-depth features (n = 3013), ONE filter per GPU.  With --gpus N the
                N ranks run N independent filters (seed offset = rank): weak scaling by filter instance.
  c2            320x240, 50 features, one filter per GPU.
  c4            256 filters of 640x480 / 200 features in total, sharded f mod N (strong scaling).
  c5            1280x720, --features N' features (update/search stress), one filter per GPU.
One JSON line on stdout (rank 0).  `value` = filter-frames/s with the frame's keypoints already resident
in HBM; `e2e` = the same through the public C ABI with host buffers (H2D of keypoints + D2H of the result
record inside the timed region).  L2 is flushed (256 MB memset) before every timed iteration, outside the
timed interval; device time by CUDA events on the library's stream, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(W=320, H=240, N=50, filters=1, desc="synthetic 320x240, 50 inverse-depth features, single filter"),
    "c3": dict(W=640, H=480, N=500, filters=1, desc="synthetic 640x480, 500 features, single filter (P 3013x3013 FP64)"),
    "c4": dict(W=640, H=480, N=200, filters=256, desc="256 independent 640x480 / 200-feature filters, sharded f mod G"),
    "c5": dict(W=1280, H=720, N=1000, filters=1, desc="1280x720 stress, N features, single filter"),
}
METRIC = "EKF frames/sec at N features"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--features", type=int, default=0, help="override the feature count (c5 sweep)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--filter-warm", type=int, default=30, help="frames run before the warm-up so the filter has converged")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 8:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": sorted(reasons),
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ------------------------------------------------------------------------------------------------
def make_scenarios(wl, my_filters, features):
    from openekfmonoslam_b200.scenario import Scenario
    N = features or wl["N"]
    uniq = sorted(set(my_filters))
    # at most 8 distinct scenes per rank; further filters reuse them (identical work, different index)
    scenes = {}
    for i, f in enumerate(uniq):
        scenes[f] = Scenario(wl["W"], wl["H"], N, seed_offset=f) if i < 8 else scenes[uniq[i % 8]]
    return scenes, N


def cpu_oracle_sample(wl, N, budget_s=25.0, max_frames=20):
    """Times the reference's CPU path (1 thread -- the reference is single-threaded) on the first frames of the same
    workload until ~budget_s of CPU time is spent; returns frames/s and the sample description.
    kind "reference": oracle/_ref/libref.so, the reference's own sources (EKF::step and everything below it)
    compiled against oracle/cvshim -- used when it was built (needs /root/reference at build time; the .so travels to
    the GPU box).  kind "port": the oracle restatement otherwise."""
    from openekfmonoslam_b200.scenario import Scenario
    from oracle import ref_lib
    sc = Scenario(wl["W"], wl["H"], N)
    x, P, ft, fo, desc, _ = sc.init_map()
    use_ref = ref_lib.available(build=False)
    if use_ref:
        f = ref_lib.ReferenceFilter(sc.params)
    else:
        from oracle.oracle_lib import OracleFilter
        f = OracleFilter(sc.params)
    f.set_state(x, P, ft, fo, desc)
    frames, spent, phases = 0, 0.0, {}
    while frames < max_frames and (frames == 0 or spent < budget_s):
        kp, ds = sc.frame(frames + 1)
        t0 = time.perf_counter()
        info = f.step(kp, ds)
        spent += time.perf_counter() - t0
        frames += 1
        for k, v in (info or {}).items():
            if k.startswith("us_"):
                phases[k[3:]] = phases.get(k[3:], 0.0) + v
    what = ("reference sources (EKF::step) compiled against oracle/cvshim, g++ -O2" if use_ref
            else "oracle (literal restatement of the reference algorithm, g++ -O2)")
    out = dict(value=frames / spent, unit=UNIT, cores=1, kind="reference" if use_ref else "port",
               sample=f"{what}, 1 thread, frames 1..{frames} of the workload ({spent:.1f} s CPU); "
                      f"k differs slightly from the steady-state frames timed on the GPU")
    if phases:
        out["us_per_phase"] = {k: v / frames for k, v in phases.items()}
    out["dense_products_all_cores"] = dense_products_all_cores(13 + 6 * N, min(2 * int(0.65 * N), 13 + 6 * N))
    return out


def dense_products_all_cores(n, k):
    """Fairness line of SURVEY 8d: the reference's single-threaded cv::gemm is not what a tuned CPU library would do, so the dense
    products of its update (Update.cpp:92-109,214-218: P H^T, H (P H^T) + R, (P H^T) S^-1, (I - K H) P) are also timed with numpy /
    OpenBLAS on all host cores, once, at the workload's n and its typical low-innovation k.  Not the reference, not the target."""
    rng = np.random.default_rng(0)
    A = rng.normal(size=(n, 16))
    P = A @ A.T + np.eye(n)
    H = rng.normal(size=(k, n))
    t0 = time.perf_counter()
    PHt = P @ H.T
    S = H @ PHt + np.eye(k)
    K = PHt @ np.linalg.inv(S)
    P2 = (np.eye(n) - K @ H) @ P
    dt = time.perf_counter() - t0
    flops = 2.0 * n * n * k + 2.0 * n * k * k + 2.0 * n * k * k + 2.0 * k ** 3 + 2.0 * n * n * k + 2.0 * n ** 3
    return {"seconds_per_update": dt, "gflops": flops / dt / 1e9, "n": n, "k": k, "cores": os.cpu_count(),
            "what": "numpy/OpenBLAS, all cores, one low-innovation update in the reference's literal form", "checksum": float(P2[0, 0])}


def measure_fp64_peak(device):
    """cuBLAS DGEMM 8192^3 via torch (best of 10): the measured FP64 roofline denominator."""
    import torch
    a = torch.randn(8192, 8192, dtype=torch.float64, device=device)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU path on this workload, rank 0 only: oracle/_ref/libref.so (the reference's
    own sources compiled against oracle/cvshim) when it was built, else the oracle port."""
    if rank != 0:
        return
    N = args.features or wl["N"]
    budget = 25.0 if wl["N"] >= 200 else 10.0
    t0 = time.perf_counter()
    cb = cpu_oracle_sample(wl, N, budget_s=budget, max_frames=max(args.steps, 1))
    v = cb["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "features": N, "filters_total": 1,
                       "note": "single-threaded CPU path, one filter (the reference is single-threaded; no GPU used)"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    from openekfmonoslam_b200.capi import EkfBatch, RECORD_BYTES

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    total_filters = wl["filters"] if wl["filters"] > 1 else world   # c4: fixed total; others: one per GPU
    from openekfmonoslam_b200.sharding import assign_filters
    my_filters = assign_filters(total_filters, world, rank)
    F = len(my_filters)
    scenes, N = make_scenarios(wl, my_filters, args.features)
    T0, W_, K = args.filter_warm, args.warmup, args.steps
    n_frames = T0 + W_ + 2 * K                       # device-resident pass, then host-fed pass
    gpu = EkfBatch(next(iter(scenes.values())).params, F, N, 2 * N + 256, device=local)
    for kv in filter(None, os.environ.get("EKFB_OPTS", "").split(",")):   # developer switches (ekfb_set_option), e.g. "4=128"
        gpu.set_option(*(int(x) for x in kv.split("=")))
    frames_of = {}
    for f in sorted(set(my_filters)):
        sc = scenes[f]
        if id(sc) not in frames_of:
            frames_of[id(sc)] = [sc.frame(t) for t in range(1, n_frames + 1)]
    inits = {}
    for i, f in enumerate(my_filters):
        sc = scenes[f]
        if id(sc) not in inits:
            inits[id(sc)] = sc.init_map()
        x, P, ft, fo, desc, _ = inits[id(sc)]
        gpu.set_state(i, x, P, ft, fo, desc)
        gpu.load_sequence(i, frames_of[id(sc)])
    # pinned host copies of the host-fed frames (e2e leg)
    pin = {}
    for i, f in enumerate(my_filters):
        fr = frames_of[id(scenes[f])]
        pin[i] = [(torch.from_numpy(fr[t][0]).pin_memory(), torch.from_numpy(fr[t][1]).pin_memory())
                  for t in range(T0 + W_ + K, n_frames)]
    rec_dev = torch.zeros(F * RECORD_BYTES, dtype=torch.uint8, device=dev)
    rec_all = torch.zeros(world * F * RECORD_BYTES, dtype=torch.uint8, device=dev) if world > 1 else None
    flush = not args.no_l2_flush

    def barrier():
        gpu.sync()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def gather_records():
        """result gather across ranks: ~1.5 KB per filter over NCCL (the only collective of the path)"""
        gpu.write_records_device(rec_dev.data_ptr())
        gpu.sync()
        if world > 1:
            dist.all_gather_into_tensor(rec_all, rec_dev)

    # ---- filter convergence + warm-up (untimed) ----
    for t in range(T0 + W_):
        gpu.select_frame(t)
        gpu.step()
    gather_records()
    barrier()

    # ---- timed: K frames, keypoints resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    gpu.downdate_timing(True)
    l0 = gpu.kernel_launches()
    dev_ms = 0.0
    t_wall0 = time.perf_counter()
    for s in range(K):
        t = T0 + W_ + s
        if flush:
            gpu.flush_l2()
        gpu.select_frame(t)
        gpu.timer_record(0)
        gpu.step()
        if s % 16 == 15 or s == K - 1:
            gpu.write_records_device(rec_dev.data_ptr())
        gpu.timer_record(1)
        dev_ms += gpu.timer_elapsed_ms(0, 1)
        if world > 1 and (s % 16 == 15 or s == K - 1):
            dist.all_gather_into_tensor(rec_all, rec_dev)
    barrier()
    wall_value = time.perf_counter() - t_wall0
    launches = gpu.kernel_launches() - l0
    dd = gpu.downdate_stats()
    gpu.downdate_timing(False)
    info = gpu.frame_info(0)

    # ---- timed: K frames end to end (host keypoints in pinned memory -> H2D, step, D2H of the record) ----
    e2e_s, h2d = 0.0, 0
    for s in range(K):
        if flush:
            gpu.flush_l2()
        gpu.sync()
        t0 = time.perf_counter()
        cur = [pin[i][s] for i in range(F)]
        gpu.set_keypoints_batch_raw([xy.data_ptr() for xy, _ in cur], [ds.data_ptr() for _, ds in cur], [xy.shape[0] for xy, _ in cur])
        h2d += sum(xy.numel() * 4 + ds.numel() for xy, ds in cur)
        gpu.step()
        recs = gpu.records()                           # D2H + sync
        e2e_s += time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    barrier()
    assert abs(np.linalg.norm(np.array(recs[0].x_cam)[3:7]) - 1.0) < 1e-9

    # ---- max over ranks ----
    agg = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_max = float(agg[0]), float(agg[1])
    cnt = torch.tensor([float(F), float(launches), float(h2d)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    filters_total, launches_total, h2d_total = int(cnt[0]), int(cnt[1]), int(cnt[2])

    if rank == 0:
        value = filters_total * K / (dev_ms_max * 1e-3)
        e2e_value = filters_total * K / e2e_max
        peak = measure_fp64_peak(dev)
        tfl = dd["flops"] / (dd["ms"] * 1e-3) / 1e12 if dd["ms"] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True,
            "scaling": "strong" if wl["filters"] > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "features": N, "state_dim": 13 + 6 * N,
                       "filters_total": filters_total, "filters_per_gpu": F, "parallelism": f"independent filters x{world}",
                       "l2": "flushed before every timed iteration (256 MB memset, outside the timed interval)" if flush
                             else "not flushed", "filter_warm_frames": T0,
                       "last_frame": {k: info[k] for k in ("n_matches", "n_hypotheses", "n_inliers", "n_rescued")}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_total // K,
                    "d2h_bytes_per_step": filters_total * RECORD_BYTES},
            "gpu_launches": launches_total,
            "roofline": {"bound": "tensor", "kernel": "k_downdate64 (covariance downdate P -= W W^T: lower 64x64 tiles, 4 CTAs/SM, FP64 DMMA m8n8k4)",
                         "achieved": tfl, "peak": peak, "unit": "TFLOP/s", "frac": tfl / peak if peak else None,
                         "traffic": traffic, "launches": dd["launches"], "avg_launch_ms": dd["ms"] / max(dd["launches"], 1),
                         "algorithmic_flops_per_launch": dd["flops"] / max(dd["launches"], 1),
                         "peak_source": "measured in this run: torch.matmul float64 8192^3 (cuBLAS DGEMM), best of 10; "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "share_of_step": dd["ms"] / dev_ms if dev_ms > 0 else None},
            "clocks": sampler.summary(),
            "wall_s_timed_value_leg": wall_value,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_oracle_sample(wl, N, budget_s=25.0 if N >= 200 else 8.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
