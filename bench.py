#!/usr/bin/env python
"""bench.py -- EKF frames/s of the B200-native hot path (BASELINE.json metric) on synthetic sequences.

    python bench.py --gpus N --steps K --warmup W [--workload c3|c3full|c3k1000|c2|c4|c5] [--impl b200|reference]

A "step" is one frame of the per-frame hot path (EKF::step order: predict, measure, match, 1-point
RANSAC, low-innovation update, rescue, high-innovation update, map-feature bookkeeping) over this
rank's filter(s).  Workloads (BASELINE.json configs):
  c3 (default)  640x480, 500 inverse-depth features (n = 3013), ONE filter per GPU (the configuration the >= 1 kHz target is
                quoted on).  With --gpus N the N ranks run N independent filters: weak scaling by filter instance.
                The default run ALSO executes a short leg of the north star's batched configuration -- c4: 256 filters of
                200 features in total, sharded f mod N -- and reports it in `config.c4_sharded` and `c4_sharded` of the
                same JSON line (strong scaling; `--no-c4-leg` skips it).
  c3full        c3 without clutter keypoints: every feature is matched (BASELINE.md's C3 row; k ~ 700 + the rescue update).
  c3k1000       c3 without clutter, outliers, descriptor flips: nearly all 500 features are low-innovation inliers (k ~ 1000).
  c2            320x240, 50 features, one filter per GPU.
  c4            the 256-filter batch as the main workload.
  c5            1280x720, --features N' features (update/search stress), one filter per GPU.
One JSON line on stdout (rank 0).  `value` = filter-frames/s with the frame's keypoints already resident in HBM;
`e2e` = the same through the public C ABI with host buffers (one packed H2D of the keypoints + D2H of the result records
inside the timed region).  Timing: CUDA events on the library's stream around every step (L2 flushed by a 256 MB memset
before every step, outside the timed interval) plus CUDA events around every NCCL gather of the result records; exactly K
steps form a block, blocks are repeated until >= ~0.6 s is timed, the MEDIAN block is reported, max over ranks.
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
    "c3full": dict(W=640, H=480, N=500, filters=1, clutter=0.0,
                   desc="synthetic 640x480, 500 features, single filter, no clutter keypoints: every feature matched (BASELINE.md C3 row)"),
    "c3k1000": dict(W=640, H=480, N=500, filters=1, clutter=0.0, outliers=0.0, flips=0.0, noise=0.1,
                    desc="synthetic 640x480, 500 features, single filter, no clutter, no outlier displacements, no descriptor bit "
                         "flips, 0.1 px noise: (nearly) every feature is a low-innovation inlier, k ~ 1000 update rows"),
    "c4": dict(W=640, H=480, N=200, filters=256, desc="256 independent 640x480 / 200-feature filters, sharded f mod G"),
    "c5": dict(W=1280, H=720, N=1000, filters=1, desc="1280x720 stress, N features, single filter"),
}
METRIC = "EKF frames/sec at N features"
UNIT = "frames/s"
GATHER_EVERY = 16      # frames between two gathers of the result records (T >= 16)
MIN_TIMED_S = 0.6
MAX_BLOCKS = 40


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
    ap.add_argument("--no-c4-leg", action="store_true", help="skip the batched 256-filter leg of the default run")
    ap.add_argument("--c4-steps", type=int, default=10, help="steps per block of the c4 leg")
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
def scenario_for(wl, N, seed_offset=0):
    from openekfmonoslam_b200.scenario import Scenario
    return Scenario(wl["W"], wl["H"], N, seed_offset=seed_offset, clutter_ratio=wl.get("clutter", 1.0),
                    outlier_frac=wl.get("outliers", 0.10), noise_px=wl.get("noise", 0.3), flip_p=wl.get("flips", 0.05))


def make_scenarios(wl, my_filters, features):
    N = features or wl["N"]
    uniq = sorted(set(my_filters))
    # at most 8 distinct scenes per rank; further filters reuse them (identical work, different index)
    scenes = {}
    for i, f in enumerate(uniq):
        scenes[f] = scenario_for(wl, N, f) if i < 8 else scenes[uniq[i % 8]]
    return scenes, N


def reference_filter(params):
    """(filter object, kind, description): oracle/_ref/libref.so -- the reference's own sources (EKF::step and everything
    below it) compiled against oracle/cvshim -- when it was built (needs /root/reference at build time; the .so travels to
    the GPU box), else the oracle restatement.  TEST INFRASTRUCTURE: the checker / CPU baseline, never the product path."""
    from oracle import ref_lib
    if ref_lib.available(build=False):
        return ref_lib.ReferenceFilter(params), "reference", "reference sources (EKF::step phases) compiled against oracle/cvshim, g++ -O2"
    from oracle.oracle_lib import OracleFilter
    return OracleFilter(params), "port", "oracle (literal restatement of the reference algorithm, g++ -O2)"


def reference_frame(r, kind, kp, ds):
    """one frame on the CPU arm; returns (seconds, matched, z, inlier, rescued)"""
    t0 = time.perf_counter()
    if kind == "reference":
        # the reference's own phase functions in EKF::step order (bit-identical to EKF::step, tests/test_s3_sequence.py)
        r.predict(); r.measure(); r.match(kp, ds); r.ransac(); r.update_li(); r.rescue(); r.update_hi()
        dt = time.perf_counter() - t0
        m, s = r.get_match(), r.get_sets()
        t0 = time.perf_counter()
        r.update_map_features()
        dt += time.perf_counter() - t0
        return dt, m["matched"], m["z"], s["inlier"], s["rescued"]
    r.step(kp, ds)
    dt = time.perf_counter() - t0
    m, ro = r.get_match(), r.get_ransac()
    return dt, m["matched"], m["z"], ro["inlier"], r.get_rescue()


def cpu_baseline_and_parity(wl, N, sc, gpu, first_frame, n_frames):
    """The reference's CPU path (1 thread -- the reference is single-threaded) on `n_frames` steady-state frames that start
    from the GPU filter's current state, timed; the GPU runs the same frames from the same state and the results are
    compared: parity at the benchmarked size, in the same run."""
    def rel_err(a, b):   # max|a-b| / max|b|: the norm BASELINE's 1e-9 is applied in (tests/conftest.py)
        return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-300)

    xg, Pg = gpu.get_state(0)
    dg, _, _ = gpu.get_descriptors(0)
    ft = np.full(N, 2, np.int32); fo = (13 + 6 * np.arange(N)).astype(np.int32)
    r, kind, what = reference_filter(sc.params)
    r.set_state(xg, Pg, ft, fo, dg)
    gpu.set_state(0, xg, Pg, ft, fo, dg)
    spent, sets_equal, ex, eP, ks = 0.0, True, 0.0, 0.0, []
    for t in range(first_frame, first_frame + n_frames):
        kp, ds = sc.frame(t)
        dt, matched, z, inl, resc = reference_frame(r, kind, kp, ds)
        spent += dt
        gpu.set_keypoints(0, kp, ds); gpu.step()
        g = gpu.feature_results(0)
        mm = matched.astype(bool)
        sets_equal = sets_equal and bool(np.array_equal(g["matched"], matched) and np.array_equal(g["z"][mm], z[mm]) and
                                         np.array_equal(g["inlier"], inl) and np.array_equal(g["rescued"], resc))
        xr, Pr = r.get_state()
        x2, P2 = gpu.get_state(0)
        ex, eP = max(ex, rel_err(x2, xr)), max(eP, rel_err(P2, Pr))
        ks.append([int(2 * inl.sum()), int(2 * resc.sum())])
    if hasattr(r, "close"):
        r.close()
    out = dict(value=n_frames / spent, unit=UNIT, cores=1, kind=kind,
               sample=f"{what}, 1 thread, {n_frames} steady-state frames ({first_frame}..{first_frame + n_frames - 1}) started from the GPU "
                      f"filter's state after its warm-up, {spent:.1f} s CPU; update rows [low, high innovation] per frame {ks}",
               frames_timed=n_frames, seconds=spent)
    parity = dict(frames=n_frames, sets_equal=sets_equal, state_rel_err=ex, cov_rel_err=eP, tolerance=1e-9,
                  ok=bool(sets_equal and ex < 1e-9 and eP < 1e-9),
                  what="matched set + matched pixels + inlier + rescued sets exact; state and full covariance max|a-b|/max|b| "
                       f"against the CPU arm ({kind}) on the same frames from the same state")
    out["dense_products_all_cores"] = dense_products_all_cores(13 + 6 * N, min(2 * int(0.65 * N), 13 + 6 * N))
    return out, parity


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


def run_reference(args, wl, rank, world, json_fd):
    """--impl reference: the reference's CPU path on this workload, rank 0 only: oracle/_ref/libref.so (the reference's own
    sources compiled against oracle/cvshim) when it was built, else the oracle port.  The timed frames are steady-state frames:
    the filter state they start from is produced by running the warm-up frames first -- on the GPU library when a device is
    there (the CPU arm needs ~20 s per C3 frame; the state provider is not part of the timed path), else the frames are the
    cold frames 1.. of the sequence and the line says so."""
    if rank != 0:
        return
    N = args.features or wl["N"]
    sc = scenario_for(wl, N)
    x, P, ft, fo, desc, _ = sc.init_map()
    T0 = args.filter_warm + args.warmup
    warm = "cold frames from the initial map (no CUDA device to produce the warmed-up state)"
    first = 1
    try:
        import torch
        if torch.cuda.is_available():
            from openekfmonoslam_b200.capi import EkfBatch
            g = EkfBatch(sc.params, 1, N, 2 * N + 256, device=int(os.environ.get("LOCAL_RANK", "0")))
            g.set_state(0, x, P, ft, fo, desc)
            for t in range(1, T0 + 1):
                g.set_keypoints(0, *sc.frame(t)); g.step()
            x, P = g.get_state(0)
            desc, _, _ = g.get_descriptors(0)
            g.close()
            first = T0 + 1
            warm = f"state after {T0} warm-up frames (same frames as the b200 arm's warm-up)"
    except Exception as exc:   # the reference arm must not depend on the product
        warm += f" [{type(exc).__name__}]"
    r, kind, what = reference_filter(sc.params)
    r.set_state(x, P, ft, fo, desc)
    per_frame_s = 20.0 if N >= 400 else (3.0 if N >= 150 else 0.05)
    frames = int(max(3, min(args.steps, 75.0 / per_frame_s)))
    t_wall = time.perf_counter()
    spent, ks = 0.0, []
    for t in range(first, first + frames):
        dt, matched, z, inl, resc = reference_frame(r, kind, *sc.frame(t))
        spent += dt
        ks.append([int(2 * inl.sum()), int(2 * resc.sum())])
    v = frames / spent
    cb = dict(value=v, unit=UNIT, cores=1, kind=kind, frames_timed=frames, seconds=spent,
              sample=f"{what}, 1 thread, {frames} frames ({first}..{first + frames - 1}) from the {warm}; update rows "
                     f"[low, high innovation] per frame {ks}")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": frames,
            "steps_requested": args.steps, "warmup": 0, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "features": N, "state_dim": 13 + 6 * N, "filters_total": 1,
                       "filters_per_gpu": 1,
                       "note": "single-threaded CPU path, one filter (the reference is single-threaded; no GPU on the timed path); "
                               "`steps` is the number of frames actually timed"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_wall}
    emit(json_fd, line)


# ------------------------------------------------------------------------------------------------
class Leg:
    """One workload on this rank's shard of filters: setup, warm-up, timed blocks (device resident), e2e blocks (host fed)."""

    def __init__(self, args, wl, features, rank, world, local, dev, K, W_, T0, flush):
        import torch
        from openekfmonoslam_b200.capi import EkfBatch, RECORD_BYTES
        from openekfmonoslam_b200.sharding import assign_filters
        self.torch, self.args, self.wl, self.rank, self.world, self.dev = torch, args, wl, rank, world, dev
        self.K, self.W_, self.T0, self.flush = K, W_, T0, flush
        self.RECORD_BYTES = RECORD_BYTES
        self.total_filters = wl["filters"] if wl["filters"] > 1 else world   # c4: fixed total; others: one per GPU
        self.my_filters = assign_filters(self.total_filters, world, rank)
        self.F = len(self.my_filters)
        self.scenes, self.N = make_scenarios(wl, self.my_filters, features)
        self.gpu = EkfBatch(next(iter(self.scenes.values())).params, max(self.F, 1), self.N, 2 * self.N + 256, device=local)
        self.rec_dev = torch.zeros(max(self.F, 1) * RECORD_BYTES, dtype=torch.uint8, device=dev)
        self.n_max = (self.total_filters + world - 1) // world
        self.rec_pad = torch.zeros(self.n_max * RECORD_BYTES, dtype=torch.uint8, device=dev)
        self.rec_all = torch.zeros(world * self.n_max * RECORD_BYTES, dtype=torch.uint8, device=dev) if world > 1 else None
        self.scene_list = []      # distinct scenes of this rank, and for every local filter the index of its scene
        self.scene_of = []
        for f in self.my_filters:
            sc = self.scenes[f]
            if not any(sc is s for s in self.scene_list):
                self.scene_list.append(sc)
            self.scene_of.append([i for i, s in enumerate(self.scene_list) if s is sc][0])
        inits = [sc.init_map() for sc in self.scene_list]
        for i in range(self.F):
            x, P, ft, fo, desc, _ = inits[self.scene_of[i]]
            self.gpu.set_state(i, x, P, ft, fo, desc)
        self.t_next = 1     # next frame number of the synthetic sequence

    def load(self, count):
        """generate frames t_next .. t_next+count-1 and make them the device-resident sequence; returns them (per scene)"""
        fr = [[sc.frame(t) for t in range(self.t_next, self.t_next + count)] for sc in self.scene_list]
        for i in range(self.F):
            self.gpu.load_sequence(i, fr[self.scene_of[i]])
        self.t_next += count
        return fr

    def barrier(self):
        import torch.distributed as dist
        self.gpu.sync()
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def gather(self):
        """the only collective of the path: all_gather of the ~1.5 KB per-filter records over NCCL; returns its device time [ms]"""
        import torch.distributed as dist
        if self.world == 1:
            return 0.0
        torch = self.torch
        self.gpu.sync()     # the records were written on the library's stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.rec_pad[: self.F * self.RECORD_BYTES] = self.rec_dev[: self.F * self.RECORD_BYTES]
        dist.all_gather_into_tensor(self.rec_all, self.rec_pad)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    def warm(self):
        fr = self.load(self.T0 + self.W_)
        del fr
        t0 = time.perf_counter()
        for t in range(self.T0 + self.W_):
            self.gpu.select_frame(t)
            self.gpu.step()
        self.gpu.sync()
        est = (time.perf_counter() - t0) / max(self.T0 + self.W_, 1)
        self.gpu.write_records_device(self.rec_dev.data_ptr())
        self.gather()
        self.barrier()
        return est

    def timed_blocks(self, R):
        """R blocks of exactly K steps; returns per-block device ms (steps + gathers), gather ms list, launches"""
        gpu, K = self.gpu, self.K
        self.load(R * K)
        self.barrier()
        blocks, gathers = [], []
        l0 = gpu.kernel_launches()
        gpu.downdate_timing(True)
        for b in range(R):
            ms = 0.0
            for s in range(K):
                if self.flush:
                    gpu.flush_l2()
                gpu.select_frame(b * K + s)
                due = (s % GATHER_EVERY == GATHER_EVERY - 1) or s == K - 1
                gpu.timer_record(0)
                gpu.step()
                if due:
                    gpu.write_records_device(self.rec_dev.data_ptr())
                gpu.timer_record(1)
                ms += gpu.timer_elapsed_ms(0, 1)
                if due:
                    g = self.gather()
                    gathers.append(g)
                    ms += g
            blocks.append(ms)
        self.barrier()
        launches = gpu.kernel_launches() - l0
        lms, lfl = gpu.downdate_launches()
        dd = gpu.downdate_stats()
        gpu.downdate_timing(False)
        # low-innovation launches (the bulk of the flop) and high-innovation launches (few rows: bandwidth / latency bound) apart
        if len(lms):
            big = lfl >= 0.5 * lfl.max()
            for name, sel in (("low_innovation", big), ("high_innovation", ~big)):
                if sel.any():
                    dd[name] = {"launches": int(sel.sum()), "avg_launch_ms": float(lms[sel].mean()),
                                "algorithmic_flops_per_launch": float(lfl[sel].mean()),
                                "tflops": float(lfl[sel].sum() / (lms[sel].sum() * 1e-3) / 1e12)}
        return blocks, gathers, launches, dd

    def e2e_blocks(self, R):
        """R blocks of K steps fed from pinned host memory: ONE packed H2D pair per step for all filters of the handle
        (ekfb_set_keypoints_packed), the step, D2H of the result records (+ the gather when due).  Host wall clock per step, from
        before the H2D is issued until the D2H has completed; the L2 flush between steps (timing hygiene) is outside the clock."""
        torch, gpu, K, F = self.torch, self.gpu, self.K, self.F
        fr = [[sc.frame(t) for t in range(self.t_next, self.t_next + R * K)] for sc in self.scene_list]
        self.t_next += R * K
        pins = []
        for t in range(R * K):
            cur = [fr[self.scene_of[i]][t] for i in range(F)]
            off = np.cumsum([0] + [xy.shape[0] for xy, _ in cur]).astype(np.int32)
            xy = torch.from_numpy(np.ascontiguousarray(np.concatenate([c[0] for c in cur]))).pin_memory()
            ds = torch.from_numpy(np.ascontiguousarray(np.concatenate([c[1] for c in cur]))).pin_memory()
            pins.append((xy, ds, off))
        del fr
        self.barrier()
        blocks, h2d = [], 0
        recs = None
        for b in range(R):
            blk = 0.0
            for s in range(K):
                xy, ds, off = pins[b * K + s]
                if self.flush:          # timing hygiene, not part of the path: the flush and its completion stay off the clock
                    gpu.flush_l2()
                    gpu.sync()
                t0 = time.perf_counter()
                gpu.set_keypoints_packed_raw(xy.data_ptr(), ds.data_ptr(), off)
                h2d += xy.numel() * 4 + ds.numel()
                gpu.step()
                recs = gpu.records()                           # D2H + sync
                if self.world > 1 and ((s % GATHER_EVERY == GATHER_EVERY - 1) or s == K - 1):
                    gpu.write_records_device(self.rec_dev.data_ptr())
                    self.gather()
                blk += time.perf_counter() - t0
            blocks.append(blk)
        self.barrier()
        assert abs(np.linalg.norm(np.array(recs[0].x_cam)[3:7]) - 1.0) < 1e-9
        return blocks, h2d // (R * K)


def all_ranks(torch, dist, world, dev, values):
    """[world][len(values)] list of every rank's values"""
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    if world == 1:
        return [t.tolist()]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def run_leg(args, wl, features, rank, world, local, dev, K, W_, T0, flush):
    """returns (leg, result dict on every rank)"""
    import torch
    import torch.distributed as dist
    leg = Leg(args, wl, features, rank, world, local, dev, K, W_, T0, flush)
    est = leg.warm()
    est_max = max(r[0] for r in all_ranks(torch, dist, world, dev, [est]))
    R = int(min(MAX_BLOCKS, max(3, np.ceil(MIN_TIMED_S / max(K * est_max, 1e-6)))))
    blocks, gathers, launches, dd = leg.timed_blocks(R)
    info = leg.gpu.frame_info(0)
    Re = int(min(R, max(1, np.ceil(MIN_TIMED_S / max(K * est_max, 1e-6)))))
    e2e, h2d = leg.e2e_blocks(Re)
    med = float(np.median(blocks))
    e2e_med = float(np.median(e2e))
    per_rank = all_ranks(torch, dist, world, dev, [med, e2e_med, float(leg.F), float(launches), float(h2d),
                                                   float(np.mean(gathers)) if gathers else 0.0, float(min(blocks)), float(max(blocks))])
    dev_ms_max = max(r[0] for r in per_rank)
    e2e_max = max(r[1] for r in per_rank)
    filters_total = int(sum(r[2] for r in per_rank))
    res = dict(value=filters_total * K / (dev_ms_max * 1e-3), e2e_value=filters_total * K / e2e_max,
               ms_per_step=dev_ms_max / K, filters_total=filters_total, filters_per_gpu=leg.F, blocks=R, e2e_blocks=Re,
               launches_total=int(sum(r[3] for r in per_rank)), launches_per_step=sum(r[3] for r in per_rank) / (R * K * world),
               h2d_bytes_per_step=int(sum(r[4] for r in per_rank)),
               per_rank_ms_per_step=[r[0] / K for r in per_rank],
               per_rank_e2e_ms_per_step=[1e3 * r[1] / K for r in per_rank],
               collective_us=1e3 * max(r[5] for r in per_rank), gathers_per_block=len(gathers) // max(R, 1),
               block_ms_min_max=[min(r[6] for r in per_rank), max(r[7] for r in per_rank)],
               dd=dd, info=info, dev_ms_total=float(sum(blocks)), N=leg.N)
    return leg, res


def main():
    args = parse()
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner, ...) are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        run(args, json_fd)
    finally:
        sys.stdout.flush()


def emit(json_fd, line):
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def run(args, json_fd):
    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world, json_fd)
        return

    import torch
    import torch.distributed as dist
    from openekfmonoslam_b200.capi import RECORD_BYTES

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flush = not args.no_l2_flush
    K, W_, T0 = args.steps, args.warmup, args.filter_warm

    sampler = ClockSampler(local)
    sampler.start()
    t_wall0 = time.perf_counter()
    leg, res = run_leg(args, wl, args.features, rank, world, local, dev, K, W_, T0, flush)
    wall_main = time.perf_counter() - t_wall0
    sampler.stop_flag.set()       # clocks are sampled during the timed legs only, not while the CPU baseline runs
    sampler.join(timeout=2)
    N = res["N"]

    # ---- CPU baseline + parity at this size, same run (rank 0, N = 1, single-filter workloads) ----
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and wl["filters"] == 1:
        frames = 2 if N >= 400 else (4 if N >= 150 else 20)
        cpu, parity = cpu_baseline_and_parity(wl, N, leg.scene_list[0], leg.gpu, leg.t_next, frames)
    leg.gpu.close()
    del leg
    torch.cuda.empty_cache()

    # ---- the north star's batched configuration in the same run: 256 filters sharded f mod N ----
    c4 = None
    if args.workload == "c3" and not args.no_c4_leg and not args.features:
        sampler4 = ClockSampler(local)
        sampler4.start()
        leg4, r4 = run_leg(args, WORKLOADS["c4"], 0, rank, world, local, dev, args.c4_steps, max(3, min(W_, 5)), min(T0, 20), flush)
        c4 = {"workload": "c4", "desc": WORKLOADS["c4"]["desc"], "filter_frames_per_s": r4["value"], "unit": "filter-frames/s",
              "e2e": r4["e2e_value"], "ms_per_step": r4["ms_per_step"], "steps_per_block": args.c4_steps, "blocks": r4["blocks"],
              "filters_total": r4["filters_total"], "filters_per_gpu": r4["filters_per_gpu"], "scaling": "strong",
              "per_rank_ms_per_step": r4["per_rank_ms_per_step"], "collective_us": r4["collective_us"],
              "gathers_per_block": r4["gathers_per_block"], "h2d_bytes_per_step": r4["h2d_bytes_per_step"],
              "d2h_bytes_per_step": r4["filters_total"] * RECORD_BYTES,
              "last_frame_filter0": {k: r4["info"][k] for k in ("n_matches", "n_hypotheses", "n_inliers", "n_rescued", "status")},
              "timing": "device time of every step (CUDA events, library stream) + device time of every NCCL gather, median block, "
                        "max over ranks"}
        leg4.gpu.close()
        del leg4
        sampler4.stop_flag.set()
        sampler4.join(timeout=2)
        c4["clocks"] = sampler4.summary()

    if rank == 0:
        dd = res["dd"]
        peak = measure_fp64_peak(dev)
        tfl = dd["flops"] / (dd["ms"] * 1e-3) / 1e12 if dd["ms"] > 0 else 0.0
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get(args.workload)
            traffic_src = tj.get("_source")
        info = res["info"]
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if wl["filters"] > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "features": N, "state_dim": 13 + 6 * N,
                       "filters_total": res["filters_total"], "filters_per_gpu": res["filters_per_gpu"],
                       "parallelism": f"independent filters x{world}",
                       "l2": "flushed before every timed iteration (256 MB memset, outside the timed interval)" if flush
                             else "not flushed", "filter_warm_frames": T0,
                       "blocks": res["blocks"], "block_ms_min_max": res["block_ms_min_max"],
                       "timing": "exactly `steps` steps per block, blocks repeated until >= 0.6 s is timed, median block reported; "
                                 "device time of every step + of every NCCL gather (CUDA events); max over ranks",
                       "update_rows_last_frame": {"low_innovation_k": 2 * info["n_inliers"], "high_innovation_k": 2 * info["n_rescued"]},
                       "last_frame": {k: info[k] for k in ("n_matches", "n_hypotheses", "n_inliers", "n_rescued", "status")},
                       "c4_sharded": c4},
            "e2e": {"value": res["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": res["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": res["filters_total"] * RECORD_BYTES, "blocks": res["e2e_blocks"],
                    "what": "per step: pinned host keypoints -> one packed H2D pair -> ekfb_step -> D2H of the records (+ the NCCL gather when due), "
                            "host wall clock from before the H2D to after the D2H has completed, summed over the block's steps (the L2 flush "
                            "between steps and its completion are outside the clock), median block"},
            "gpu_launches": res["launches_total"], "gpu_launches_per_step": res["launches_per_step"],
            "per_rank_ms_per_step": res["per_rank_ms_per_step"], "collective_us": res["collective_us"],
            "roofline": {"bound": "tensor", "kernel": "k_downdate_tma (covariance downdate P -= W W^T: persistent CTAs take the lower 64x64 tiles from a dynamic "
                                                     "queue, TMA tensor loads / stores + mbarrier ring, FP64 DMMA m8n8k4)",
                         "achieved": tfl, "peak": peak, "unit": "TFLOP/s", "frac": tfl / peak if peak else None,
                         "traffic": traffic, "traffic_source": traffic_src, "launches": dd["launches"],
                         "avg_launch_ms": dd["ms"] / max(dd["launches"], 1),
                         "algorithmic_flops_per_launch": dd["flops"] / max(dd["launches"], 1),
                         "peak_source": "measured in this run: torch.matmul float64 8192^3 (cuBLAS DGEMM), best of 10; "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "share_of_step": dd["ms"] / res["dev_ms_total"] if res["dev_ms_total"] > 0 else None,
                         "by_update": {k2: dict(v2, frac=v2["tflops"] / peak) for k2, v2 in dd.items() if isinstance(v2, dict)}},
            "clocks": sampler.summary(),
            "wall_s_main_leg": wall_main,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
            line["parity"] = parity
        if c4 is not None:
            line["c4_sharded"] = c4
        emit(json_fd, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
