"""Summarises Nsight Compute reports (ncu --set full ... -o X) into the JSON evidence kept under profiles/:
per kernel launch the duration, DRAM / L2 / shared-memory traffic, achieved bandwidths, occupancy and pipe utilisation, and --
with --traffic OUT -- the measured DRAM bytes per launch of the covariance downdate that bench.py reports as `roofline.traffic`
(so that number comes from a capture, not from a hand-edited file).

    python tools/ncu_summary.py gpurun_out/r02b/prof_frame_c3.ncu-rep [more.ncu-rep] > profiles/r02_kernels_ncu.json
    python tools/ncu_summary.py REP --traffic profiles/roofline_traffic.json --workload c3
Reads the report with `ncu -i REP --page raw --csv` (works without a GPU)."""
import csv
import io
import json
import subprocess
import sys

HBM_PEAK_GBS = 6542.1      # MEASURED_PEAKS.json (driver-written, this pool's B200)
WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "global_ld_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum": "global_st_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts.sum": "global_ldgsts_sectors",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_active.avg": "sm_cycles_active",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
              "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def read_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        if len(r) != len(header):
            continue
        name = r[header.index("Kernel Name")]
        rec = {"kernel": name[5:] if name.startswith("void ") else name, "id": int(r[header.index("ID")])}
        for i, h in enumerate(header):
            if h in WANT and r[i] != "":
                try:
                    val = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                rec[WANT[h]] = val * UNIT_SCALE.get(units[i], 1.0)
        res.append(rec)
    return res


def derive(rec):
    d = rec.get("duration_us")
    if d:
        dr = rec.get("dram_read_bytes", 0.0) + rec.get("dram_write_bytes", 0.0)
        rec["dram_bytes"] = dr
        rec["dram_gbs"] = dr / d / 1e3
        rec["dram_frac_of_measured_hbm_peak"] = rec["dram_gbs"] / HBM_PEAK_GBS
        sect = rec.get("global_ld_sectors", 0.0) + rec.get("global_st_sectors", 0.0) + rec.get("global_ldgsts_sectors", 0.0)
        rec["global_bytes_requested"] = 32.0 * sect        # 32-byte sectors the SMs asked L1/L2 for (loads, stores, cp.async)
        rec["global_gbs"] = rec["global_bytes_requested"] / d / 1e3
        if "smem_wavefronts" in rec:
            rec["smem_gbs"] = rec["smem_wavefronts"] * 128.0 / d / 1e3     # 128 bytes per shared-memory wavefront
    return rec


def main():
    args = sys.argv[1:]
    traffic_out, workload = None, "c3"
    if "--traffic" in args:
        i = args.index("--traffic"); traffic_out = args[i + 1]; del args[i:i + 2]
    if "--workload" in args:
        i = args.index("--workload"); workload = args[i + 1]; del args[i:i + 2]
    launches = []
    for p in args:
        for rec in read_report(p):
            rec["report"] = p
            launches.append(derive(rec))
    if traffic_out:
        dd = [r for r in launches if r["kernel"].startswith("k_downdate")]
        if not dd:
            raise SystemExit("no k_downdate launch in the reports")
        per_launch = sum(r["dram_bytes"] for r in dd) / len(dd)
        try:
            cur = json.load(open(traffic_out))
        except Exception:
            cur = {}
        cur[workload] = per_launch
        cur.pop("note", None)
        cur["_source"] = (f"tools/ncu_summary.py over {args}: mean of dram__bytes_read.sum + dram__bytes_write.sum of the "
                          f"{len(dd)} k_downdate launches of one frame ({[round(r['dram_bytes'] / 1e6, 1) for r in dd]} MB)")
        json.dump(cur, open(traffic_out, "w"), indent=1)
    # aggregate per kernel name
    agg = {}
    for r in launches:
        a = agg.setdefault(r["kernel"].split("(")[0], {"launches": 0, "duration_us": 0.0})
        a["launches"] += 1
        a["duration_us"] += r.get("duration_us", 0.0)
    print(json.dumps({"hbm_peak_gbs": HBM_PEAK_GBS, "per_kernel_total": agg, "launches": launches}, indent=1))


if __name__ == "__main__":
    main()
