#!/usr/bin/env python
"""Per-kernel and per-stage DRAM traffic of one bench step from an `ncu --set full` report of
benchmarks/profile_step.py (1 step profiled):

    python benchmarks/ncu_dram_bytes.py gpurun_out/r02_step.ncu-rep [--out profiles/r02]

writes <out>_dram_bytes.json  (what bench.py's `roofline.traffic` / `stages[].traffic` read: dram__bytes_read.sum +
dram__bytes_write.sum per step, summed over the launches of a stage) and <out>_step_ncu_summary.txt (one block per
launch: duration, DRAM bytes and achieved GB/s against MEASURED_PEAKS.json, issue-active, warp-instructions, occupancy
limiter, pipes, top stall reasons).  Launches are assigned to the stages of pipeline.FusedSplatStep by their order."""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum"]


def to_bytes(val, unit):
    v = float(val.replace(",", "")) if val else 0.0
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def to_us(val, unit):
    v = float(val.replace(",", "")) if val else 0.0
    return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(unit, 1.0)


def stage_of(name, state):
    n = name
    if "project_fwd_kernel" in n:
        state["s"] = "isect_prepare"
        return "project_fwd"
    if "emit_boundaries_kernel" in n:
        state["s"] = "isect_fill"
    if "raster_fwd" in n:
        state["s"] = "after_fwd"
        return "raster_fwd"
    if "raster_bwd" in n:
        return "raster_bwd"
    if "project_bwd_kernel" in n:
        return "project_bwd"
    if "loss_" in n or "ssim_" in n:
        return "loss"
    if any(k in n for k in ("scan_", "radix_", "emit_", "compose_ids")):
        return state["s"] if state["s"] in ("isect_prepare", "isect_fill") else "isect_other"
    if "adam_arena" in n:
        return "adam"
    if "strategy_update" in n:
        return "strategy_update"
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02"))
    ap.add_argument("--workload", default=None)
    ap.add_argument("--what", default="ONE fused step (benchmarks/profile_step.py)")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    state = {"s": "isect_prepare"}
    stages, kernels, lines = {}, {}, []
    for r in rows:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("qed::", "")
        st = stage_of(name, state)
        us = to_us(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]])
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us else 0.0
        S = stages.setdefault(st, {"dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0, "duration_us": 0.0, "launches": 0})
        S["dram_bytes"] += rd + wr
        S["dram_read"] += rd
        S["dram_write"] += wr
        S["duration_us"] += us
        S["launches"] += 1
        Kk = kernels.setdefault(short, {"dram_bytes": 0.0, "duration_us": 0.0, "launches": 0})
        Kk["dram_bytes"] += rd + wr
        Kk["duration_us"] += us
        Kk["launches"] += 1
        lines.append(f"===== [{st}] {short}")
        lines.append(f"{'duration':60s} {us:10.2f} us    DRAM read {rd / 1e6:8.2f} MB  write {wr / 1e6:8.2f} MB  -> {gbs:7.0f} GB/s = {gbs / peak:.3f} of {peak:.0f} (measured)")
        for w in WANT[3:]:
            if w in idx:
                lines.append(f"{w:60s} {r[idx[w]]} {units[idx[w]]}")
        top = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:6]
        for v, h in top:
            lines.append(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.2f}")
    for S in stages.values():
        S["dram_gbs_over_stage_kernel_time"] = S["dram_bytes"] / (S["duration_us"] * 1e-6) / 1e9 if S["duration_us"] else None
    total_us = sum(S["duration_us"] for S in stages.values())
    head = [f"# ncu --set full --clock-control none of {a.what}, {len(rows)} launches, {total_us:.1f} us of kernel time",
            "# (cold-cache, serialised replays: compare SHARES and bytes, not absolute times); HBM peak = MEASURED_PEAKS.json",
            "# stage                launches   kernel us   share    DRAM MB (read + write)      GB/s over kernel time"]
    for st, S in sorted(stages.items(), key=lambda kv: -kv[1]["duration_us"]):
        head.append(f"# {st:20s} {S['launches']:5d} {S['duration_us']:11.1f} {100 * S['duration_us'] / total_us:6.1f} %  {S['dram_bytes'] / 1e6:9.2f} "
                    f"({S['dram_read'] / 1e6:.2f} + {S['dram_write'] / 1e6:.2f})   {S['dram_gbs_over_stage_kernel_time']:.0f}")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out + "_step_ncu_summary.txt", "w").write("\n".join(head + lines) + "\n")
    sys.path.insert(0, ROOT)
    workload = a.workload
    if workload is None:
        import bench

        workload = bench.workload_name(argparse.Namespace(gaussians=1_000_000, width=1920, height=1080, mode="RGB+ED"))
    json.dump({"source": os.path.basename(a.report), "how": "ncu --set full --clock-control none --profile-from-start off python benchmarks/profile_step.py",
               "workload": workload, "hbm_peak_gbs": peak, "stages": stages, "kernels": kernels}, open(a.out + "_dram_bytes.json", "w"), indent=1)
    print("\n".join(head))


if __name__ == "__main__":
    main()
