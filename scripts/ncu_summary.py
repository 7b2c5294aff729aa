"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt [profiles/kernel_counters.json key kernel_substring units]

With the optional arguments the executed FP64 thread-instruction counts, DRAM bytes and FP64 pipe utilisation of the first
kernel whose name contains `kernel_substring` are stored under `key` (bench.py reads them for roofline.executed_frac /
roofline.traffic); `units` = stage-iterations that launch processed (horizon x problems).
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]
STALL = "smsp__average_warps_issue_stalled_"


def units_row(hdr, units, key):
    return units[hdr.index(key)]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    units_all = units
    lines = [f"source: {rep} (ncu --set full --clock-control none; per-launch values, kernel replayed in isolation)"]
    traffic = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("")
        lines.append(f"kernel {d['Kernel Name']}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:82s} {d[k]:>18s} {units[hdr.index(k)]}")
        stalls = sorted(((float(d[k]), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k in hdr
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k]), reverse=True)
        lines.append("  stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]))
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            rd = float(d["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(d["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]]
            lines.append(f"  dram traffic per launch (read + write): {(rd + wr) / 1e9:.3f} GB")
            traffic.setdefault(d["Kernel Name"].split("(")[0], []).append(rd + wr)
        except (KeyError, ValueError):
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 6:
        path, key, sub, units = sys.argv[3], sys.argv[4], sys.argv[5], float(sys.argv[6])
        try:
            cur = json.load(open(path))
        except (OSError, ValueError):
            cur = {}
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            if sub not in d["Kernel Name"]:
                continue
            cyc = float(d["sm__cycles_elapsed.max"])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            rd = float(d["dram__bytes_read.sum"]) * scale[units_row(hdr, units_all, "dram__bytes_read.sum")]
            wr = float(d["dram__bytes_write.sum"]) * scale[units_row(hdr, units_all, "dram__bytes_write.sum")]
            dur = float(d["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[units_row(hdr, units_all, "gpu__time_duration.sum")]
            cur[key] = {"dfma": float(d["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"]) * cyc,
                        "dmul": float(d["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"]) * cyc,
                        "dadd": float(d["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"]) * cyc,
                        "units": units, "dram_bytes_per_launch": rd + wr,
                        "fp64_pipe_active_frac": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]) / 100.0,
                        "launch_ms_ncu": dur,
                        "source": f"{out} ({d['Kernel Name'].split('(')[0]}, one launch of {units:.0f} stage-iterations: "
                                  "smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum.per_cycle_elapsed x sm__cycles_elapsed.max; "
                                  "dram__bytes_read.sum + dram__bytes_write.sum)"}
            break
        json.dump(cur, open(path, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
