"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt [traffic.json key]
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


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
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
    if len(sys.argv) > 4:
        path, key = sys.argv[3], sys.argv[4]
        try:
            cur = json.load(open(path))
        except (OSError, ValueError):
            cur = {}
        vals = [v for vs in traffic.values() for v in vs]
        cur[key] = sum(vals) / len(vals)
        cur[key + "_source"] = out
        json.dump(cur, open(path, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
