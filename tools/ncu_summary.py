#!/usr/bin/env python3
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the numbers DESIGN.md
and profiles/ quote: duration, DRAM bytes, pipe utilisation, issue rate, stall reasons."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    col = {h: i for i, h in enumerate(hdr)}
    print("kernel:", vals[col.get("Kernel Name", 4)][:100])
    for k in KEYS:
        if k in col:
            print(f"{k:78s} {vals[col[k]]:>16s} {units[col[k]]}")
    print("-- warp stall reasons (average warps per issue-active cycle) --")
    st = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            st.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    for v, name in sorted(st, reverse=True)[:10]:
        print(f"  {name:28s} {v:8.3f}")


if __name__ == "__main__":
    main()
