#!/usr/bin/env python
"""Print the handful of ncu metrics this repo tracks from a .ncu-rep (run where ncu is installed).

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print(r[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                print(f"  {w:70s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
        if "--stalls" in sys.argv:
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") or \
                        h.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in h:
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v > 0.05:
                        print(f"  {h:70s} {v:16.3f}")


if __name__ == "__main__":
    main()
