#!/usr/bin/env python
"""traffic.json from ncu reports (run where ncu is installed; profiles/capture.sh calls it on the box).

    python profiles/ncu_traffic.py --out gpurun_out/traffic_r2.json cfg5:16=prof_cfg5.ncu-rep T:1=prof_T1.ncu-rep ...

Per workload key ("name:frames per launch"): dram__bytes_read.sum + dram__bytes_write.sum summed over
the kernels of ONE step found in the report (bench.py copies it into roofline.traffic), the time
of each kernel, and the issue-side counters bench.py reports for workloads that are not HBM-bound
(thread-instructions per output pixel, FP64 pipe and issue-slot utilisation).
"""
import argparse
import csv
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        def get(name, default=None):
            if name not in hdr:
                return default
            i = hdr.index(name)
            try:
                return float(r[i].replace(",", "")) * UNIT.get(units[i], 1)
            except ValueError:
                return default
        yield r[hdr.index("Kernel Name")], get


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("items", nargs="+", help="key=report.ncu-rep")
    args = ap.parse_args()
    with open(os.path.join(REPO, "tests", "golden", "full_configs.json")) as fh:
        golden = json.load(fh)
    table = {"_comment": "written by profiles/ncu_traffic.py from the `ncu --set full` captures of profiles/capture.sh: "
                         "dram bytes, times and issue-side counters of the kernels of ONE step / call (ncu serialises grids "
                         "that overlap in a plain run). bench.py copies `bytes` into roofline.traffic."}
    for item in args.items:
        key, rep = item.split("=", 1)
        name, frames = key.split(":")
        px = golden["cfg4" if name == "cfg5" else name]["out_pixels"] * int(frames)
        kernels, read, write, inst, us = [], 0.0, 0.0, 0.0, 0.0
        fp64 = issue = 0.0
        for kname, get in rows_of(rep):
            t = get("gpu__time_duration.sum", 0.0)
            r, w = get("dram__bytes_read.sum", 0.0), get("dram__bytes_write.sum", 0.0)
            n = get("smsp__inst_executed.sum", 0.0)
            kernels.append({"kernel": kname.split("(")[0], "us": round(t, 2), "read": int(r), "write": int(w),
                            "registers": get("launch__registers_per_thread"),
                            "warp_inst": int(n),
                            "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
                            "dram_pct_of_peak": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
            read += r
            write += w
            inst += n
            us += t
            fp64 += t * (get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0.0) or 0.0)
            issue += t * (get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) or 0.0)
        if not kernels:
            continue
        table[key] = {"bytes": int(read + write), "read": int(read), "write": int(write), "us_serialised": round(us, 2),
                      "inst_per_px": round(inst * 32 / px, 1), "fp64_pipe_pct": round(fp64 / us, 1),
                      "issue_active_pct": round(issue / us, 1), "kernels": kernels,
                      "source": "profiles/" + re.sub(r"^prof_([^_]+)_", r"\1_ncu_full_", os.path.basename(rep)).replace(".ncu-rep", ".txt")}
    with open(args.out, "w") as fh:
        json.dump(table, fh, indent=1)
    print(json.dumps({k: (v if k == "_comment" else {kk: v[kk] for kk in ("bytes", "us_serialised", "inst_per_px")})
                      for k, v in table.items()}, indent=1))


if __name__ == "__main__":
    main()
