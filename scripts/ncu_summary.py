#!/usr/bin/env python
"""Summarise an .ncu-rep into the two text files kept under profiles/:
   <out>_keymetrics.csv  -- the handful of metrics DESIGN.md / bench.py quote
   <out>_details.csv     -- `ncu --page details --csv` of the same report
usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    with open(out + "_keymetrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "value", "unit"])
        for r in rows[2:]:
            d = dict(zip(h, r))
            for k in KEYS:
                if k in d:
                    w.writerow([d["Kernel Name"].split("(")[0], k, d[k], u[h.index(k)]])
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    open(out + "_details.csv", "w").write(det)


if __name__ == "__main__":
    main()
