#!/usr/bin/env python
"""Summarise an ncu --set full report (.ncu-rep) into the handful of counters DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg.per_second",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"ncu --set full --clock-control none: {path.split('/')[-1]}")
    for r in rows[2:]:
        print("---")
        print("Kernel Name".ljust(100), r[idx["Kernel Name"]][:120])
        for w in WANT:
            if w in idx:
                print(w.ljust(100), r[idx[w]].rjust(18), units[idx[w]])


if __name__ == "__main__":
    main(sys.argv[1])
