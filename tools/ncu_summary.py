#!/usr/bin/env python
"""Condenses an `ncu --set full` report (or its `--page raw --csv` export) into the handful of metrics the
roofline discussion uses.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe"),
]


def main():
    path = sys.argv[1]
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        text = open(path).read()
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{path}`\n")
    for n, r in enumerate(rows[2:]):
        print(f"## launch {n}: `{r[col['Kernel Name']]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in KEYS:
            if key in col:
                print(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
        print()


if __name__ == "__main__":
    main()
