#!/usr/bin/env python
"""Summarise an `ncu --set full` report into one row per kernel class (mean over the captured launches).
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_full_summary.csv"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

COLS = OrderedDict([
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__waves_per_multiprocessor", "waves_per_sm"),
])
SCALE = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "second": 1e6,
         "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        a = agg.setdefault(name, {"n": 0})
        a["n"] += 1
        for m, short in COLS.items():
            if m not in idx:
                continue
            try:
                v = float(r[idx[m]].replace(",", ""))
            except ValueError:
                continue
            v *= SCALE.get(units[idx[m]], 1.0)
            a[short] = a.get(short, 0.0) + v
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches"] + list(COLS.values()))
    for name, a in agg.items():
        w.writerow([name, a["n"]] + ["%.4g" % (a.get(s, float("nan")) / a["n"]) for s in COLS.values()])


if __name__ == "__main__":
    main()
