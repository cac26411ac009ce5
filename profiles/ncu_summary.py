#!/usr/bin/env python3
"""Extract the metrics we quote (bench.py roofline.traffic) from an .ncu-rep into a small CSV.
Usage: python profiles/ncu_summary.py report.ncu-rep > profiles/rN_ncu_<kernel>.csv"""
import csv
import subprocess
import sys

KEEP = ("Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__pipe_alu_cycles_active.avg.pct", "sm__pipe_fp64_cycles_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct", "smsp__issue_active.avg.pct",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct", "sm__cycles_elapsed.avg ",
        "smsp__sass_thread_inst_executed_op_dfma", "smsp__sass_thread_inst_executed_op_integer", "sm__inst_executed_pipe_uniform",
        "smsp__thread_inst_executed_per_inst_executed.ratio")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    w.writerow(["launch", "metric", "unit", "value"])
    for li, vals in enumerate(rows[2:]):
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) or k in h for k in KEEP) and v != "":
                w.writerow([li, h, u, v])


if __name__ == "__main__":
    main(sys.argv[1])
