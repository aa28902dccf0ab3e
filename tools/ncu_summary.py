#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the text form kept under profiles/:  tools/ncu_summary.py in.ncu-rep "title" > out.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__", "sm__cycles_elapsed.avg ", "sm__throughput.avg.pct",
        "sm__pipe_fmaheavy_cycles_active.avg.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "smsp__inst_executed.sum", "smsp__inst_executed.max",
        "smsp__inst_executed.min", "smsp__issue_active.avg.pct", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_per_inst_issued", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum ", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "Kernel Name", "Block Size",
        "Grid Size", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ")


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f"# {title}")
    print(f"# source: ncu --set full --clock-control none --import-source on ({rep.split('/')[-1]}), summarised by tools/ncu_summary.py")
    for h, u, v in zip(hdr, units, vals):
        if any((h + " ").startswith(k) or k in (h + " ") for k in KEEP):
            print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main()
