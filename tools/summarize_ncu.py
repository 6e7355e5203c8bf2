#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i`, no GPU needed) into the small CSV kept under profiles/.
Usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/out.csv"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__cycles_elapsed.avg.per_second",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [h for h in hdr if h in KEEP or "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "metric", "unit", "value"])
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            for c in cols:
                w.writerow([name[:120], c, units[hdr.index(c)], r[hdr.index(c)]])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
