#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV + markdown for profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(rep, out_prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out_prefix + ".csv", "w") as fh:
        w = csv.writer(fh)
        w.writerow(["launch", "kernel", "metric", "unit", "value"])
        for n, r in enumerate(rows[2:]):
            kern = r[col["Kernel Name"]]
            for k in KEYS:
                if k in col:
                    w.writerow([n, kern[:60], k, units[col[k]], r[col[k]]])
            for h, i in col.items():
                if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                    w.writerow([n, kern[:60], h.replace("smsp__average_warps_issue_stalled_", "stall_"), units[i], r[i]])
    print("wrote", out_prefix + ".csv", "launches:", len(rows) - 2)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
