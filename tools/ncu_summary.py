#!/usr/bin/env python
"""Key metrics of an ncu report (one block per profiled launch), as markdown.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_<kernel>.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary: {path}\n")
    for r in rows[2:]:
        print(f"## `{r[idx['Kernel Name']][:100]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for key, label in WANT:
            if key in idx:
                print(f"| {label} (`{key}`) | {r[idx[key]]} | {units[idx[key]]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
