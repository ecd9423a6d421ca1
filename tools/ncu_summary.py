#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i rep --page raw --csv`) into the short metric list kept under profiles/."""
import csv
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__", "sm__inst_executed_pipe",
        "sm__inst_issued", "sm__pipe_alu", "sm__throughput", "sm__warps_active", "smsp__average_warps_issue_stalled",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg", "smsp__thread_inst_executed_per_inst_executed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_shared", "sm__sass_inst_executed_op_shared")


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title}")
    for val in rows[2:]:
        name = dict(zip(hdr, val)).get("Kernel Name", "")
        print(f"# kernel: {name}")
        print("metric,unit,value")
        for h, u, v in zip(hdr, units, val):
            if any(h.startswith(k) for k in KEEP):
                print(f"{h},{u},{v}")


if __name__ == "__main__":
    main()
