"""Summarise an `ncu --page raw --csv` dump: stall reasons, pipes, DRAM bytes, duration."""
import csv
import re
import sys


def f(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main(path, row=2):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        print("==", d.get("Kernel Name", ("?",))[0], "grid", d.get("Grid Size", ("?",))[0], "block", d.get("Block Size", ("?",))[0])
        st = [(k, d[k][0]) for k in d if re.search(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k)]
        print("stalls (warps per issue):")
        for k, v in sorted(st, key=lambda kv: -f(kv[1]))[:8]:
            print("   %-28s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_alu.sum",
                "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum",
                "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]
        for k in keys:
            if k in d:
                print("   %-62s %s %s" % (k, d[k][0], d[k][1]))


if __name__ == "__main__":
    main(sys.argv[1])
