"""Summarise an ncu report (--set full --import-source on, built with -lineinfo): headline metrics of the first
kernel and its warp instructions / lane utilisation / stall samples per CUDA source line.

    python tools/ncu_lines.py gpurun_out/march.ncu-rep [min_percent]
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
           "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def page(rep, *args):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout)))


def main():
    rep = sys.argv[1]
    floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    raw = page(rep, "--page", "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    print("| metric | value | unit |\n|---|---|---|")
    for m in METRICS:
        if m in hdr:
            print("| %s | %s | %s |" % (m, vals[hdr.index(m)], units[hdr.index(m)]))
    rows = page(rep, "--page", "source", "--print-source", "cuda,sass")
    cur, hdr, line, agg = None, None, None, {}
    for r in rows:
        if r and r[0] in ("File Name", "File Path"):
            cur, hdr = r[1].split("/")[-1], None
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) >= 12:
            if r[0] != "":
                line = (cur, int(r[0]), r[1].strip()[:96])
            else:
                try:
                    a = agg.setdefault(line, [0, 0, 0])
                    a[0] += int(r[7]); a[1] += int(r[8]); a[2] += int(r[6])
                except ValueError:
                    pass
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print("\nwarp instructions %d, stall samples %d; lines with >= %.1f %% of either:" % (tot, tots, floor))
    for k, a in sorted(agg.items(), key=lambda kv: (str(kv[0][0]), kv[0][1])):
        if a[0] * 100 >= tot * floor or a[2] * 100 >= tots * floor:
            print("%-14s %4d  inst %5.1f%%  lanes %4.1f  stalls %5.1f%%  %s" % (k[0], k[1], 100 * a[0] / tot, a[1] / max(a[0], 1), 100 * a[2] / tots, k[2]))


if __name__ == "__main__":
    main()
