#!/usr/bin/env python3
"""tools/pipes_table.py <raw.csv> [log2n] -- per-kernel dynamic instruction mix from `ncu -i REP --page raw --csv`
(sections ComputeWorkloadAnalysis InstructionStats SchedulerStats WarpStateStats): warp instructions per element-row,
issue-slot and pipe utilisation, stall reasons per issue."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return float('nan')
print("%-10s %7s %6s %5s | %5s %5s %5s %5s %5s | %5s %5s %5s %5s %5s %4s" % ("kernel", "us", "inst/e", "issue", "alu%", "fma%", "fp64%", "xu%", "lsu%", "warps", "math", "nosel", "short", "long", "regs"))
for r in rows[2:]:
    m = re.search(r'k_map_flat<(?:nuk::)?(\w+)', r[ix['Kernel Name']])
    nm = m.group(1) if m else r[ix['Kernel Name']][:20]
    t = g(r, 'gpu__time_duration.sum')
    print("%-10s %7.1f %6.1f %5.1f | %5.1f %5.1f %5.1f %5.1f %5.1f | %5.1f %5.2f %5.2f %5.2f %5.2f %4d" % (
        nm, t if t > 5 else t * 1000, g(r, 'smsp__inst_executed.sum') * 32 / n, g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        g(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'), g(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
        g(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'), g(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
        g(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'), g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        g(r, 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio'), g(r, 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio'),
        g(r, 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'), g(r, 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'), g(r, 'launch__registers_per_thread')))
