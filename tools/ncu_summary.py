#!/usr/bin/env python
"""Summarise ncu outputs (run where ncu is installed, no GPU needed).
  launches: tools/ncu_summary.py launches <launches.csv>
  full    : tools/ncu_summary.py full <report.ncu-rep> [out.csv]
  source  : tools/ncu_summary.py source <report.ncu-rep> [N]
"""
import collections
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum']


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) > vi:
            agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} n={len(v):4d} mean={sum(v)/len(v)/1e3:8.2f} us min={min(v)/1e3:8.2f} max={max(v)/1e3:8.2f} share={100*sum(v)/tot:5.1f}%")


def full(rep, out=None):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [['metric', 'unit'] + [f'launch{i}' for i in range(len(data))]]
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            lines.append([k, units[i]] + [r[i] for r in data])
    for ln in lines:
        print(",".join(ln))
    if out:
        with open(out, "w") as f:
            csv.writer(f).writerows(lines)


def source(rep, n=30):
    rows = ncu_csv(rep, "source")
    ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    seg = rows[ends[0] + 2: ends[1] if len(ends) > 1 else len(rows)]
    tot = sum(int(r[4]) for r in seg if r[4].isdigit())
    print("total samples", tot)
    idx = sorted(range(len(seg)), key=lambda i: -int(seg[i][4]) if seg[i][4].isdigit() else 0)[:n]
    for i in idx:
        r = seg[i]
        prev = seg[i - 1][1].strip()[:50] if i else ""
        print(f"{r[4]:>5s} {100*int(r[4])/max(tot,1):5.1f}% exec={r[5]:>7s}  {r[1].strip()[:70]:70s} | prev: {prev}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "full":
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
