#!/usr/bin/env python
"""Summarise an ncu report: per-kernel raw metrics of interest and the executed-instruction mix per sample.
usage: scripts/ncu_summary.py gpurun_out/x.ncu-rep [samples_per_launch]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
per_launch = float(sys.argv[2]) if len(sys.argv) > 2 else 1e9
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:95s} {units[i]:12s} {[r[i] for r in data]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
start, end = hi[0], (hi[1] - 1 if len(hi) > 1 else len(rows))
h = rows[start]; ci = {x: i for i, x in enumerate(h)}
body = [r for r in rows[start + 1:end] if len(r) == len(h)]
c = Counter(); st = Counter()
for r in body:
    toks = r[ci['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    c[op] += int(r[ci['Instructions Executed']]); st[op] += int(r[ci['# Samples']])
ws = per_launch / 32
print("\nexecuted warp-instructions per warp of 32 samples (= instructions per sample)")
for op, v in c.most_common(28):
    print(f"  {op:10s} {v / ws:7.2f}   stall-samples {st[op]}")
print(f"  total      {sum(c.values()) / ws:7.2f}")
stall_cols = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]
tot = Counter()
for r in body:
    for x in stall_cols:
        tot[x] += int(r[ci[x]] or 0)
print("\nwarp stall samples (all):", {k: v for k, v in tot.most_common(8)})
if len(sys.argv) > 3:
    print("\nhottest SASS lines")
    for r in sorted(body, key=lambda r: -int(r[ci['# Samples']]))[:40]:
        print(f"  {r[ci['# Samples']]:>7s} {r[ci['Instructions Executed']]:>12s}  {r[ci['Source']]}")
