"""Summarise `ncu -i rep --page source --csv`: stall-reason totals, hottest SASS lines, opcode histogram.

  ncu -i gpurun_out/x.ncu-rep --page source --csv > /tmp/src.csv
  python profiles/ncu_source_summary.py /tmp/src.csv [top_n]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = collections.Counter()
samples = 0
for r in data:
    samples += int(r[ci["# Samples"]] or 0)
    for s in stall_cols:
        tot[s] += int(r[ci[s]] or 0)
print(f"kernel: {rows[0][1][:100]}\nsamples {samples}")
print("stall totals:", ", ".join(f"{k[6:]} {v} ({100 * v / max(1, samples):.0f}%)" for k, v in tot.most_common(10)))
ops = collections.Counter()
ops_s = collections.Counter()
for r in data:
    op = r[ci["Source"]].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
    op = op.split(".")[0] if not op.startswith("MUFU") else op
    ops[op] += int(r[ci["Instructions Executed"]] or 0)
    ops_s[op] += int(r[ci["# Samples"]] or 0)
n = sum(ops.values())
print("opcode histogram (warp instructions, share; samples):")
for k, v in ops.most_common(24):
    print(f"  {k:14s} {v:10d} {100 * v / n:5.1f}%   samples {ops_s[k]:7d} {100 * ops_s[k] / max(1, samples):5.1f}%")
print(f"hottest {topn} lines by samples:")
idx = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:topn]
for i in idx:
    r = data[i]
    why = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:2]
    print(f"  #{i:5d} {int(r[ci['# Samples']]):7d}  {r[ci['Source']].strip()[:70]:70s} {why}")
