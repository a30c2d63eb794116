"""Summarise `ncu --page source --csv` output: stall samples by opcode and the top instructions.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 | python profiles/stall_summary.py"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
si = hdr.index("Warp Stall Sampling (All Samples)")
ii = hdr.index("Instructions Executed")
src = hdr.index("Source")


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


data = [r for r in rows[hi + 1:] if len(r) > max(si, ii)]
tot = sum(num(r[si]) for r in data) or 1
print(rows[0][1] if rows[0] else "", "| samples", tot, "| SASS instructions", len(data),
      "| executed warp-instr %.1fM" % (sum(num(r[ii]) for r in data) / 1e6))
byop = defaultdict(lambda: [0, 0])
for r in data:
    toks = r[src].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    byop[op][0] += num(r[si])
    byop[op][1] += num(r[ii])
for op, (s, n) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:16]:
    print("%-10s stall %5.1f%%  executed %8.2fM" % (op, 100.0 * s / tot, n / 1e6))
print()
for r in sorted(data, key=lambda r: -num(r[si]))[:int(sys.argv[1]) if len(sys.argv) > 1 else 20]:
    print("%5.2f%%  %7.2fM  %s" % (100.0 * num(r[si]) / tot, num(r[ii]) / 1e6, r[src].strip()[:100]))
