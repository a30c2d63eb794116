"""Stall samples of one kernel by opcode and by site, from `ncu -i X.ncu-rep --page source --csv` (captured with --import-source on).
usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_stalls.py "<substring of the kernel name>" """
import collections
import csv
import sys

want = sys.argv[1]
cur = None
hdr = None
agg = collections.Counter()
tot = 0
top = []
for r in csv.reader(sys.stdin):
    if r and r[0] == "Kernel Name":
        cur = r[1].replace("fb::", "").replace("(int)", "").replace("(bool)", "")
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if cur and want in cur and hdr and len(r) > 4 and r[0].startswith("0x"):
        try:
            n = int(r[2])
        except ValueError:
            continue
        toks = r[1].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        agg[op] += n
        tot += n
        top.append((n, r[1].strip()[:70]))
print(want, "stall samples:", tot)
if tot:
    print("  by opcode:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in agg.most_common(10)))
    print("  top sites:", "; ".join("%.1f%% %s" % (100 * n / tot, s) for n, s in sorted(top, reverse=True)[:6]))
