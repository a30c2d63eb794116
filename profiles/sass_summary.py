"""Opcode histogram per kernel of libflutas_b200.so (cuobjdump -sass): which kernels move data with TMA (UTMALDG), which through
the LSU (LDG/STG), how much shared-memory and FP64 work each carries.  usage: python profiles/sass_summary.py [regex]"""
import collections
import re
import subprocess
import sys

so = "flutas_b200/csrc/libflutas_b200.so"
want = re.compile(sys.argv[1] if len(sys.argv) > 1 else ".")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name = None
hist = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("fb::", "").replace("(int)", "").replace("(bool)", "")
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        hist[name][m.group(1)] += 1
KEYS = ["UTMALDG", "UTMASTG", "SYNCS", "LDG", "STG", "LDGSTS", "LDS", "STS", "SHFL", "BAR", "WARPSYNC", "DFMA", "DADD", "DMUL", "MUFU", "LDL", "STL"]
print("%-86s %s" % ("kernel", " ".join("%7s" % k for k in KEYS)))
for k in sorted(hist):
    if want.search(k) and sum(hist[k].values()) > 50:
        print("%-86s %s" % (k[:86], " ".join("%7d" % hist[k][q] for q in KEYS)))
