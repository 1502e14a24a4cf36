"""Opcode evidence for the shipped library: per kernel, the count of tensor-core / TMA / TMEM SASS mnemonics.

    python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt     (cuobjdump -sass of awesome_b200/csrc/libawb.so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "awesome_b200", "csrc", "libawb.so")
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "SYNCS", "LDGSTS", "HMMA", "FFMA", "FSET", "MUFU",
       "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "ACQBULK", "UTCATOMSWS"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        per[cur][m.group(1)] += 1
        per[cur]["_total"] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a); mnemonic counts")
print("# library total: " + "  ".join(f"{k}={tot[k]}" for k in KEY if tot[k]))
print(f"{'kernel':70s} {'instr':>7s}  " + " ".join(f"{k:>8s}" for k in KEY[:12]))
for name, c in per.items():
    if c["_total"] < 64:
        continue
    print(f"{name[:70]:70s} {c['_total']:7d}  " + " ".join(f"{c[k]:8d}" for k in KEY[:12]))
