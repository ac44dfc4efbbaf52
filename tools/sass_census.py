#!/usr/bin/env python
"""SASS opcode census of libmltcnn.so per kernel: the mnemonics that prove tcgen05 (UTCHMMA), TMEM loads (LDTM), TMA tensor
loads (UTMALDG), bulk copies (UBLKCP) and tcgen05 commits / mbarriers (UTCBAR, SYNCS) are what the product kernels execute.
    python tools/sass_census.py > profiles/r02/sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fastintercu_vvc_b200", "libmltcnn.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS", "ATOMS", "RED", "ATOM"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
cur, cnt = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if cur and m and m.group(1) in OPS:
        cnt[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a): instruction counts per kernel")
tot = collections.Counter()
for name, c in zip(names, cnt.values()):
    tot.update(c)
    if c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"] or c["LDTM"]:
        print(f"{name[:160]}\n    " + " ".join(f"{o}={c[o]}" for o in OPS if c[o]))
print(f"# {len(cnt)} kernels; kernels without tcgen05 / TMA instructions (head, staging probes, descriptor builders, SIMT cross-check engine) omitted above")
print("TOTAL " + " ".join(f"{o}={tot[o]}" for o in OPS))
