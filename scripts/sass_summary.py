#!/usr/bin/env python
"""SASS evidence of the Blackwell path: per kernel of libb200dm.so, how many tcgen05 / TMEM / TMA / cluster
instructions it contains (cuobjdump -sass).  Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (bf16/f16),
LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit -> mbarrier, UTCATOMSWS = TMEM
alloc / dealloc, SYNCS = mbarrier ops, UCGABAR = cluster barrier, STAS = st.async into a peer CTA's shared memory,
ACQBULK = griddepcontrol.wait (programmatic dependent launch), HMMA / LDSM = mma.sync / ldmatrix (the attention kernels),
MUFU.TANH = tanh.approx, REDG = red.global (split-K weight gradients).

  python scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lightning-generative-models_b200", "b200dm", "libb200dm.so")
PAT = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "STAS", "ACQBULK", "HMMA",
       "LDSM", "MUFU.TANH", "REDG"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, counts, ninstr = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        ninstr[cur] = 0
        continue
    if cur is None or "/*" not in line:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m:
        continue
    op = m.group(1)
    ninstr[cur] += 1
    for p in PAT:
        if op.startswith(p):
            counts[cur][p] += 1
total = collections.Counter()
print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels (sm_100a); columns = instruction counts in the SASS")
print("kernel".ljust(72), "instr", " ".join(p.rjust(9) for p in PAT))
for k, c in counts.items():
    name = re.sub(r"\(.*", "", demangle(k)).replace("b200dm::", "").replace("void ", "")
    total.update(c)
    if sum(c.values()) == 0:
        continue
    print(name[:72].ljust(72), str(ninstr[k]).rjust(5), " ".join(str(c.get(p, 0)).rjust(9) for p in PAT))
print("TOTAL".ljust(72), str(sum(ninstr.values())).rjust(5), " ".join(str(total.get(p, 0)).rjust(9) for p in PAT))
