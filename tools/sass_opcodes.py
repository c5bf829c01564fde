"""Opcode census of the shipped library (cuobjdump -sass): python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.md
DMMA = FP64 tensor-core mma (mma.sync.m8n8k4.f64), UBLKCP = cp.async.bulk (the bulk-copy / TMA engine),
SYNCS = mbarrier operations, FENCE.VIEW.ASYNC = fence.proxy.async / fence.mbarrier_init, LDL/STL = spills."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cocons_b200", "libcocons_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["DMMA", "UBLKCP", "SYNCS", "FENCE.VIEW.ASYNC", "UTMALDG", "HMMA", "LDGSTS", "LDS", "LDG", "STG", "DFMA", "DMUL",
        "DADD", "MUFU", "BAR.SYNC", "ATOM", "RED", "MEMBAR", "LDL", "STL"]
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ops, total = collections.Counter(), 0
    for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f):
        total += 1
        for w in want:
            if m.group(1).startswith(w):
                ops[w] += 1
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    rows.append((re.sub(r"\(.*\)$", "", dem), total, ops))
print("# SASS opcode census of `%s` (sm_100a)\n" % os.path.relpath(so, ROOT))
print(__doc__.split("\n", 1)[1])
print("| kernel | SASS instr | " + " | ".join(want) + " |")
print("|---|---:|" + "---:|" * len(want))
for name, total, ops in sorted(rows, key=lambda r: -r[1]):
    print("| `%s` | %d | " % (name[:80], total) + " | ".join(str(ops.get(w, 0) or "") for w in want) + " |")
