"""Instruction mix of a kernel from an .ncu-rep captured with --import-source on: executed warp
instructions per SASS opcode (and the hottest instructions), to see where an issue-bound kernel
spends its slots.

    python tools/ncu_opmix.py gpurun_out/x.ncu-rep [top]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc, iex, ist = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter()
stall = collections.Counter()
total = 0
hot = []
for r in rows[2:]:
    if len(r) <= iex:
        continue
    try:
        n = int(r[iex])
    except ValueError:
        continue
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "MUFU", "SHFL")) and "." in op else "")
    ops[op] += n
    stall[op] += int(r[ist] or 0)
    total += n
    hot.append((n, src))
print("total warp instructions executed: %d" % total)
for op, n in ops.most_common(top):
    print("  %-14s %12d  %5.1f %%   stall samples %d" % (op, n, 100.0 * n / total, stall[op]))
