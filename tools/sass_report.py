"""Resource usage and SASS mnemonic counts of every kernel in libfitnoise_b200.so (cuobjdump).

    python tools/sass_report.py > profiles/rNN_sass_and_resources.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fitnoise_b200", "_lib", "libfitnoise_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(.*", "", o).replace("(int)", "") for o in out]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            f = dict(re.findall(r"(\w+):(\d+)", line))
            usage[cur] = (int(f.get("REG", 0)), int(f.get("STACK", 0)), int(f.get("SHARED", 0)), int(f.get("LOCAL", 0)))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    keys = ("UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "LDS", "STS", "SHFL", "MUFU", "HMMA", "UTCHMMA")
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1).split(".")[0]
                if op in keys:
                    counts[cur][op] += 1
                counts[cur]["_all"] += 1
    names = list(counts)
    pretty = dict(zip(names, demangle(names)))
    print("Resource usage and SASS mnemonic counts of libfitnoise_b200.so (nvcc 12.9, -O3 -gencode arch=compute_100a,code=sm_100a")
    print("-lineinfo).  UBLKCP = cp.async.bulk global->shared (the TMA engine); SYNCS = mbarrier; DFMA/DMUL/DADD = fp64 pipe;")
    print("no HMMA/UTC*MMA anywhere: fp64, HBM-bound work -- tensor cores are not used by design.")
    print()
    print("%-46s %5s %6s %7s %7s | %6s %5s %5s %5s %5s %5s %5s %5s %5s" % (
        "kernel", "regs", "stack", "smem", "instr", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "LDS", "STS", "SHFL", "MUFU"))
    for n in sorted(names, key=lambda k: pretty[k]):
        r = usage.get(n, (0, 0, 0, 0))
        c = counts[n]
        assert c["HMMA"] == 0 and c["UTCHMMA"] == 0
        print("%-46s %5d %6d %7d %7d | %6d %5d %5d %5d %5d %5d %5d %5d %5d" % (
            pretty[n][:46], r[0], r[1], r[2], c["_all"], c["UBLKCP"], c["SYNCS"], c["DFMA"], c["DMUL"], c["DADD"],
            c["LDS"], c["STS"], c["SHFL"], c["MUFU"]))


if __name__ == "__main__":
    main()
