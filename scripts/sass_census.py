#!/usr/bin/env python
"""SASS mnemonic census of the built library (no GPU needed): which kernels issue FP64 tensor-core MMAs (DMMA),
TMA bulk copies (UBLKCP) with mbarrier waits (SYNCS), cp.async (LDGSTS) and global reductions (REDG).
    cuobjdump -sass forte2_b200/libforte2_b200.so | python scripts/sass_census.py > profiles/<tag>_sass_census.md"""
import collections
import re
import subprocess
import sys

KEYS = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "REDG", "ATOMG", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    cur, counts = None, collections.OrderedDict()
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            counts[cur][m.group(1).split(".")[0]] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
    fam = collections.OrderedDict()
    for mangled, name in zip(counts, names):
        short = re.sub(r"\(.*", "", name).replace("void f2::", "").replace("f2::", "").replace("void ", "")
        fam.setdefault(short.split("<")[0], []).append(counts[mangled])
    print("# SASS mnemonic census of forte2_b200/libforte2_b200.so (sm_100a)\n")
    print("Static instruction counts per kernel (range over the template instances), from `cuobjdump -sass`;")
    print("DMMA = FP64 tensor-core MMA, UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier operations,")
    print("LDGSTS = cp.async, REDG = red.global.add.f64.\n")
    print("| kernel (instances) | SASS instructions | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for name, lst in fam.items():
        tot = [sum(c.values()) for c in lst]
        cells = []
        for k in KEYS:
            lo, hi = min(c.get(k, 0) for c in lst), max(c.get(k, 0) for c in lst)
            cells.append(str(lo) if lo == hi else f"{lo}–{hi}")
        span = str(tot[0]) if min(tot) == max(tot) else f"{min(tot)}–{max(tot)}"
        print(f"| `{name}` ({len(lst)}) | {span} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
