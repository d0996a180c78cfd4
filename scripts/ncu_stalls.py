"""Aggregate the warp-stall samples of the first kernel of an .ncu-rep by opcode and by stall reason.
usage: python scripts/ncu_stalls.py x.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; idx = {x: i for i, x in enumerate(h)}
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
by_op, by_stall, inst = {}, {}, {}
tot = 0
for r in rows[2:]:
    if len(r) < len(h) or r[0] == "Kernel Name":
        break
    srcl = r[idx["Source"]].strip()
    if not srcl:
        continue
    parts = srcl.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    n = int(r[idx["# Samples"]] or 0)
    tot += n
    by_op[op] = by_op.get(op, 0) + n
    inst[op] = inst.get(op, 0) + int(r[idx["Instructions Executed"]] or 0)
    for c in stall_cols:
        by_stall[(op, c)] = by_stall.get((op, c), 0) + int(r[idx[c]] or 0)
print("total samples", tot)
for op, n in sorted(by_op.items(), key=lambda kv: -kv[1])[:14]:
    st = sorted(((c, v) for (o, c), v in by_stall.items() if o == op), key=lambda kv: -kv[1])[:3]
    print(f"{op:8s} samples {n:9d} ({n/tot:5.1%})  executed {inst[op]:12d}  " + ", ".join(f"{c[6:]}={v/max(n,1):.0%}" for c, v in st))
print("all instructions executed:", sum(inst.values()))
