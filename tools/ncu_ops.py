"""Opcode mix (executed warp-instructions) from an `ncu --page source --csv` dump. usage: ncu_ops.py file section"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k = int(sys.argv[2])
lo = secs[k]; hi_ = secs[k + 1] if k + 1 < len(secs) else len(rows)
hdr = rows[lo + 1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[lo + 2:hi_] if len(r) == len(hdr)]
ops = collections.Counter(); samp = collections.Counter()
tot = 0
for r in data:
    src = r[ci["Source"]].split()
    op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
    op = op.split(".")[0]
    n = int(r[ci["Instructions Executed"]] or 0)
    ops[op] += n; tot += n
    samp[op] += int(r[ci["# Samples"]] or 0)
print(rows[lo][1][:120]); print("total warp-instr", tot, "static", len(data))
for op, n in ops.most_common(28):
    print(f"  {op:12s} {n:10d} {100*n/tot:5.1f}%   samples {samp[op]}")
