"""Top stall sites from an `ncu --page source --csv` dump (SASS view).  usage: ncu_src.py file [section] [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(sys.argv) < 3 or sys.argv[2] == "list":
    for k, i in enumerate(secs):
        print(k, rows[i][1][:150])
    sys.exit()
k = int(sys.argv[2])
lo = secs[k]; hi_ = secs[k + 1] if k + 1 < len(secs) else len(rows)
print(rows[lo][1][:200])
hdr = rows[lo + 1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[lo + 2:hi_] if len(r) == len(hdr)]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
top = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:n]
for i in sorted(top):
    r = data[i]
    stalls = {h: int(r[ci[h]] or 0) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
    s = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {int(r[ci['# Samples']]):6d} {100*int(r[ci['# Samples']])/max(tot,1):5.1f}%  {r[ci['Source']][:80]:80s} {s}")
