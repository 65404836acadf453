"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

usage: launch_summary.py launches.csv [steps_in_capture] > profiles/rNN_launches_summary.txt
Times are cold-cache, serialised per-launch durations: compare SHARES, not absolutes.
"""
import csv
import collections
import re
import sys


def short(name: str) -> str:
    name = name.replace("void ", "")
    fn = name.split("<")[0].split("(")[0]
    if fn.startswith(("pcoe::", "v4::", "v5::", "v6::")):          # ncu drops the outer namespace of nested ones
        targs = name[len(fn):].split(">(")[0] if "<" in name else ""
        args = re.findall(r"(?:pcoe|v4|v5|v6)::(\w+)", targs)
        fn = fn if fn.startswith("pcoe::") else "pcoe::" + fn
        return f"{fn}<{','.join(args)}>" if args else fn
    return fn[:70]


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((short(r[ci["Kernel Name"]]), us))
    tot = sum(u for _, u in rows)
    agg = collections.OrderedDict()
    for k, u in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += u
    print(f"# {len(rows)} launches, {tot:.1f} us total over {steps} step(s): {tot / steps:.1f} us of kernel time per step")
    print(f"{'kernel':78s} {'n/step':>7s} {'us/step':>9s} {'share':>7s}")
    for k, (n, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:78s} {n / steps:7.1f} {u / steps:9.1f} {100 * u / tot:6.1f}%")
    ours = sum(u for k, (n, u) in agg.items() if k.startswith("pcoe::"))
    print(f"# libpcoe kernels: {ours / steps:.1f} us/step ({100 * ours / tot:.1f}%), torch/cuBLAS/NCCL kernels: {(tot - ours) / steps:.1f} us/step")


if __name__ == "__main__":
    main()
