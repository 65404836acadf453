"""Summarise one step's `ncu --set full` capture of the libpcoe GEMM / grouping kernels.

    ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/r01_ncu_full_summary.txt profiles/traffic.json

The capture is taken with  -k regex:'tc4_|tc5_|knn_kernel'  over exactly one eager training step of
bench.py (config c2, bf16), so the launches arrive in a fixed order; this tool attaches the profile names
bench.py uses (sa1_fwd_l1 ...) by that order, prints the table the roofline discussion in DESIGN.md cites
and writes traffic.json (kernel name -> dram__bytes_read.sum + dram__bytes_write.sum per launch), which
bench.py reports as `roofline.traffic`.
"""
import csv
import json
import sys

ORDER_X3 = (["knn_kernel(sa1)", "sa1_fwd_l1", "sa1_fwd_l2", "sa1_fwd_l3", "knn_kernel(sa2)", "sa2_fwd_l1", "sa2_fwd_l2", "sa2_fwd_l3",
             "sa3_fwd_l1", "sa3_fwd_l2", "sa3_fwd_l3"] +
            [f"sa{l}_bwd_{k}{i}" for l in (3, 2) for i in (3, 2, 1) for k in ("wgrad", "dgrad")] +
            ["sa1_bwd_wgrad3", "sa1_bwd_dgrad3", "sa1_bwd_wgrad2", "sa1_bwd_dgrad2"])   # bf16x3: -k regex:'x3_|knn_k32' (dW1 of SA1 comes
                                                                                        # out of sa1_bwd_dgrad2's epilogue: no wgrad1 launch)
ORDER = (["knn_kernel(sa1)", "sa1_fwd_l1", "sa1_fwd_l2", "sa1_fwd_l3", "knn_kernel(sa2)", "sa2_fwd_l1", "sa2_fwd_l2",
          "sa2_fwd_l3", "sa3_fwd_l1", "sa3_fwd_l2", "sa3_fwd_l3", "sa3_bwd_wgrad3", "sa3_bwd_dgrad3", "sa3_bwd_wgrad2",
          "sa3_bwd_dgrad2", "sa3_bwd_wgrad1", "sa3_bwd_dgrad1", "sa2_bwd_l3", "sa2_bwd_l2", "sa2_bwd_l1", "sa1_bwd_l3",
          "sa1_bwd_l2"])   # sa1_bwd_l1 no longer runs: dW1 is accumulated by sa1_bwd_l2's epilogue (MaskStatsW1)
COLS = [("gpu__time_duration.sum", "us", 1.0),
        ("dram__bytes_read.sum", "rd MB", 1.0),
        ("dram__bytes_write.sum", "wr MB", 1.0),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1.0),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0)]


def to_bytes(v: float, unit: str) -> float:
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main():
    raw, out_txt, out_json = sys.argv[1:4]
    mode = sys.argv[4] if len(sys.argv) > 4 else "bf16"
    global ORDER
    if mode == "bf16x3":
        ORDER = ORDER_X3
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    if len(data) != len(ORDER):
        print(f"warning: {len(data)} launches captured, expected {len(ORDER)}", file=sys.stderr)
    lines, traffic = [], {}
    lines.append(f"{'kernel':18s} " + " ".join(f"{n:>9s}" for _, n, _ in COLS) + "   ncu kernel name")
    for k, r in enumerate(data):
        name = ORDER[k] if k < len(ORDER) else f"launch{k}"
        vals = []
        for col, _, _ in COLS:
            if col not in ci:
                vals.append(float("nan")); continue
            v = float(r[ci[col]].replace(",", "") or "nan")
            u = units[ci[col]]
            if col.startswith("dram__bytes"):
                v = to_bytes(v, u) / 1e6
            if col == "gpu__time_duration.sum":
                v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
            vals.append(v)
        rd = to_bytes(float(r[ci["dram__bytes_read.sum"]].replace(",", "")), units[ci["dram__bytes_read.sum"]])
        wr = to_bytes(float(r[ci["dram__bytes_write.sum"]].replace(",", "")), units[ci["dram__bytes_write.sum"]])
        traffic[name] = rd + wr
        lines.append(f"{name:18s} " + " ".join(f"{v:9.2f}" for v in vals) + "   " + r[ci["Kernel Name"]][:70])
    open(out_txt, "w").write("\n".join(lines) + "\n")
    # traffic.json is keyed by precision mode: {"bf16": {kernel: bytes}, "bf16x3": {...}}
    try:
        allt = json.load(open(out_json))
        if allt and not isinstance(next(iter(allt.values())), dict):
            allt = {"bf16": allt}
    except Exception:
        allt = {}
    allt[mode] = traffic
    json.dump(allt, open(out_json, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
