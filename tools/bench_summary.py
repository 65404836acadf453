"""Print the headline fields and the per-kernel table of a bench.py JSON line (stdin or file)."""
import json
import sys

src = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in src.strip().splitlines() if l.startswith("{")][-1])
for k in ("value", "ms_per_step", "e2e", "gpu_launches_per_step", "roofline", "cpu_baseline", "modes", "clocks"):
    print(k, d.get(k))
ks = sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])
print("total kernel ms", sum(k["ms_per_step"] for _, k in ks))
for n, k in ks[:50]:
    print(f"{n:28s} {k['ms_per_step'] * 1e3:8.1f} us  x{k['launches_per_step']:.0f}  {k.get('bound', '')} "
          f"{k.get('frac', 0):.3f} tf={k.get('tensor_frac', 0):.3f} gbs={k.get('gbs', 0):.0f}")
