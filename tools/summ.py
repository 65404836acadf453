import json,sys
for f in sys.argv[1:]:
    try:
        for l in open(f):
            if l.startswith("{"):
                d=json.loads(l); print(f, "value",round(d["value"]), "ms",round(d["ms_per_step"],3), "e2e",d.get("e2e"), "cpu",d.get("cpu_baseline"), "launches",d.get("gpu_launches"))
                tot=0
                for k,v in (d.get("kernels") or {}).items():
                    print(f"  {k:26s} n={v['launches_per_step']:4.1f} ms={v['ms_per_step']:.4f} {v.get('achieved',0):8.1f} {v.get('unit','')} frac={v.get('frac',0):.4f}")
                    tot+=v['ms_per_step']
                print("  total libpcoe ms", tot)
    except Exception as e: print(f, "ERR", e)
