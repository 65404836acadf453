"""Debug-build helper: per-role clock trace of CTA 0 of the v4 kernels (tools/build_trace.sh).
usage: PCOE_LIB=.../libpcoe_trace.so python tools/trace_sa.py sa1|sa2 [bwd]"""
import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcoe
lib = pcoe._lib.load()
lib.pcoe_debug_trace.restype = C.c_int
lib.pcoe_debug_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
cuda = torch.device("cuda:0")
B = 64
which = sys.argv[1] if len(sys.argv) > 1 else "sa1"
bwd = len(sys.argv) > 2
N, S, K, D, mlp = dict(sa1=(1024, 128, 32, 0, [64, 64, 128]), sa2=(128, 32, 32, 128, [128, 128, 256]))[which]
layer = pcoe.PointNetSetAbstraction(S, K, D, mlp, precision="bf16", sampler="randperm_device").to(cuda).train()
xyz = pcoe.synthetic.clouds(1, B, N, 0).to(cuda)
pts = torch.randn(B, N, D, device=cuda, requires_grad=True) if D else None
names = {0: "start", 1: "prologue_done", 2: "end", 10: "  prod_store_begin", 11: "  prod_store_end", 12: "  prod_loads_issued", 20: "    mma_full", 21: "    mma_go", 22: "    mma_issued",
         30: "epi_begin", 31: "epi_end", 32: "epi_finish", 35: "epi_arrived", 33: "epi_ld_done", 34: "epi_block_done", 36: "epi_dw_flush"}
for it in range(3):
    _, out = layer(xyz, pts)
    out.sum().backward()
    torch.cuda.synchronize()
    lib.pcoe_debug_trace(None, 0, 1)
_, out = layer(xyz, pts)
torch.cuda.synchronize()
if bwd:
    lib.pcoe_debug_trace(None, 0, 1)
    out.sum().backward()
    torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 2700))()
n = lib.pcoe_debug_trace(buf, 2700, 1)
ev = [(buf[3 * i + 2], buf[3 * i], buf[3 * i + 1]) for i in range(n)]
ev.sort()   # kernels are separated in time: sort by clock, split at tag 0
ker, cur = [], []
for t, tag, idx in ev:
    if tag == 0 and cur:
        ker.append(cur); cur = []
    cur.append((t, tag, idx))
ker.append(cur)
for k, evs in enumerate(ker):
    t0 = min(e[0] for e in evs)
    print(f"--- kernel {k}: {len(evs)} events, span {(max(e[0] for e in evs) - t0) / 1.9e3:.2f} us (at 1.9 GHz)")
    for t, tag, idx in sorted(evs):
        print(f"   {(t - t0) / 1.9e3:8.2f} us  {names.get(tag, tag):24s} {idx}")
