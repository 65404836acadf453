import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcoe
from oracle import sa_torch
torch.manual_seed(3)
cuda = torch.device("cuda:0")
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
shapes = sys.argv[1:] or ["sa1", "sa2"]
for shape in shapes:
    B = int(os.environ.get("DBG_B", "8"))
    N, S, K, D, mlp, ga = dict(sa1=(1024, 128, 32, 0, [64, 64, 128], False), sa2=(128, 32, 32, 128, [128, 128, 256], False),
                               sa3=(32, None, None, 256, [256, 512, 1024], True))[shape]
    layer = pcoe.PointNetSetAbstraction(S, K, D, mlp, group_all=ga, precision="bf16").to(cuda).train()
    g = torch.Generator().manual_seed(17)
    xyz = torch.randn(B, N, 3, generator=g)
    xyz = xyz / xyz.norm(dim=-1).amax(1).view(B, 1, 1)
    pts = torch.randn(B, N, D, generator=g) if D else None
    sd0 = sa_torch.clone_state({f"sa.{k}": v for k, v in layer.state_dict().items()}, dtype=torch.float64, requires_grad=True)
    fps = None if ga else torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)])
    pts_c = pts.to(cuda).requires_grad_(True) if D else None
    print(shape, "forward...", flush=True)
    _, out = layer(xyz.to(cuda), pts_c, fps_idx=None if ga else fps.to(cuda))
    torch.cuda.synchronize()
    print(shape, "forward done", flush=True)
    grp = None if ga else layer.last_group_idx.long().cpu()
    opts = pts.double().requires_grad_(True) if D else None
    _, oy, _ = sa_torch.set_abstraction(sd0, "sa", xyz.double(), opts, group_all=ga, nsample=K, fps_idx=fps, group_idx=grp)
    print(shape, "fwd rel", rel(out, oy), flush=True)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.to(cuda))
    torch.cuda.synchronize()
    print(shape, "backward done", flush=True)
    oy.backward(gout.double())
    rels = {n: rel(p.grad, sd0[f"sa.{n}"].grad) for n, p in layer.named_parameters()
            if not (n.startswith("convs") and n.endswith("bias"))}
    if D:
        rels["grad_feats"] = rel(pts_c.grad, opts.grad)
    print(shape, "grad rel " + ", ".join(f"{k}={v:.1e}" for k, v in rels.items()), flush=True)
