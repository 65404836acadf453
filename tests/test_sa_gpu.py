"""Set-abstraction kernels (pcoe_sa_forward/backward through the drop-in module) vs the reference's
recorded results (tests/golden/sa.npz) and vs the fp64 oracle at the model's real layer shapes."""
import numpy as np
import pytest
import torch

from oracle import sa_torch

pytestmark = pytest.mark.gpu

# fp32 (CUDA-core) mode gates; the teacher-forced backward sees the same routing as the oracle
# gradients are discontinuous in the forward values (arg-max / ReLU re-routing between fp32 and fp64):
# SURVEY 7.3 measures 7.4e-4 for an exact-arithmetic fp32 implementation at the deepest weight
FWD_RTOL, GRAD_REL = 2e-5, 2e-3


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _build(pcoe, g, tag, cuda, precision="fp32"):
    B, N, S, K, D, c1, c2, c3, ga = g[f"{tag}_cfg"].tolist()
    layer = pcoe.PointNetSetAbstraction(S or None, K or None, D, [c1, c2, c3], group_all=bool(ga), precision=precision)
    sd = {k[len(tag) + 5:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith(f"{tag}_sd0.")}
    layer.load_state_dict(sd, strict=True)
    return layer.to(cuda), (B, N, S, K, D, ga)


@pytest.mark.parametrize("tag", ["small", "nofeat", "sa2", "gall"])
def test_sa_golden_train_forward_backward_buffers(pcoe, golden, cuda, tag):
    g = golden("sa")
    layer, (B, N, S, K, D, ga) = _build(pcoe, g, tag, cuda)
    layer.train()
    xyz = torch.from_numpy(g[f"{tag}_xyz"]).to(cuda)
    pts = torch.from_numpy(g[f"{tag}_pts"]).to(cuda).requires_grad_(True) if D else None
    fps = None if ga else torch.from_numpy(g[f"{tag}_fps"]).to(cuda)
    new_xyz, out = layer(xyz, pts, fps_idx=fps)
    if not ga:   # T1: the kNN kernel reproduces the reference's neighbour sets on the recorded centroids
        assert np.array_equal(np.sort(layer.last_group_idx.cpu().numpy(), -1), np.sort(g[f"{tag}_grp"], -1))
    assert out.shape == tuple(g[f"{tag}_out"].shape)
    assert torch.allclose(out.cpu(), torch.from_numpy(g[f"{tag}_out"]), rtol=1e-4, atol=2e-5)
    out.backward(torch.from_numpy(g[f"{tag}_gout"]).to(cuda))
    for name, p in layer.named_parameters():
        want = torch.from_numpy(g[f"{tag}_grad.{name}"])
        if name.startswith("convs") and name.endswith("bias"):
            assert float(p.grad.abs().max()) == 0.0            # cancelled exactly; the reference has rounding noise
            continue
        assert _rel(p.grad, want) < 2e-3, name                 # the fp32 reference itself is ~1e-3 from fp64 here
    if D:
        assert _rel(pts.grad, torch.from_numpy(g[f"{tag}_gpts"])) < 2e-3
    for k in g.files:
        if k.startswith(f"{tag}_sd1."):
            got = layer.state_dict()[k[len(tag) + 5:]].cpu()
            assert torch.allclose(got.float(), torch.from_numpy(g[k]).float(), rtol=1e-5, atol=1e-6), k


@pytest.mark.parametrize("tag", ["small", "nofeat", "sa2", "gall"])
def test_sa_golden_eval_forward(pcoe, golden, cuda, tag):
    g = golden("sa")
    layer, (B, N, S, K, D, ga) = _build(pcoe, g, tag, cuda)
    sd = layer.state_dict()
    for k in g.files:                                          # eval output was recorded after one train step
        if k.startswith(f"{tag}_sd1."):
            sd[k[len(tag) + 5:]] = torch.from_numpy(g[k]).to(cuda)
    layer.load_state_dict(sd)
    layer.eval()
    xyz = torch.from_numpy(g[f"{tag}_xyz"]).to(cuda)
    pts = torch.from_numpy(g[f"{tag}_pts"]).to(cuda) if D else None
    fps = None if ga else torch.from_numpy(g[f"{tag}_fps"]).to(cuda)
    with torch.no_grad():
        _, out = layer(xyz, pts, fps_idx=fps)
    assert torch.allclose(out.cpu(), torch.from_numpy(g[f"{tag}_out_eval"]), rtol=1e-4, atol=2e-5)
    rm_before = layer.bns[0].running_mean.clone()
    with torch.no_grad():
        layer(xyz, pts, fps_idx=fps)
    assert torch.equal(rm_before, layer.bns[0].running_mean)   # eval never touches the buffers


@pytest.mark.parametrize("shape", ["sa1", "sa2", "sa3"])
def test_sa_vs_fp64_oracle_at_model_shapes(pcoe, cuda, shape):
    """T2 (teacher-forced): same indices, same weights, fp64 oracle; fp32 kernels must be fp32-accurate."""
    torch.manual_seed(3)
    B = 8
    cfg = dict(sa1=(1024, 128, 32, 0, [64, 64, 128], False), sa2=(128, 32, 32, 128, [128, 128, 256], False),
               sa3=(32, None, None, 256, [256, 512, 1024], True))[shape]
    N, S, K, D, mlp, ga = cfg
    layer = pcoe.PointNetSetAbstraction(S, K, D, mlp, group_all=ga).to(cuda).train()
    with torch.no_grad():
        for bn in layer.bns:
            bn.weight.uniform_(-0.5, 1.5)
            bn.bias.uniform_(-0.2, 0.2)
    g = torch.Generator().manual_seed(17)
    xyz = torch.randn(B, N, 3, generator=g)
    xyz = xyz / xyz.norm(dim=-1).amax(1).view(B, 1, 1)
    pts = torch.randn(B, N, D, generator=g) if D else None
    sd0 = sa_torch.clone_state({f"sa.{k}": v for k, v in layer.state_dict().items()}, dtype=torch.float64, requires_grad=True)
    fps = None if ga else torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)])
    pts_c = pts.to(cuda).requires_grad_(True) if D else None
    _, out = layer(xyz.to(cuda), pts_c, fps_idx=None if ga else fps.to(cuda))
    grp = None if ga else layer.last_group_idx.long().cpu()
    opts = pts.double().requires_grad_(True) if D else None
    _, oy, _ = sa_torch.set_abstraction(sd0, "sa", xyz.double(), opts, group_all=ga, nsample=K, fps_idx=fps, group_idx=grp)
    assert _rel(out, oy) < FWD_RTOL
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.to(cuda))
    oy.backward(gout.double())
    for name, p in layer.named_parameters():
        if name.startswith("convs") and name.endswith("bias"):
            continue
        # sa1 has 8*128*128 maxima over 32 rows each: a handful of fp32-vs-fp64 arg-max / ReLU flips
        # each re-route one O(1) gradient element (measured 2.2e-3 on bns.1.bias), hence the wider gate
        assert _rel(p.grad, sd0[f"sa.{name}"].grad) < (5e-3 if shape == "sa1" else GRAD_REL), name
    if D:
        assert _rel(pts_c.grad, opts.grad) < GRAD_REL
    for i in range(3):
        for buf in ("running_mean", "running_var"):
            assert torch.allclose(getattr(layer.bns[i], buf).cpu().double(), sd0[f"sa.bns.{i}.{buf}"], rtol=1e-5, atol=1e-6)
        assert int(layer.bns[i].num_batches_tracked) == 1


def test_sa_errors(pcoe, cuda):
    layer = pcoe.PointNetSetAbstraction(None, None, 0, [8, 8, 8], group_all=True).to(cuda).train()
    with pytest.raises(NotImplementedError):
        layer(torch.zeros(2, 48, 3, device=cuda), None)        # 48 points: group size not a power of two
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        layer(torch.zeros(1, 1, 3, device=cuda), None)         # the reference's BatchNorm raises the same
    with pytest.raises(ValueError):
        layer(torch.zeros(2, 3, 32, device=cuda), None)        # (B,3,N) is not accepted by the SA layer


# ---- bf16 tensor-core mode (tcgen05, fp32 accumulate): stated tolerances ----------------------
# operands are rounded to bf16 (8-bit mantissa, 3.9e-3 relative): a chain of three GEMM+BN layers
# lands at ~1e-2 relative on the outputs and, through the discontinuous max/ReLU routing, at up to
# ~1e-1 on the deepest weight gradients (SURVEY 7.3 measured 3.5e-3..2e-2 and 0.36..0.44 end to end).
# With mixed-sign BatchNorm weights (negative channels pool through the group minimum) the worst tensor measured
# 0.21 (SA1, first BatchNorm bias); with the default all-positive init 0.18.
BF16_FWD, BF16_GRAD = 3e-2, 2.5e-1


@pytest.mark.parametrize("shape", ["sa1", "sa2", "sa3"])
def test_sa_bf16_tensor_core_vs_fp64_oracle(pcoe, cuda, shape):
    torch.manual_seed(3)
    B = 8
    N, S, K, D, mlp, ga = dict(sa1=(1024, 128, 32, 0, [64, 64, 128], False), sa2=(128, 32, 32, 128, [128, 128, 256], False),
                               sa3=(32, None, None, 256, [256, 512, 1024], True))[shape]
    layer = pcoe.PointNetSetAbstraction(S, K, D, mlp, group_all=ga, precision="bf16").to(cuda).train()
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():       # mixed-sign BatchNorm weights: negative channels pool through the group MINIMUM
        for bn in layer.bns:
            sign = torch.where(torch.rand(bn.weight.shape, generator=g) < 0.3, -1.0, 1.0)
            bn.weight.copy_((0.5 + torch.rand(bn.weight.shape, generator=g)) * sign)
            bn.bias.copy_(torch.rand(bn.bias.shape, generator=g) * 0.4 - 0.1)
    xyz = torch.randn(B, N, 3, generator=g)
    xyz = xyz / xyz.norm(dim=-1).amax(1).view(B, 1, 1)
    pts = torch.randn(B, N, D, generator=g) if D else None
    sd0 = sa_torch.clone_state({f"sa.{k}": v for k, v in layer.state_dict().items()}, dtype=torch.float64, requires_grad=True)
    fps = None if ga else torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)])
    pts_c = pts.to(cuda).requires_grad_(True) if D else None
    _, out = layer(xyz.to(cuda), pts_c, fps_idx=None if ga else fps.to(cuda))
    grp = None if ga else layer.last_group_idx.long().cpu()
    opts = pts.double().requires_grad_(True) if D else None
    _, oy, _ = sa_torch.set_abstraction(sd0, "sa", xyz.double(), opts, group_all=ga, nsample=K, fps_idx=fps, group_idx=grp)
    fwd = _rel(out, oy)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.to(cuda))
    oy.backward(gout.double())
    rels = {n: _rel(p.grad, sd0[f"sa.{n}"].grad) for n, p in layer.named_parameters()
            if not (n.startswith("convs") and n.endswith("bias"))}
    if D:
        rels["grad_feats"] = _rel(pts_c.grad, opts.grad)
    print(f"\n[bf16 {shape}] fwd rel {fwd:.2e}; grad rel " + ", ".join(f"{k}={v:.1e}" for k, v in rels.items()))
    assert fwd < BF16_FWD
    assert max(rels.values()) < BF16_GRAD
    for i in range(3):
        assert torch.allclose(layer.bns[i].running_var.cpu().double(), sd0[f"sa.bns.{i}.running_var"], rtol=3e-2, atol=1e-3)


def test_sa_bf16_eval_matches_fp32_eval(pcoe, golden, cuda):
    g = golden("sa")
    for tag in ("sa2", "gall", "small"):
        l32, (B, N, S, K, D, ga) = _build(pcoe, g, tag, cuda, "fp32")
        l16, _ = _build(pcoe, g, tag, cuda, "bf16")
        l32.eval(); l16.eval()
        xyz = torch.from_numpy(g[f"{tag}_xyz"]).to(cuda)
        pts = torch.from_numpy(g[f"{tag}_pts"]).to(cuda) if D else None
        fps = None if ga else torch.from_numpy(g[f"{tag}_fps"]).to(cuda)
        with torch.no_grad():
            _, a = l32(xyz, pts, fps_idx=fps)
            _, b = l16(xyz, pts, fps_idx=fps)
        assert _rel(b, a) < BF16_FWD, tag


@pytest.mark.parametrize("shape", ["sa1", "sa2"])
def test_sa_bf16_sparse_last_layer_matches_stored_y3(pcoe, cuda, shape, monkeypatch):
    """The v4 train path does not store the last layer's pre-activations y3: its BatchNorm backward is rewritten with
    the Gram matrix of the layer's input (sa_tc4.cuh, DySparse4).  A/B against the older formulation (y3 stored in
    bf16 and re-read; PCOE_SA_STORE_Y3=1) on the same inputs: identical forward, gradients equal to bf16 rounding."""
    B = 8
    N, S, K, D, mlp = dict(sa1=(1024, 128, 32, 0, [64, 64, 128]), sa2=(128, 32, 32, 128, [128, 128, 256]))[shape]
    g = torch.Generator().manual_seed(23)
    xyz = torch.randn(B, N, 3, generator=g)
    xyz = (xyz / xyz.norm(dim=-1).amax(1).view(B, 1, 1)).to(cuda)
    pts = torch.randn(B, N, D, generator=g).to(cuda) if D else None
    fps = torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)]).to(cuda)
    gout = None
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PCOE_SA_STORE_Y3", mode)
        torch.manual_seed(5)
        layer = pcoe.PointNetSetAbstraction(S, K, D, mlp, precision="bf16").to(cuda).train()
        with torch.no_grad():
            for bn in layer.bns:
                bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
        p = pts.clone().requires_grad_(True) if D else None
        _, out = layer(xyz, p, fps_idx=fps)
        if gout is None:
            gout = torch.randn(out.shape, generator=g).to(cuda)
        out.backward(gout)
        res[mode] = (out.detach(), {n: q.grad.clone() for n, q in layer.named_parameters()}, p.grad.clone() if D else None)
    assert torch.equal(res["0"][0], res["1"][0])
    rels = {n: _rel(res["0"][1][n], res["1"][1][n]) for n in res["0"][1] if not (n.startswith("convs") and n.endswith("bias"))}
    if D:
        rels["grad_feats"] = _rel(res["0"][2], res["1"][2])
    print(f"\n[bf16 sparse-l3 vs stored-y3 {shape}] " + ", ".join(f"{k}={v:.1e}" for k, v in rels.items()))
    assert max(rels.values()) < 2e-2


def test_sa_bf16_tail_tile_falls_back_to_stored_y3(pcoe, cuda):
    """Rows that do not fill whole 128-row tiles (B*S*K % 128 != 0): the Gram-matrix last layer needs whole tiles, so
    the layer keeps the stored-y3 kernels - same tolerances against the fp64 oracle."""
    torch.manual_seed(4)
    B, N, S, K, D, mlp = 3, 200, 10, 32, 0, [64, 64, 128]            # M = 960 rows = 7.5 tiles
    layer = pcoe.PointNetSetAbstraction(S, K, D, mlp, precision="bf16").to(cuda).train()
    g = torch.Generator().manual_seed(19)
    xyz = torch.randn(B, N, 3, generator=g)
    xyz = xyz / xyz.norm(dim=-1).amax(1).view(B, 1, 1)
    sd0 = sa_torch.clone_state({f"sa.{k}": v for k, v in layer.state_dict().items()}, dtype=torch.float64, requires_grad=True)
    fps = torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)])
    _, out = layer(xyz.to(cuda), None, fps_idx=fps.to(cuda))
    grp = layer.last_group_idx.long().cpu()
    _, oy, _ = sa_torch.set_abstraction(sd0, "sa", xyz.double(), None, group_all=False, nsample=K, fps_idx=fps, group_idx=grp)
    assert _rel(out, oy) < BF16_FWD
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.to(cuda))
    oy.backward(gout.double())
    rels = {n: _rel(p.grad, sd0[f"sa.{n}"].grad) for n, p in layer.named_parameters()
            if not (n.startswith("convs") and n.endswith("bias"))}
    print("\n[bf16 tail tile] " + ", ".join(f"{k}={v:.1e}" for k, v in rels.items()))
    assert max(rels.values()) < BF16_GRAD


@pytest.mark.parametrize("precision,ab_tol,oracle_tol", [("bf16", 2e-2, None), ("bf16x3", 1e-3, 5e-3)])
def test_sa1_layer1_weight_gradient_from_the_layer2_epilogue(pcoe, cuda, monkeypatch, precision, ab_tol, oracle_tol):
    """SA1 (no input features): dW1 is assembled from  dz1^T x0,  x0^T x0  and  sum x0  accumulated by the layer-2
    backward epilogue (sa_tc4.cuh MaskStatsW1 / sa_tc6.cuh MaskStatsW6) instead of a separate pass over dz1 and y1.
    A/B against the layer-1 backward kernel (PCOE_SA_BWD_L1_KERNEL=1) on the same inputs, and both against the fp64 oracle."""
    B, N, S, K, mlp = 8, 1024, 128, 32, [64, 64, 128]
    g = torch.Generator().manual_seed(29)
    xyz = torch.randn(B, N, 3, generator=g)
    xyz = xyz / xyz.norm(dim=-1).amax(1).view(B, 1, 1)
    fps = torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)])
    gout = torch.randn(B, S, mlp[-1], generator=g)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PCOE_SA_BWD_L1_KERNEL", mode)
        torch.manual_seed(6)
        layer = pcoe.PointNetSetAbstraction(S, K, 0, mlp, precision=precision).to(cuda).train()
        with torch.no_grad():
            for bn in layer.bns:
                bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
        sd0 = sa_torch.clone_state({f"sa.{k}": v for k, v in layer.state_dict().items()}, dtype=torch.float64, requires_grad=True)
        _, out = layer(xyz.to(cuda), None, fps_idx=fps.to(cuda))
        out.backward(gout.to(cuda))
        res[mode] = (out.detach(), {n: q.grad.clone() for n, q in layer.named_parameters()}, sd0, layer.last_group_idx.long().cpu())
    assert torch.equal(res["0"][0], res["1"][0])
    for n in ("convs.0.weight", "bns.0.weight", "bns.0.bias", "convs.1.weight", "convs.2.weight"):
        r = _rel(res["0"][1][n], res["1"][1][n])
        print(f"[W1-in-epilogue vs layer-1 kernel, {precision}] {n}: {r:.1e}")
        assert r < ab_tol, n
    sd0, grp = res["0"][2], res["0"][3]
    _, oy, _ = sa_torch.set_abstraction(sd0, "sa", xyz.double(), None, group_all=False, nsample=K, fps_idx=fps, group_idx=grp)
    oy.backward(gout.double())
    r_new = _rel(res["0"][1]["convs.0.weight"], sd0["sa.convs.0.weight"].grad)
    r_old = _rel(res["1"][1]["convs.0.weight"], sd0["sa.convs.0.weight"].grad)
    print(f"[dW1 vs fp64 oracle, {precision}] epilogue path {r_new:.2e}, layer-1 kernel {r_old:.2e}")
    assert r_new < (BF16_GRAD if oracle_tol is None else oracle_tol)


# ---- bf16x3 mode (tcgen05, split operands, fp32 stored activations): the fp32-accurate tensor-core mode ----------
# every operand keeps 16 significant bits (hi + lo bf16 planes), the dropped lo*lo product is 2^-16 relative, the
# accumulator is fp32: a layer output lands at ~1e-5 relative - below the 5e-5 .. 1.8e-4 the reference's own fp32 run
# is away from its fp64 run - and the weight gradients (discontinuous in the forward values through the
# max / ReLU routing) at the level of the CUDA-core fp32 mode.
X3_FWD, X3_GRAD = 5e-5, 5e-3


@pytest.mark.parametrize("shape", ["sa1", "sa2", "sa3"])
def test_sa_bf16x3_tensor_core_vs_fp64_oracle(pcoe, cuda, shape):
    torch.manual_seed(3)
    B = 8
    N, S, K, D, mlp, ga = dict(sa1=(1024, 128, 32, 0, [64, 64, 128], False), sa2=(128, 32, 32, 128, [128, 128, 256], False),
                               sa3=(32, None, None, 256, [256, 512, 1024], True))[shape]
    layer = pcoe.PointNetSetAbstraction(S, K, D, mlp, group_all=ga, precision="bf16x3").to(cuda).train()
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():       # mixed-sign BatchNorm weights: negative channels pool through the group MINIMUM
        for bn in layer.bns:
            sign = torch.where(torch.rand(bn.weight.shape, generator=g) < 0.3, -1.0, 1.0)
            bn.weight.copy_((0.5 + torch.rand(bn.weight.shape, generator=g)) * sign)
            bn.bias.copy_(torch.rand(bn.bias.shape, generator=g) * 0.4 - 0.1)
    xyz = torch.randn(B, N, 3, generator=g)
    xyz = xyz / xyz.norm(dim=-1).amax(1).view(B, 1, 1)
    pts = torch.randn(B, N, D, generator=g) if D else None
    sd0 = sa_torch.clone_state({f"sa.{k}": v for k, v in layer.state_dict().items()}, dtype=torch.float64, requires_grad=True)
    fps = None if ga else torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)])
    pts_c = pts.to(cuda).requires_grad_(True) if D else None
    _, out = layer(xyz.to(cuda), pts_c, fps_idx=None if ga else fps.to(cuda))
    grp = None if ga else layer.last_group_idx.long().cpu()
    opts = pts.double().requires_grad_(True) if D else None
    _, oy, _ = sa_torch.set_abstraction(sd0, "sa", xyz.double(), opts, group_all=ga, nsample=K, fps_idx=fps, group_idx=grp)
    fwd = _rel(out, oy)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.to(cuda))
    oy.backward(gout.double())
    rels = {n: _rel(p.grad, sd0[f"sa.{n}"].grad) for n, p in layer.named_parameters()
            if not (n.startswith("convs") and n.endswith("bias"))}
    if D:
        rels["grad_feats"] = _rel(pts_c.grad, opts.grad)
    print(f"\n[bf16x3 {shape}] fwd rel {fwd:.2e}; grad rel " + ", ".join(f"{k}={v:.1e}" for k, v in rels.items()))
    assert fwd < X3_FWD
    assert max(rels.values()) < X3_GRAD
    for i in range(3):
        for buf in ("running_mean", "running_var"):
            assert torch.allclose(getattr(layer.bns[i], buf).cpu().double(), sd0[f"sa.bns.{i}.{buf}"], rtol=1e-4, atol=1e-5)
        assert int(layer.bns[i].num_batches_tracked) == 1


def test_sa_bf16x3_eval_and_ragged_tile(pcoe, cuda):
    """Eval mode (running statistics folded) in bf16x3 equals the fp32 mode; B = 1 group-all gives M = 32 rows, a
    quarter of one 128-point tile (the B = 1 inference shape of train.py:228-246 after SA2)."""
    torch.manual_seed(5)
    for (B, N, S, K, D, mlp, ga) in [(3, 256, 32, 32, 64, [64, 128, 128], False), (1, 32, None, None, 256, [256, 512, 1024], True)]:
        a = pcoe.PointNetSetAbstraction(S, K, D, mlp, group_all=ga, precision="fp32").to(cuda)
        b = pcoe.PointNetSetAbstraction(S, K, D, mlp, group_all=ga, precision="bf16x3").to(cuda)
        with torch.no_grad():
            for bn in a.bns:
                bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5); bn.weight.uniform_(-1, 1.5)
        b.load_state_dict(a.state_dict())
        a.eval(); b.eval()
        xyz = torch.rand(B, N, 3, device=cuda)
        pts = torch.randn(B, N, D, device=cuda)
        fps = None if ga else torch.stack([torch.randperm(N)[:S] for _ in range(B)]).to(cuda)
        with torch.no_grad():
            _, ya = a(xyz, pts, fps_idx=fps)
            _, yb = b(xyz, pts, fps_idx=fps)
        assert _rel(yb, ya) < X3_FWD, (B, N)
