"""End-to-end drop-in models on the GPU vs the reference's recorded runs and the fp64 oracle (T3)."""
import math

import numpy as np
import pytest
import torch

from oracle import losses as ol, sa_torch

pytestmark = pytest.mark.gpu

KINDS = [("vonmises", "PointNetPPVonMises"), ("mvm", "PointNetPPMvM"), ("8dir", "PointNetPP8Dir"), ("xyz", "PointNetPPXYZ")]


def _make(pcoe, g, kind, cls, cuda):
    torch.manual_seed(1000)
    model = getattr(pcoe, cls)()
    model.drop.p = 0.0
    if kind == "mvm":
        sd = model.state_dict()
        sd["head_mu.weight"] = torch.from_numpy(g["mvm_head_mu_w"])
        sd["head_pi.weight"] = torch.from_numpy(g["mvm_head_pi_w"])
        model.load_state_dict(sd)
    return model.to(cuda).train()


def _loss(pcoe, kind, res, g, cuda):
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    if kind == "vonmises":
        return pcoe.kl_von_mises(res[0], res[1], t("mu_gt"), t("kappa_gt")).mean()
    if kind == "mvm":
        return pcoe.match_loss(res[0], res[1], res[2], t("vm_gt"), t("vm_gt"), t("K_gt")).mean()
    if kind == "8dir":
        return pcoe.kl_loss_per_sample_from_logits(res, t("p8")).mean()
    return (res[0] * torch.tensor([1.0, 2.0, 3.0], device=cuda)).sum() + (res[1] ** 2 * torch.tensor([0.5, -1.0, 2.0], device=cuda)).sum()


@pytest.mark.parametrize("kind,cls", KINDS)
def test_model_matches_reference_run(pcoe, golden, cuda, kind, cls):
    """Same seed -> same initial weights and the same host-replayed random subsets as the reference
    (T1: fps_idx bit-exact); outputs, loss, gradients and BN buffers match its recorded step."""
    g = golden("models")
    model = _make(pcoe, g, kind, cls, cuda)
    xyz = torch.from_numpy(g["xyz"]).to(cuda)
    torch.manual_seed(42)
    res = model(xyz)
    assert np.array_equal(model.sa1.last_fps_idx.cpu().numpy(), g[f"{kind}_fps1"])
    assert np.array_equal(model.sa2.last_fps_idx.cpu().numpy(), g[f"{kind}_fps2"])
    res_t = res if isinstance(res, tuple) else (res,)
    for i, r in enumerate(res_t):
        assert torch.allclose(r.cpu(), torch.from_numpy(g[f"{kind}_out{i}"]), rtol=2e-3, atol=2e-4), (kind, i)
    loss = _loss(pcoe, kind, res, g, cuda)
    want = float(g[f"{kind}_loss"])
    assert abs(float(loss) - want) <= 1e-3 * max(1.0, abs(want))       # north-star tolerance: 1e-3 relative
    loss.backward()
    gmax = max(float(g[k]) for k in g.files if k.startswith(f"{kind}_gnorm."))
    for name, p in model.named_parameters():
        gn = float(g[f"{kind}_gnorm.{name}"])
        if ".convs." in name and name.endswith("bias"):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        got = 0.0 if p.grad is None else float(p.grad.norm())
        # gradients are discontinuous in the forward values (max/ReLU routing): the fp32 reference is
        # itself 2-4 % from its fp64 run; norms must agree to that level
        assert abs(got - gn) <= 5e-2 * gn + 1e-4 * gmax, (name, got, gn)   # atol: gradients that are ~0 analytically
    for k in g.files:
        if k.startswith(f"{kind}_sd1."):
            got = model.state_dict()[k[len(kind) + 5:]].cpu()
            assert torch.allclose(got, torch.from_numpy(g[k]), rtol=1e-4, atol=1e-5), k


@pytest.mark.parametrize("kind,cls", KINDS)
def test_model_vs_fp64_oracle_T3(pcoe, golden, cuda, kind, cls):
    """T3: loss <= 1e-3 relative vs the fp64 oracle; gradients vs the fp64 oracle no worse than the
    fp32 reference's own deviation from fp64 (recorded golden grads), floor 1e-3."""
    g = golden("models")
    model = _make(pcoe, g, kind, cls, cuda)
    sd64 = sa_torch.clone_state(model.state_dict(), dtype=torch.float64, requires_grad=True)
    xyz = torch.from_numpy(g["xyz"])
    fps1, fps2 = torch.from_numpy(g[f"{kind}_fps1"]), torch.from_numpy(g[f"{kind}_fps2"])
    ores = sa_torch.model_forward(kind, sd64, xyz.double(), fps1, fps2)
    t = lambda k: torch.from_numpy(g[k])
    if kind == "vonmises":
        oloss = ol.kl_von_mises_single(ores[0], ores[1], t("mu_gt").double(), t("kappa_gt").double()).mean()
    elif kind == "mvm":
        oloss = ol.match_loss(ores[0], ores[1], ores[2], t("vm_gt").double(), t("K_gt")).mean()
    elif kind == "8dir":
        oloss = ol.soft_ce(ores, t("p8").double()).mean()
    else:
        oloss = (ores[0] * torch.tensor([1.0, 2.0, 3.0])).sum() + (ores[1] ** 2 * torch.tensor([0.5, -1.0, 2.0])).sum()
    oloss.backward()
    torch.manual_seed(42)
    res = model(xyz.to(cuda))
    loss = _loss(pcoe, kind, res, g, cuda)
    loss.backward()
    assert abs(float(loss) - float(oloss)) <= 1e-3 * max(1.0, abs(float(oloss)))
    for name in ("sa1.convs.0.weight", "sa2.convs.1.weight", "sa3.convs.2.weight", "fc1.weight"):
        o = sd64[name].grad
        mine = dict(model.named_parameters())[name].grad.double().cpu()
        ref32 = torch.from_numpy(g[f"{kind}_grad.{name}"]).double()
        rows = ref32.shape[0]
        self_dev = float((ref32 - o[:rows]).norm() / o[:rows].norm().clamp_min(1e-30))
        ours = float((mine[:rows] - o[:rows]).norm() / o[:rows].norm().clamp_min(1e-30))
        assert ours <= max(1e-3, 1.5 * self_dev), (name, ours, self_dev)


def test_state_dict_round_trip_and_eval(pcoe, golden, cuda):
    g = golden("models")
    a = _make(pcoe, g, "vonmises", "PointNetPPVonMises", cuda)
    xyz = torch.from_numpy(g["xyz"]).to(cuda)
    a(xyz)                                                     # one train step moves the BN buffers
    b = pcoe.PointNetPPVonMises().to(cuda)
    b.load_state_dict(a.state_dict(), strict=True)
    a.eval(); b.eval()
    idx1, idx2 = a.sa1.last_fps_idx, a.sa2.last_fps_idx
    with torch.no_grad():
        torch.manual_seed(5); ra = a(xyz)
        torch.manual_seed(5); rb = b(xyz)
    assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1])
    # eval vs the oracle in eval mode on the same checkpoint / indices
    sd = sa_torch.clone_state(a.state_dict())
    o = sa_torch.model_forward("vonmises", sd, xyz.cpu(), a.sa1.last_fps_idx.long().cpu(), a.sa2.last_fps_idx.long().cpu(), training=False)
    assert torch.allclose(ra[0].cpu(), o[0], rtol=2e-3, atol=2e-4) and torch.allclose(ra[1].cpu(), o[1], rtol=2e-3, atol=2e-4)


def test_mvm_accepts_both_layouts_and_zero_init_quirk(pcoe, cuda):
    """(B,3,N) input is accepted (pointnet_pp_mvM.py:15-27); with the reference's zero-initialised
    head_mu the fallback substitutes mu = 0 and blocks its gradient."""
    torch.manual_seed(0)
    m = pcoe.PointNetPPMvM().to(cuda).train()
    xyz = torch.randn(4, 200, 3, device=cuda)
    torch.manual_seed(1); mu, kappa, w = m(xyz)
    torch.manual_seed(1); mu2, kappa2, w2 = m(xyz.transpose(1, 2).contiguous())
    assert torch.equal(kappa, kappa2) and torch.equal(w, w2)
    assert (mu == 0).all() and torch.allclose(w, torch.full_like(w, 0.25))
    assert (kappa > 0).all() and (kappa <= 80).all()
    (mu.sum() + kappa.sum()).backward()
    assert float(m.head_mu.weight.grad.abs().max()) == 0.0
    with pytest.raises(ValueError):
        m(torch.zeros(2, 5, 7, device=cuda))


def test_other_heads_and_samplers_run(pcoe, cuda):
    xyz = torch.randn(4, 512, 3, device=cuda)
    for cls in ("PointNetPP", "PointNetPPXYZ_Schedmit", "PointNetPPFwd"):
        m = getattr(pcoe, cls)().to(cuda).train()
        r = m(xyz)
        r = r if isinstance(r, tuple) else (r,)
        sum(x.sum() for x in r).backward()
        assert all(torch.isfinite(x).all() for x in r)
        assert m.sa1.convs[0].weight.grad is not None and torch.isfinite(m.sa1.convs[0].weight.grad).all()
    m = pcoe.PointNetPP8Dir(sampler="fps", grouper="ball", radius=0.8).to(cuda).train()
    out = m(xyz / xyz.norm(dim=-1).amax())
    assert out.shape == (4, 8) and torch.isfinite(out).all()
    assert pcoe.DIRS_8.shape == (8, 3) and abs(float(pcoe.DIRS_8[1, 0]) - 0.7071) < 1e-6
    th, p = pcoe.mvm_density_on_grid(torch.zeros(2, 4, device=cuda), torch.ones(2, 4, device=cuda), torch.full((2, 4), 0.25, device=cuda))
    assert p.shape == (2, 359) and torch.allclose(p.sum(-1), torch.ones(2, device=cuda), atol=1e-5)


def test_mvm_head_fused_matches_torch_formulation(pcoe, cuda):
    """pcoe_mvm_head_fwd/_bwd vs the reference's elementwise formulation (models/pointnet_pp_mvM.py:91-125) in torch
    fp32: values and gradients, including the zero-vector fallback, the eps branch of normalize, the softplus
    threshold and the kappa clamp."""
    torch.manual_seed(5)
    model = pcoe.PointNetPPMvM().to(cuda)
    with torch.no_grad():
        for m in (model.head_pi, model.head_mu, model.head_kappa):
            m.weight.normal_(0, 0.3); m.bias.normal_(0, 0.3)
        model.head_kappa.bias[0] = 30.0           # softplus threshold (> 20) and clamp_max(80) untouched
        model.head_kappa.bias[1] = 200.0          # clamped: zero gradient
        model.head_mu.weight[0:2].zero_(); model.head_mu.bias[0:2].zero_()            # mu_raw == 0: fallback, no grad
        model.head_mu.weight[2:4].mul_(1e-6); model.head_mu.bias[2:4].mul_(1e-6)      # |v| < eps: u = v / eps
    feat = torch.randn(33, 256, device=cuda)
    fa, fb = feat.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    got = pcoe.models._MvMHead.apply(model.head_pi(fa), model.head_mu(fa), model.head_kappa(fa), model.temp, model.kappa_max)
    want = model._head_torch(fb)
    for g, w, name in zip(got, want, ("mu", "kappa", "weight")):
        assert torch.allclose(g, w, rtol=1e-5, atol=1e-6), name
    coef = [torch.randn_like(t) for t in got]
    sum((c * t).sum() for c, t in zip(coef, got)).backward()
    ga = [p.grad.clone() for p in model.parameters() if p.grad is not None]
    model.zero_grad()
    sum((c * t).sum() for c, t in zip(coef, want)).backward()
    gb = [p.grad.clone() for p in model.parameters() if p.grad is not None]
    assert torch.allclose(fa.grad, fb.grad, rtol=1e-4, atol=1e-5)
    assert len(ga) == len(gb) and len(ga) >= 6
    for a, b in zip(ga, gb):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4 * float(b.abs().max()) + 1e-6)


@pytest.mark.parametrize("direct", [False, True])
def test_mvm_fused_trunk_matches_torch_modules(pcoe, cuda, direct):
    """pcoe.trunk.MvMTrunkHead (fc1..heads as libpcoe fp32 kernels, one autograd node) vs the same parameters run
    through torch.nn modules (the reference's formulation, fp32): outputs and every parameter / input gradient.
    `direct`: gradients added straight into a FlatGradBuffer (two backward passes = accumulation)."""
    torch.manual_seed(7)
    model = pcoe.PointNetPPMvM().to(cuda).train()
    model.drop.p = 0.0                                    # dropout draws from different random streams
    with torch.no_grad():
        for m in (model.head_pi, model.head_mu, model.head_kappa):
            m.weight.normal_(0, 0.2); m.bias.normal_(0, 0.2)
        model.ln1.weight.uniform_(0.5, 1.5); model.ln1.bias.uniform_(-0.3, 0.3)
    for B in (64, 5):
        if direct:
            buf = pcoe.dp.FlatGradBuffer(model)
            assert model.direct_grad_accumulation
        x = torch.randn(B, 1024, device=cuda)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        coef = [torch.randn(B, 4, device=cuda) for _ in range(3)]
        reps = 2 if direct else 1
        model.zero_grad(set_to_none=not direct)
        if direct:
            buf.zero_()
        for _ in range(reps):
            got = model._fused(xa)
            sum((c * t).sum() for c, t in zip(coef, got)).backward()
        ga = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None and not n.startswith("sa")}
        for p in model.parameters():
            p.grad = None
        model.direct_grad_accumulation = False
        for _ in range(reps):
            feat = model.drop(torch.relu(model.ln2(model.fc2(model.drop(torch.relu(model.ln1(model.fc1(xb))))))))
            want = model._head_torch(feat)
            sum((c * t).sum() for c, t in zip(coef, want)).backward()
        for g, w, name in zip(got, want, ("mu", "kappa", "weight")):
            assert torch.allclose(g, w, rtol=1e-4, atol=1e-5), (B, name)
        assert torch.allclose(xa.grad, xb.grad, rtol=1e-3, atol=1e-5 * float(xb.grad.abs().max()) + 1e-7), B
        for n, p in model.named_parameters():
            if n.startswith("sa"):
                continue
            assert n in ga, n
            assert torch.allclose(ga[n], p.grad, rtol=1e-3, atol=1e-4 * float(p.grad.abs().max()) + 1e-7), (B, n)
        for p in model.parameters():
            p.grad = None


def test_mvm_fused_trunk_dropout_statistics_and_eval(pcoe, cuda):
    """Dropout inside the fused trunk: keep rate 1-p, survivors scaled by 1/(1-p), a fresh mask per call, gradients
    only through kept units; eval mode is deterministic and equals the torch modules."""
    torch.manual_seed(8)
    lib = pcoe._lib.load()
    B, N, p = 64, 512, 0.4
    x = torch.randn(B, N, device=cuda)
    g, b = torch.ones(N, device=cuda), torch.full((N,), 3.0, device=cuda)       # shift up: ReLU keeps everything
    out, mask = torch.empty_like(x), torch.empty(B, N, dtype=torch.uint8, device=cuda)
    mean, rstd = torch.empty(B, device=cuda), torch.empty(B, device=cuda)
    cnt = torch.zeros(1, dtype=torch.int64, device=cuda)
    masks = []
    for it in range(2):
        cnt.add_(1)
        pcoe._lib.check(lib.pcoe_ln_relu_dropout_fwd(x.data_ptr(), 1, None, None, g.data_ptr(), b.data_ptr(), B, N, 1e-5, p, 1, 1234, cnt.data_ptr(),
                                                     out.data_ptr(), mean.data_ptr(), rstd.data_ptr(), mask.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream))
        masks.append(mask.clone())
        keep = mask.float().mean().item()
        assert abs(keep - (1 - p)) < 0.02
        ref = torch.relu(torch.nn.functional.layer_norm(x, (N,), g, b, 1e-5)) / (1 - p)
        assert torch.allclose(out[mask.bool()], ref[mask.bool()], rtol=1e-5, atol=1e-5)
        assert float(out[~mask.bool()].abs().max()) == 0.0
    assert (masks[0] != masks[1]).float().mean().item() > 0.3                    # a new mask per call (device counter)
    model = pcoe.PointNetPPMvM().to(cuda).eval()
    xin = torch.randn(16, 1024, device=cuda)
    with torch.no_grad():
        a = model._fused(xin)
        bb = model._head_torch(model._global_feat.__func__(type("S", (), {"_sa_features": staticmethod(lambda z: z), "drop": model.drop,
                               "ln1": model.ln1, "ln2": model.ln2, "fc1": model.fc1, "fc2": model.fc2})(), xin))
    for u, v in zip(a, bb):
        assert torch.allclose(u, v, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("cls,B,N", [("PointNetPP8Dir", 4, 2048), ("PointNetPPXYZ", 4, 8192)])
def test_baseline_configs_c3_c4_shapes(pcoe, cuda, cls, B, N):
    """BASELINE configs[2] (8-direction head, 2048 points) and configs[3] (xyz head, 8192 points per cloud) at a
    reduced batch: the bf16 tensor-core path against the fp32 parity path of the same module on the same subsets
    (forward, loss gradient norms), and sampling / grouping indices against the CPU oracle."""
    from oracle import sampling as osmp
    torch.manual_seed(11)
    m32 = getattr(pcoe, cls)(precision="fp32").to(cuda).train()
    m16 = getattr(pcoe, cls)(precision="bf16").to(cuda).train()
    m16.load_state_dict(m32.state_dict())
    xyz = pcoe.synthetic.clouds(3, B, N, 5).to(cuda)
    outs = []
    for m in (m32, m16):
        torch.manual_seed(42)                  # same host randperm subsets and dropout masks
        res = m(xyz)
        res = res if isinstance(res, tuple) else (res,)
        gw = torch.Generator().manual_seed(9)   # fixed linear functional (sum of squares is constant for unit-vector heads)
        sum((r * torch.randn(r.shape, generator=gw).to(cuda)).sum() for r in res).backward()
        outs.append((torch.cat([r.flatten() for r in res]).detach(), m))
    idx32, idx16 = outs[0][1].sa1.last_fps_idx, outs[1][1].sa1.last_fps_idx
    assert torch.equal(idx32, idx16)
    # kNN sets of SA1 against the exact CPU neighbours (first cloud)
    new_xyz = xyz[0, idx32[0].long()].cpu().numpy()
    want, margin = osmp.knn(new_xyz[None], xyz[:1].cpu().numpy(), 32)
    got = outs[1][1].sa1.last_group_idx[:1].cpu().numpy()
    n, eq, excused, bad = osmp.knn_rows_match(got, want, margin)
    assert bad == 0 and excused <= max(1, n // 1000)
    # BatchNorm1d over a batch of 2-4 clouds amplifies bf16 noise in the trunk: compare the SA features instead
    torch.manual_seed(42)
    f32 = outs[0][1]._sa_features(xyz).detach()
    torch.manual_seed(42)
    f16 = outs[1][1]._sa_features(xyz).detach()
    rel = float((f16 - f32).norm() / f32.norm())
    print(f"\n[{cls} {B}x{N}] bf16 vs fp32 SA features rel {rel:.2e}")
    assert rel < 5e-2
    # eval mode (running statistics, BatchNorm folded): bf16 kernels (half-block epilogues on the 64-channel layers of
    # SA1 included) against the fp32 path, same subsets
    m32.eval(); m16.eval()
    with torch.no_grad():
        torch.manual_seed(43)
        e32 = m32._sa_features(xyz)
        torch.manual_seed(43)
        e16 = m16._sa_features(xyz)
    rel_eval = float((e16 - e32).norm() / e32.norm())
    print(f"[{cls} {B}x{N}] eval-mode bf16 vs fp32 SA features rel {rel_eval:.2e}")
    assert rel_eval < 5e-2
    g32 = torch.cat([p.grad.flatten() for p in outs[0][1].sa1.parameters() if p.grad is not None]).norm()
    g16 = torch.cat([p.grad.flatten() for p in outs[1][1].sa1.parameters() if p.grad is not None]).norm()
    assert torch.isfinite(g16) and 0.5 < float(g16 / g32) < 2.0


def test_single_cloud_full_resolution_eval_inference(pcoe, cuda):
    """The reference's inference call (train.py:228-246): ONE whole cloud (all ~10 000 vertices, no resampling) through
    PointNetPPXYZ_Schedmit in eval mode.  Checked against the torch-CPU oracle on the same checkpoint and the same
    host-generator subsets, with non-trivial running statistics (a trained checkpoint's BatchNorm state)."""
    B, N = 1, 10000
    torch.manual_seed(77)
    model = pcoe.PointNetPPXYZ_Schedmit()
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.uniform_(-0.2, 0.2, generator=g)
    sd = sa_torch.clone_state(model.state_dict(), dtype=torch.float64)
    model = model.to(cuda).eval()
    xyz = pcoe.synthetic.clouds(11, B, N)
    torch.manual_seed(5)
    with torch.no_grad():
        vy, vz = model(xyz.to(cuda))
    fps1, fps2 = model.sa1.last_fps_idx.long().cpu(), model.sa2.last_fps_idx.long().cpu()
    torch.manual_seed(5)                                     # the reference's draw order: sa1 then sa2 (:28)
    assert torch.equal(fps1, torch.stack([torch.randperm(N)[:128] for _ in range(B)]))
    assert torch.equal(fps2, torch.stack([torch.randperm(128)[:32] for _ in range(B)]))
    oy, oz = sa_torch.model_forward("schedmit", sd, xyz.double(), fps1, fps2, training=False, update_buffers=False)
    assert vy.shape == (1, 3) and vz.shape == (1, 3)
    assert float((vy.cpu().double() - oy).abs().max()) < 1e-3 and float((vz.cpu().double() - oz).abs().max()) < 1e-3
    assert abs(float(vy.norm()) - 1.0) < 1e-5               # unit vectors (F.normalize, Pointnet_pp_xyz_Schedmit.py:88-90)
