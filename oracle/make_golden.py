"""Pin the oracle against the UNMODIFIED reference and write tests/golden/*.npz.

Run in the build container (needs /root/reference or $PCOE_REF):  python -m oracle.make_golden
Every oracle function is first checked against the reference function it restates (assert), then
the reference's own outputs are stored so that the GPU box - which has no reference tree - can
replay them.  Nothing is copied from the reference: model code is imported from it, the three loss
functions are lifted at run time by ast-slicing the training scripts (they cannot be imported:
they need matplotlib and /home/pablo paths).
"""
from __future__ import annotations

import ast
import importlib.util
import math
import os
import re
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("PCOE_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import losses as olosses, sa_torch, sampling  # noqa: E402


# ---------------------------------------------------------------------------------------------
def load_reference():
    sys.path.insert(0, REF)
    import models.base as base                                  # noqa
    import models.pointnet_pp_8dir as m8                        # noqa
    from models.pointnet_pp_vonMises import PointNetPPVonMises  # noqa
    from models.pointnet_pp_mvM import PointNetPPMvM            # noqa
    from models.Pointnet_pp_xyz import PointNetPPXYZ            # noqa
    spec = importlib.util.spec_from_file_location("pp_demo", os.path.join(REF, "PointNet++Demo.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    return dict(base=base, m8=m8, VonMises=PointNetPPVonMises, MvM=PointNetPPMvM, XYZ=PointNetPPXYZ, demo=demo)


def lift(path: str, names: list[str], extra: dict) -> dict:
    """exec only the named top-level functions of a reference script."""
    src = open(os.path.join(REF, path), encoding="utf-8").read()
    tree = ast.parse(src)
    ns = dict(torch=torch, math=math, F=F, np=np, **extra)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def unit_clouds(seed: int, B: int, N: int) -> torch.Tensor:
    """centred, unit-ball, duplicate-free clouds (SURVEY 8d generator)."""
    g = torch.Generator("cpu").manual_seed(seed)
    x = torch.randn(B, N, 3, generator=g)
    x = x - x.mean(1, keepdim=True)
    return (x / x.norm(dim=-1).amax(1).view(B, 1, 1)).contiguous()


# ---------------------------------------------------------------------------------------------
def golden_sampling(ref) -> dict:
    demo, base = ref["demo"], ref["base"]
    out = {}
    cases = [("a", unit_clouds(11, 3, 256)), ("b", torch.rand(2, 301, 3, generator=torch.Generator().manual_seed(5)))]
    for tag, xyz in cases:
        B, N, _ = xyz.shape
        S = 64
        start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(3))
        real_randint = torch.randint
        torch.randint = lambda *a, **k: start.clone()           # PointNet++Demo.py:20 draws the start index
        try:
            fps_ref = demo.farthest_point_sample(xyz, S)
        finally:
            torch.randint = real_randint
        fps_or = sampling.farthest_point_sample(xyz.numpy(), S, start.numpy())
        assert np.array_equal(fps_ref.numpy(), fps_or), "FPS oracle != reference"
        new_xyz = demo.index_points(xyz, fps_ref)
        out[f"{tag}_xyz"], out[f"{tag}_start"], out[f"{tag}_fps"] = xyz.numpy(), start.numpy(), fps_ref.numpy()
        for r, ns in ((0.2, 16), (0.4, 32), (0.05, 8)):
            bq_ref = demo.query_ball_point(r, ns, xyz, new_xyz)
            bq_or = sampling.ball_query(r, ns, xyz.numpy(), new_xyz.numpy())
            assert np.array_equal(bq_ref.numpy(), bq_or), "ball-query oracle != reference"
            out[f"{tag}_ball_{r}_{ns}"] = bq_ref.numpy()
        k = 32
        knn_ref = base.query_ball_point(new_xyz, xyz, k).numpy()
        knn_or, margin = sampling.knn(new_xyz.numpy(), xyz.numpy(), k)
        n, eq, tie, bad = sampling.knn_rows_match(knn_ref, knn_or, margin)
        assert bad == 0 and eq + tie == n, f"kNN oracle != reference ({bad} rows)"
        print(f"  kNN[{tag}]: {eq}/{n} rows set-equal to the reference, {tie} near-tie rows")
        out[f"{tag}_knn_sorted"] = np.sort(knn_ref, -1)
    return out


# ---------------------------------------------------------------------------------------------
def run_ref_sa(ref, layer, xyz, points, grad_out_seed=7):
    """forward+backward of a reference SA module, capturing fps_idx / group idx via the module
    globals it resolves (pointnet_pp_8dir.py:4,29-31)."""
    m8 = ref["m8"]
    cap = {}
    real_ip, real_qb = m8.index_points, m8.query_ball_point

    def ip(p, idx):
        if idx.dim() == 2 and "fps" not in cap:
            cap["fps"] = idx.clone()
        return real_ip(p, idx)

    def qb(new_xyz, xyz_, k):
        idx = real_qb(new_xyz, xyz_, k)
        cap["grp"] = idx.clone()
        return idx

    m8.index_points, m8.query_ball_point = ip, qb
    try:
        if points is not None:
            points = points.clone().requires_grad_(True)
        new_xyz, out = layer(xyz, points)
    finally:
        m8.index_points, m8.query_ball_point = real_ip, real_qb
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(grad_out_seed))
    out.backward(g)
    return new_xyz, out, g, points.grad if points is not None else None, cap


def golden_sa(ref) -> dict:
    SA = ref["m8"].PointNetSetAbstraction
    out = {}
    cfgs = {
        "small": dict(B=2, N=64, S=16, K=8, D=4, mlp=[16, 16, 32], group_all=False),
        "nofeat": dict(B=2, N=96, S=16, K=16, D=0, mlp=[16, 24, 32], group_all=False),
        "sa2": dict(B=2, N=128, S=32, K=32, D=128, mlp=[128, 128, 256], group_all=False),
        "gall": dict(B=3, N=32, S=None, K=None, D=8, mlp=[16, 32, 64], group_all=True),
    }
    for tag, c in cfgs.items():
        torch.manual_seed(100 + len(tag))
        layer = SA(c["S"], c["K"], c["D"], c["mlp"], group_all=c["group_all"])
        with torch.no_grad():                                 # non-trivial BN affine + running stats
            for bn in layer.bns:
                bn.weight.uniform_(-1.0, 1.5)                 # negative gammas exercise the min route
                bn.bias.uniform_(-0.3, 0.3)
                bn.running_mean.uniform_(-0.2, 0.2)
                bn.running_var.uniform_(0.5, 1.5)
        sd0 = {k: v.detach().clone() for k, v in layer.state_dict().items()}
        xyz = unit_clouds(21, c["B"], c["N"])
        pts = torch.randn(c["B"], c["N"], c["D"], generator=torch.Generator().manual_seed(9)) if c["D"] else None
        layer.train()
        new_xyz, y, g, gpts, cap = run_ref_sa(ref, layer, xyz, pts)
        # oracle restatement on the same indices, same initial state
        sd = {f"sa.{k}": v.clone() for k, v in sd0.items()}
        osd = sa_torch.clone_state(sd, requires_grad=True)
        opts = pts.clone().requires_grad_(True) if pts is not None else None
        _, oy, _ = sa_torch.set_abstraction(osd, "sa", xyz, opts, group_all=c["group_all"], nsample=c["K"],
                                            fps_idx=cap.get("fps"), group_idx=cap.get("grp"), training=True)
        oy.backward(g)
        assert torch.allclose(oy, y, rtol=1e-4, atol=1e-5), f"SA oracle forward != reference ({tag})"
        for name, p in layer.named_parameters():
            og = osd[f"sa.{name}"].grad
            if name.startswith("convs") and name.endswith("bias"):
                # cancelled exactly by the batch-mean subtraction: both sides are rounding noise
                wn = 1.0 + float(dict(layer.named_parameters())[name.replace("bias", "weight")].grad.norm())
                assert float(og.abs().max()) < 1e-3 * wn and float(p.grad.abs().max()) < 1e-3 * wn, f"conv bias grad ({tag})"
                continue
            rel = float((og - p.grad).norm() / p.grad.norm().clamp_min(1e-12))
            assert rel < 2e-3, f"SA oracle grad {name} ({tag}): rel {rel:.2e}"
        for k, v in layer.state_dict().items():
            if "running" in k:
                assert torch.allclose(osd[f"sa.{k}"], v, rtol=1e-5, atol=1e-6), f"SA oracle buffer {k} ({tag})"
        # eval-mode forward with the post-step buffers
        layer.eval()
        with torch.no_grad():
            real = ref["m8"].index_points
            _, y_eval, _, _, _ = (None, None, None, None, None)
        # eval forward needs the same fps_idx: feed it through the RNG-free oracle check only
        with torch.no_grad():
            esd = sa_torch.clone_state({f"sa.{k}": v for k, v in layer.state_dict().items()})
            _, oy_eval, _ = sa_torch.set_abstraction(esd, "sa", xyz, pts, group_all=c["group_all"], nsample=c["K"],
                                                     fps_idx=cap.get("fps"), group_idx=cap.get("grp"), training=False)
            x = (torch.cat([xyz.unsqueeze(1), pts.unsqueeze(1)], -1) if pts is not None else xyz.unsqueeze(1)) if c["group_all"] else None
            if not c["group_all"]:
                nx = real(xyz, cap["fps"])
                x = real(xyz, cap["grp"]) - nx.unsqueeze(2)
                if pts is not None:
                    x = torch.cat([x, real(pts, cap["grp"])], -1)
            x = x.permute(0, 3, 1, 2)
            for conv, bn in zip(layer.convs, layer.bns):      # the reference's own modules, eval mode
                x = F.relu(bn(conv(x)))
            y_eval = torch.max(x, 3)[0].permute(0, 2, 1)
        assert torch.allclose(oy_eval, y_eval, rtol=1e-4, atol=1e-5), f"SA oracle eval != reference ({tag})"
        out[f"{tag}_cfg"] = np.array([c["B"], c["N"], c["S"] or 0, c["K"] or 0, c["D"], *c["mlp"], int(c["group_all"])])
        out[f"{tag}_xyz"] = xyz.numpy()
        if pts is not None:
            out[f"{tag}_pts"], out[f"{tag}_gpts"] = pts.numpy(), gpts.numpy()
        if not c["group_all"]:
            out[f"{tag}_fps"], out[f"{tag}_grp"] = cap["fps"].numpy(), cap["grp"].numpy()
        out[f"{tag}_out"], out[f"{tag}_gout"], out[f"{tag}_out_eval"] = y.detach().numpy(), g.numpy(), y_eval.numpy()
        for k, v in sd0.items():
            out[f"{tag}_sd0.{k}"] = v.numpy()
        for k, v in layer.state_dict().items():
            if "running" in k or "num_batches" in k:
                out[f"{tag}_sd1.{k}"] = v.numpy()
        for name, p in layer.named_parameters():
            out[f"{tag}_grad.{name}"] = p.grad.numpy()
        print(f"  SA[{tag}]: oracle == reference (fwd, bwd, buffers, eval)")
    return out


# ---------------------------------------------------------------------------------------------
def golden_models(ref, loss_fns) -> dict:
    out = {}
    B, N = 4, 256
    xyz = unit_clouds(31, B, N)
    out["xyz"] = xyz.numpy()
    g = torch.Generator().manual_seed(77)
    mu_gt = (torch.rand(B, generator=g) * 2 - 1) * math.pi
    kappa_gt = torch.tensor([8.0, 0.0, 8.0, 8.0])
    K_gt = torch.tensor([1, 2, 4, 1])
    vm_gt = torch.zeros(B, 4, 3)
    for b in range(B):
        for j in range(int(K_gt[b])):
            vm_gt[b, j] = torch.tensor([math.remainder(float(mu_gt[b]) + j * 2 * math.pi / int(K_gt[b]), 2 * math.pi), 8.0, 1.0 / int(K_gt[b])])
    p8 = F.softmax(torch.randn(B, 8, generator=g), dim=1)
    out.update(mu_gt=mu_gt.numpy(), kappa_gt=kappa_gt.numpy(), K_gt=K_gt.numpy(), vm_gt=vm_gt.numpy(), p8=p8.numpy())
    kinds = {"vonmises": ref["VonMises"], "mvm": ref["MvM"], "8dir": ref["m8"].PointNetPP8Dir, "xyz": ref["XYZ"]}
    for kind, cls in kinds.items():
        torch.manual_seed(1000)
        model = cls()
        model.drop.p = 0.0                                    # dropout off: parity runs are RNG-free past sampling
        if kind == "mvm":                                     # un-zero the heads so mu/weight gradients are live
            with torch.no_grad():
                gen = torch.Generator().manual_seed(5)
                model.head_mu.weight.normal_(0, 0.05, generator=gen)
                model.head_pi.weight.normal_(0, 0.05, generator=gen)
        model.train()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        torch.manual_seed(42)
        fps1 = torch.stack([torch.randperm(N)[:128] for _ in range(B)])
        fps2 = torch.stack([torch.randperm(128)[:32] for _ in range(B)])
        torch.manual_seed(42)                                 # the model draws the same permutations (:28)
        res = model(xyz)

        def loss_of(r):
            if kind == "vonmises":
                return loss_fns["single"](r[0], r[1], mu_gt, kappa_gt).mean()
            if kind == "mvm":
                return loss_fns["match"](r[0], r[1], r[2], vm_gt, vm_gt, K_gt).mean()
            if kind == "8dir":
                return loss_fns["ce"](r, p8).mean()
            return (r[0] * torch.tensor([1.0, 2.0, 3.0])).sum() + (r[1] ** 2 * torch.tensor([0.5, -1.0, 2.0])).sum()

        loss = loss_of(res)
        loss.backward()
        # oracle on the same checkpoint and replayed indices
        osd = sa_torch.clone_state(sd0, requires_grad=True)
        ores = sa_torch.model_forward(kind, osd, xyz, fps1, fps2, training=True)
        if kind == "vonmises":
            oloss = olosses.kl_von_mises_single(ores[0], ores[1], mu_gt, kappa_gt).mean()
        elif kind == "mvm":
            oloss = olosses.match_loss(ores[0], ores[1], ores[2], vm_gt, K_gt).mean()
        elif kind == "8dir":
            oloss = olosses.soft_ce(ores, p8).mean()
        else:
            oloss = loss_of(ores)
        oloss.backward()
        res_t = res if isinstance(res, tuple) else (res,)
        ores_t = ores if isinstance(ores, tuple) else (ores,)
        for a, b_ in zip(res_t, ores_t):
            assert torch.allclose(a, b_, rtol=1e-3, atol=1e-4), f"model oracle output != reference ({kind})"
        assert abs(float(loss) - float(oloss)) <= 1e-4 * max(1.0, abs(float(loss))), f"loss ({kind})"
        for i, r in enumerate(res_t):
            out[f"{kind}_out{i}"] = r.detach().numpy()
        out[f"{kind}_loss"] = np.array(float(loss))
        out[f"{kind}_fps1"], out[f"{kind}_fps2"] = fps1.numpy(), fps2.numpy()
        out[f"{kind}_sd_keys"] = np.array([f"{k}|{'x'.join(map(str, v.shape))}" for k, v in sd0.items()])
        out[f"{kind}_sd_checksum"] = np.array([float(sum(v.double().abs().sum() for v in sd0.values() if v.is_floating_point()))])
        keys = ["sa1.convs.0.weight", "sa2.convs.1.weight", "sa3.convs.2.weight", "sa3.bns.2.weight", "fc1.weight"]
        for name, p in model.named_parameters():
            gn = 0.0 if p.grad is None else float(p.grad.norm())
            out[f"{kind}_gnorm.{name}"] = np.array(gn)
            if name in keys:                                  # big tensors: first 16 rows only (fixture size)
                out[f"{kind}_grad.{name}"] = (p.grad[:16] if p.grad.numel() > 100_000 else p.grad).numpy()
        for k, v in model.state_dict().items():
            if "running" in k and k.startswith("sa"):
                out[f"{kind}_sd1.{k}"] = v.numpy()
        if kind == "mvm":
            out["mvm_head_mu_w"], out["mvm_head_pi_w"] = sd0["head_mu.weight"].numpy(), sd0["head_pi.weight"].numpy()
        print(f"  model[{kind}]: oracle == reference, loss {float(loss):.6f}")
    return out


# ---------------------------------------------------------------------------------------------
def parse_debug_log(per_k: int = 120):
    """Known-answer vectors the reference committed: results/multi_peak_vonMises_KL_debug/debug_log.txt,
    written by train_multi_peaks_vonMises_KL_debug.py:90-118.  Records: K, mu_p, kappa_p, w_p, mu_g,
    kappa_g -> matched costs, matched weights."""
    path = os.path.join(REF, "results", "multi_peak_vonMises_KL_debug", "debug_log.txt")
    num = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+|nan|inf)"
    recs, cur = [], None
    take = {1: 0, 2: 0, 4: 0}
    with open(path, encoding="utf-8") as f:
        buf = ""
        for line in f:
            m = re.match(r"\[Batch \d+\] K = (\d+)", line)
            if m:
                cur = {"K": int(m.group(1))}
                continue
            if cur is None:
                continue
            line = line.strip()
            for key, tag in (("mu_p", "μp:"), ("kappa_p", "κp:"), ("w_p", "wp:"), ("mu_g", "μg:"), ("kappa_g", "κg:"),
                             ("cost", "matched cost:")):
                if line.startswith(tag):
                    cur[key] = [float(x) for x in re.findall(num, line[len(tag):])]
            if line.startswith("matched_ws:"):
                K = cur["K"]
                ok = all(k in cur and len(cur[k]) == K for k in ("mu_p", "kappa_p", "w_p", "mu_g", "kappa_g", "cost"))
                if ok and K in take and take[K] < per_k:
                    take[K] += 1
                    recs.append(cur)
                cur = None
            if all(v >= per_k for v in take.values()):
                break
    n = len(recs)
    arr = {k: np.zeros((n, 4), dtype=np.float64) for k in ("mu_p", "kappa_p", "w_p", "mu_g", "kappa_g", "cost")}
    K = np.zeros(n, dtype=np.int64)
    for i, r in enumerate(recs):
        K[i] = r["K"]
        for k in arr:
            arr[k][i, :r["K"]] = r[k]
    return K, arr


def golden_losses(loss_fns) -> dict:
    out = {}
    K, arr = parse_debug_log()
    out["log_K"] = K
    for k, v in arr.items():
        out[f"log_{k}"] = v
    # pin the oracle (and the lifted reference) against the logged matched costs
    worst = 0.0
    for i in range(len(K)):
        k = int(K[i])
        mu = torch.tensor(arr["mu_p"][i:i + 1, :], dtype=torch.float32)
        ka = torch.tensor(arr["kappa_p"][i:i + 1, :], dtype=torch.float32)
        w = torch.tensor(arr["w_p"][i:i + 1, :], dtype=torch.float32)
        gt = torch.zeros(1, 4, 3)
        gt[0, :, 0] = torch.tensor(arr["mu_g"][i], dtype=torch.float32)
        gt[0, :, 1] = torch.tensor(arr["kappa_g"][i], dtype=torch.float32)
        _, perm = olosses.match_loss(mu, ka, w, gt, torch.tensor([k]), return_perm=True)
        cost = olosses.kl_von_mises_multi(mu[0, :k], ka[0, :k], gt[0, perm[0, :k], 0], gt[0, perm[0, :k], 1])
        want = torch.tensor(arr["cost"][i, :k], dtype=torch.float32)
        # printed inputs carry ~8 significant digits: compare to the print precision of the inputs
        err = float(((cost - want).abs() / want.abs().clamp_min(1e-2)).max())
        worst = max(worst, err)
    print(f"  debug_log: {len(K)} records, worst relative deviation of oracle matched costs {worst:.2e}")
    assert worst < 5e-4, "loss oracle does not reproduce the reference's logged matched costs"

    g = torch.Generator().manual_seed(123)
    n = 256
    mu_p = ((torch.rand(n, generator=g) * 2 - 1) * math.pi)
    mu_q = ((torch.rand(n, generator=g) * 2 - 1) * math.pi)
    kp = torch.exp(torch.rand(n, generator=g) * 9 - 5)          # 0.0067 .. 55
    kq = torch.where(torch.rand(n, generator=g) < 0.5, torch.full((n,), 8.0), torch.zeros(n))
    kp[:6] = torch.tensor([0.0, 1e-7, 1e-6, 80.0, 30.0, 3.0])
    mu_p[6:10] = torch.tensor([math.pi, -math.pi, 3.1, -3.1]); mu_q[6:10] = torch.tensor([-math.pi, math.pi, -3.1, 3.1])
    for name, fn, ofn in (("single", loss_fns["single"], olosses.kl_von_mises_single),
                          ("multi", loss_fns["multi"], olosses.kl_von_mises_multi)):
        a, b_ = mu_p.clone().requires_grad_(True), kp.clone().requires_grad_(True)
        v = fn(a, b_, mu_q, kq)
        v.sum().backward()
        a64, b64 = mu_p.double().requires_grad_(True), kp.double().requires_grad_(True)
        v64 = ofn(a64, b64, mu_q.double(), kq.double())
        v64.sum().backward()
        a32, b32 = mu_p.clone().requires_grad_(True), kp.clone().requires_grad_(True)
        v32 = ofn(a32, b32, mu_q, kq)
        assert torch.allclose(v32, v, rtol=1e-5, atol=1e-6, equal_nan=True), f"{name} KL oracle != reference"
        out[f"kl_{name}_val"], out[f"kl_{name}_dmu"], out[f"kl_{name}_dk"] = v.detach().numpy(), a.grad.numpy(), b_.grad.numpy()
        out[f"kl_{name}_val64"], out[f"kl_{name}_dmu64"], out[f"kl_{name}_dk64"] = v64.detach().numpy(), a64.grad.numpy(), b64.grad.numpy()
    out.update(kl_mu_p=mu_p.numpy(), kl_kappa_p=kp.numpy(), kl_mu_q=mu_q.numpy(), kl_kappa_q=kq.numpy())

    # match_loss: reference (scipy Hungarian) vs oracle (permutation arg-min)
    Bm = 96
    mu = ((torch.rand(Bm, 4, generator=g) * 2 - 1) * math.pi)
    ka = torch.exp(torch.rand(Bm, 4, generator=g) * 6 - 3).clamp_max(80.0)
    w = F.softmax(torch.randn(Bm, 4, generator=g), -1)
    Kg = torch.tensor([0, 1, 2, 3, 4, 1, 2, 4] * (Bm // 8))
    gt = torch.zeros(Bm, 4, 3)
    gt[..., 0] = (torch.rand(Bm, 4, generator=g) * 2 - 1) * math.pi
    gt[..., 1] = torch.where(torch.rand(Bm, 4, generator=g) < 0.8, torch.full((Bm, 4), 8.0), torch.zeros(Bm, 4))
    gt[..., 2] = 0.25
    ka[5, 0] = 300.0                                           # I0 overflows in fp32 -> nan_to_num(1e6) path
    a, b_, c_ = mu.clone().requires_grad_(True), ka.clone().requires_grad_(True), w.clone().requires_grad_(True)
    lv = loss_fns["match"](a, b_, c_, gt, gt, Kg)
    lv.sum().backward()
    oa, ob, oc = mu.clone().requires_grad_(True), ka.clone().requires_grad_(True), w.clone().requires_grad_(True)
    olv, operm = olosses.match_loss(oa, ob, oc, gt, Kg, return_perm=True)
    olv.sum().backward()
    assert torch.allclose(olv, lv, rtol=1e-5, atol=1e-6), "match_loss oracle != reference"
    for nm, x, y in (("dmu", oa.grad, a.grad), ("dkappa", ob.grad, b_.grad), ("dw", oc.grad, c_.grad)):
        bad = ~torch.isclose(x, y, rtol=1e-4, atol=1e-5, equal_nan=True)
        if bad.any():
            i = bad.nonzero()[0]
            print("   mismatch", nm, i.tolist(), float(x[tuple(i)]), float(y[tuple(i)]), "K", int(Kg[i[0]]), "kappa", ka[i[0]].tolist())
        assert not bad.any(), f"match_loss oracle grad {nm} != reference"
    out.update(m_mu=mu.numpy(), m_kappa=ka.numpy(), m_w=w.numpy(), m_gt=gt.numpy(), m_K=Kg.numpy(), m_loss=lv.detach().numpy(),
               m_dmu=a.grad.numpy(), m_dkappa=b_.grad.numpy(), m_dw=c_.grad.numpy(), m_perm=operm.numpy())

    logits = torch.randn(64, 8, generator=g) * 3
    p = F.softmax(torch.randn(64, 8, generator=g), -1)
    p[:8] = 0.125
    lg = logits.clone().requires_grad_(True)
    cv = loss_fns["ce"](lg, p)
    cv.sum().backward()
    assert torch.allclose(olosses.soft_ce(logits, p), cv, rtol=1e-6, atol=1e-7), "soft CE oracle != reference"
    out.update(ce_logits=logits.numpy(), ce_p=p.numpy(), ce_loss=cv.detach().numpy(), ce_dlogits=lg.grad.numpy())
    print("  losses: oracle == lifted reference functions (values and autograd gradients)")
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = load_reference()
    from scipy.optimize import linear_sum_assignment
    dev = torch.device("cpu")
    single = lift("train_single_peak_vonMises_KL.py", ["kl_von_mises"], {})
    multi = lift("train_multi_peaks_vonMises_KL.py", ["kl_von_mises", "match_loss"],
                 dict(device=dev, linear_sum_assignment=linear_sum_assignment))
    ce = lift("train_8dir_KL.py", ["kl_loss_per_sample_from_logits"], {})
    loss_fns = dict(single=single["kl_von_mises"], multi=multi["kl_von_mises"], match=multi["match_loss"],
                    ce=ce["kl_loss_per_sample_from_logits"])
    print("sampling / grouping"); np.savez_compressed(os.path.join(OUT, "sampling.npz"), **golden_sampling(ref))
    print("set abstraction");     np.savez_compressed(os.path.join(OUT, "sa.npz"), **golden_sa(ref))
    print("models");              np.savez_compressed(os.path.join(OUT, "models.npz"), **golden_models(ref, loss_fns))
    print("losses");              np.savez_compressed(os.path.join(OUT, "losses.npz"), **golden_losses(loss_fns))
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
