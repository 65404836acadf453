#!/usr/bin/env python
"""Round-2 golden vectors: square_distance, FPS + ball-query set abstraction (SSG), its classifier, multi-scale
grouping (MSG) and the vanilla PointNet inference path.

Run in the build container (needs /root/reference):   python oracle/make_golden_r2.py
Every oracle function is first asserted against the unmodified reference function it restates, then the reference's
outputs are written to tests/golden/{ssg_msg,pointnet}.npz.  Parameters are NOT stored: every model is constructed
under torch.manual_seed(s) and the drop-in modules reproduce the reference's default initialisation bit for bit
(checked here through a checksum of the state dict, and in the tests).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PCOE_REF", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import pointnet_torch, sa_torch, sampling  # noqa: E402


def load_demo():
    spec = importlib.util.spec_from_file_location("pp_demo", os.path.join(REF, "PointNet++Demo.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def unit_clouds(seed, B, N):
    g = torch.Generator("cpu").manual_seed(seed)
    x = torch.randn(B, N, 3, generator=g)
    x = x - x.mean(1, keepdim=True)
    return (x / x.norm(dim=-1).amax(1).view(B, 1, 1)).contiguous()


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


def fixed_fps(demo, xyz, S, start):
    """the reference's FPS with its torch.randint start draw (PointNet++Demo.py:20) pinned to `start`"""
    real = torch.randint
    torch.randint = lambda *a, **k: start.clone()
    try:
        return demo.farthest_point_sample(xyz, S)
    finally:
        torch.randint = real


def golden_ssg_msg(demo) -> dict:
    sys.path.insert(0, REF)
    from models import base
    out = {}
    # ---- square_distance (models/base.py:20-27)
    g = torch.Generator().manual_seed(9)
    src, dst = torch.randn(2, 37, 3, generator=g), torch.randn(2, 301, 3, generator=g)
    ref_d = base.square_distance(src, dst)
    assert np.allclose(sampling.square_distance(src.numpy(), dst.numpy()), ref_d.numpy(), rtol=1e-5, atol=2e-6)
    out.update(sqd_src=src.numpy(), sqd_dst=dst.numpy(), sqd_out=ref_d.numpy())

    # ---- one SimpleSetAbstraction layer with features, train-mode forward + backward (PointNet++Demo.py:74-129)
    B, N, D, S = 2, 300, 6, 48
    xyz = unit_clouds(21, B, N)
    pts = torch.randn(B, N, D, generator=g) * 0.5
    start = torch.tensor([5, 123])
    torch.manual_seed(314)
    layer = demo.SimpleSetAbstraction(S, 0.35, 32, D, [32, 48, 64])
    with torch.no_grad():                       # mixed-sign BatchNorm weights exercise the max/min pooling path
        for bn in layer.mlp_bns:
            bn.weight.uniform_(-1.0, 1.5, generator=g)
            bn.bias.uniform_(-0.3, 0.3, generator=g)
    sd0 = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    fps = fixed_fps(demo, xyz, S, start)
    real_fps = demo.farthest_point_sample
    demo.farthest_point_sample = lambda x, n: fps
    try:
        layer.train()
        nx, y = layer(xyz.transpose(1, 2).contiguous(), pts.transpose(1, 2).contiguous())
    finally:
        demo.farthest_point_sample = real_fps
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    osd = sa_torch.clone_state({f"l.{k}": v for k, v in sd0.items()}, requires_grad=True)
    onx, oy, ogi = sa_torch.ssg_layer(osd, "l", xyz, pts, fps, 0.35, 32)
    assert torch.allclose(oy.transpose(1, 2), y, rtol=1e-4, atol=1e-5), "SSG oracle forward != reference"
    (oy.transpose(1, 2) * gy).sum().backward()
    for name, p in layer.named_parameters():
        og = osd[f"l.{name}"].grad
        if "convs" in name and name.endswith("bias"):
            continue
        rel = float((og - p.grad).norm() / p.grad.norm().clamp_min(1e-20))
        assert rel < 2e-3, f"SSG oracle grad {name}: {rel:.2e}"
        out[f"ssg_grad.{name}"] = p.grad.numpy()
    out.update(ssg_xyz=xyz.numpy(), ssg_pts=pts.numpy(), ssg_start=start.numpy(), ssg_fps=fps.numpy(),
               ssg_group=ogi.numpy(), ssg_out=y.detach().numpy(), ssg_gy=gy.numpy(),
               ssg_bn_w=np.stack([np.pad(sd0[f"mlp_bns.{l}.weight"].numpy(), (0, 64 - sd0[f"mlp_bns.{l}.weight"].numel())) for l in range(3)]),
               ssg_bn_b=np.stack([np.pad(sd0[f"mlp_bns.{l}.bias"].numpy(), (0, 64 - sd0[f"mlp_bns.{l}.bias"].numel())) for l in range(3)]),
               ssg_checksum=np.array(checksum(sd0)))
    for k, v in layer.state_dict().items():
        if "running" in k:
            out[f"ssg_sd1.{k}"] = v.numpy()
    print("  SSG layer: oracle == reference (fwd, bwd)")

    # ---- PointNetPlusPlusCls (PointNet++Demo.py:178-240), train mode, dropout off, 6-channel input
    B, N = 2, 640
    x = torch.cat([unit_clouds(77, B, N), F.normalize(torch.randn(B, N, 3, generator=g), dim=-1)], -1).transpose(1, 2).contiguous()
    labels = torch.tensor([3, 17])
    torch.manual_seed(2718)
    model = demo.PointNetPlusPlusCls(num_classes=40, normal_channel=True)
    model.dropout1.p = model.dropout2.p = 0.0
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    xyz_cl = x[:, :3].transpose(1, 2).contiguous()
    s1, s2 = torch.tensor([11, 500]), torch.tensor([0, 300])
    fps1 = fixed_fps(demo, xyz_cl, 512, s1)
    l1_xyz = demo.index_points(xyz_cl, fps1)
    fps2 = fixed_fps(demo, l1_xyz, 128, s2)
    seq = iter([fps1, fps2])
    demo.farthest_point_sample = lambda x_, n: next(seq)
    try:
        model.train()
        logp = model(x)
    finally:
        demo.farthest_point_sample = real_fps
    loss = F.nll_loss(logp, labels)
    loss.backward()
    osd = sa_torch.clone_state(sd0, requires_grad=True)
    ologp = sa_torch.ssg_cls_forward(osd, x, fps1, fps2)
    assert torch.allclose(ologp, logp, rtol=1e-3, atol=1e-4), "Cls oracle != reference"
    out.update(cls_x=x.numpy(), cls_labels=labels.numpy(), cls_fps1=fps1.numpy(), cls_fps2=fps2.numpy(),
               cls_logp=logp.detach().numpy(), cls_loss=np.array(float(loss)), cls_checksum=np.array(checksum(sd0)))
    for name, p in model.named_parameters():
        out[f"cls_gnorm.{name}"] = np.array(float(p.grad.norm()))
    print(f"  Cls model: oracle == reference, loss {float(loss):.6f}")

    # ---- MSG: three scales on shared centroids; the reference has no MSG class, so the pin is per scale: branch i of the
    # oracle == the reference's SimpleSetAbstraction(radius_i, nsample_i, mlp_i) on the same centroids
    B, N, S = 2, 400, 40
    xyz = unit_clouds(5, B, N)
    start = torch.tensor([7, 77])
    fps = fixed_fps(demo, xyz, S, start)
    radii, nsamples, mlps = [0.15, 0.3, 0.6], [16, 32, 64], [[16, 16, 32], [32, 32, 64], [32, 48, 64]]
    torch.manual_seed(99)
    branches = [demo.SimpleSetAbstraction(S, r, k, 0, m) for r, k, m in zip(radii, nsamples, mlps)]
    sd = {}
    refs = []
    demo.farthest_point_sample = lambda x_, n: fps
    try:
        for i, br in enumerate(branches):
            br.train()
            for k, v in br.state_dict().items():
                kind, rest = k.split(".", 1)
                sd[f"m.{'conv_blocks' if kind == 'mlp_convs' else 'bn_blocks'}.{i}.{rest}"] = v.detach().clone()
            refs.append(br(xyz.transpose(1, 2).contiguous(), None)[1])
    finally:
        demo.farthest_point_sample = real_fps
    ref_cat = torch.cat(refs, dim=1)
    _, ocat = sa_torch.msg_layer(sa_torch.clone_state(sd), "m", xyz, None, fps, radii, nsamples)
    assert torch.allclose(ocat.transpose(1, 2), ref_cat, rtol=1e-4, atol=1e-4), "MSG oracle != reference branches"
    out.update(msg_xyz=xyz.numpy(), msg_fps=fps.numpy(), msg_out=ref_cat.detach().numpy(), msg_checksum=np.array(checksum(sd)))
    for r, k in zip(radii, nsamples):
        out[f"msg_group_{k}"] = demo.query_ball_point(r, k, xyz, demo.index_points(xyz, fps)).numpy()
    print("  MSG: oracle == reference per scale")
    return out


def golden_pointnet() -> dict:
    """models/pointnet.py:6-129 in eval mode (the inference path)."""
    sys.path.insert(0, REF)
    from models.pointnet import PointNet
    out = {}
    g = torch.Generator().manual_seed(4)
    for tag, ft, B, N in (("ft", True, 3, 200), ("noft", False, 2, 128)):
        torch.manual_seed(1234)
        model = PointNet(feature_transform=ft)
        # eval mode with non-trivial running statistics / affine parameters (a trained checkpoint has them)
        with torch.no_grad():
            for m in model.modules():
                if isinstance(m, torch.nn.BatchNorm1d):
                    m.running_mean.normal_(0, 0.2, generator=g)
                    m.running_var.uniform_(0.5, 1.5, generator=g)
                    m.weight.uniform_(-1.0, 1.5, generator=g)
                    m.bias.uniform_(-0.3, 0.3, generator=g)
        model.eval()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        x = unit_clouds(50 + B, B, N)
        with torch.no_grad():
            y = model(x)                              # (B,N,3) is transposed by the model (:124-125)
            gfeat, trans, trans_feat = model.encoder(x.transpose(1, 2))
        oy, ogf, otr, otf = pointnet_torch.pointnet_forward(sa_torch.clone_state(sd), x, feature_transform=ft)
        assert torch.allclose(oy, y, rtol=1e-4, atol=1e-5), "PointNet oracle != reference"
        assert torch.allclose(ogf, gfeat, rtol=1e-4, atol=1e-5) and torch.allclose(otr, trans, rtol=1e-4, atol=1e-5)
        if ft:
            assert torch.allclose(otf, trans_feat, rtol=1e-4, atol=1e-4)
        # the perturbed BatchNorm tensors are part of the fixture (everything else follows from the seed)
        for k, v in sd.items():
            if ".bn" in k or k.startswith("bn"):
                if v.is_floating_point():
                    out[f"{tag}_sd.{k}"] = v.numpy()
        out.update({f"{tag}_x": x.numpy(), f"{tag}_y": y.numpy(), f"{tag}_gfeat": gfeat.numpy(), f"{tag}_trans": trans.numpy(),
                    f"{tag}_checksum": np.array(checksum(sd))})
        if ft:
            out[f"{tag}_trans_feat"] = trans_feat.numpy()
        print(f"  PointNet[{tag}]: oracle == reference (eval)")
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    demo = load_demo()
    which = sys.argv[1:] or ["ssg_msg", "pointnet"]
    if "ssg_msg" in which:
        print("ssg / msg"); np.savez_compressed(os.path.join(OUT, "ssg_msg.npz"), **golden_ssg_msg(demo))
    if "pointnet" in which:
        print("pointnet"); np.savez_compressed(os.path.join(OUT, "pointnet.npz"), **golden_pointnet())
    for f in ("ssg_msg.npz", "pointnet.npz"):
        if os.path.exists(os.path.join(OUT, f)):
            print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
