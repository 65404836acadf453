"""Vanilla PointNet inference path (models/pointnet.py:6-129, eval mode) through libpcoe, against outputs recorded from
the unmodified reference (oracle/make_golden_r2.py) and the torch-CPU oracle at BASELINE configs[4] shapes."""
import numpy as np
import pytest
import torch

from oracle import pointnet_torch, sa_torch

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-3          # relative to the output scale (north star: <= 1e-3); measured ~1e-5


def _load_case(pcoe, g, tag, ft):
    torch.manual_seed(1234)
    model = pcoe.PointNet(feature_transform=ft)
    sd = model.state_dict()
    for k in list(sd):
        if f"{tag}_sd.{k}" in g.files:
            sd[k] = torch.from_numpy(g[f"{tag}_sd.{k}"])
    model.load_state_dict(sd, strict=True)
    chk = float(sum(v.double().abs().sum() for v in model.state_dict().values() if v.is_floating_point()))
    assert abs(chk - float(g[f"{tag}_checksum"])) < 1e-6 * float(g[f"{tag}_checksum"])      # same checkpoint as the reference run
    return model


def _close(a, b, tol=OUT_TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("tag,ft", [("ft", True), ("noft", False)])
def test_pointnet_eval_matches_reference_run(pcoe, golden, cuda, tag, ft):
    g = golden("pointnet")
    model = _load_case(pcoe, g, tag, ft).to(cuda).eval()
    x = torch.from_numpy(g[f"{tag}_x"]).to(cuda)            # (B,N,3); N=200 is not a multiple of 32 (padding path)
    with torch.no_grad():
        y = model(x)
        gfeat, trans, trans_feat = model.encoder(x.transpose(1, 2))
    assert _close(trans.cpu(), g[f"{tag}_trans"])
    assert _close(gfeat.cpu(), g[f"{tag}_gfeat"])
    assert _close(y.cpu(), g[f"{tag}_y"])
    if ft:
        assert _close(trans_feat.cpu(), g[f"{tag}_trans_feat"])
    else:
        assert trans_feat is None
    # channel-first input takes the same path (models/pointnet.py:124-125)
    with torch.no_grad():
        assert torch.allclose(model(x.transpose(1, 2).contiguous()), y, atol=1e-6)


def test_stn_modules_standalone(pcoe, cuda):
    torch.manual_seed(5)
    stn3, stnk = pcoe.STN3d(3).to(cuda).eval(), pcoe.STNkd(64).to(cuda).eval()
    x3 = torch.randn(3, 3, 96, device=cuda)
    xk = torch.randn(3, 64, 96, device=cuda)
    with torch.no_grad():
        t3, tk = stn3(x3), stnk(xk)
    o3 = pointnet_torch.stn(sa_torch.clone_state({f"s.{k}": v for k, v in stn3.state_dict().items()}), "s", x3.cpu(), 3)
    ok = pointnet_torch.stn(sa_torch.clone_state({f"s.{k}": v for k, v in stnk.state_dict().items()}), "s", xk.cpu(), 64)
    assert t3.shape == (3, 3, 3) and tk.shape == (3, 64, 64)
    assert _close(t3.cpu(), o3) and _close(tk.cpu(), ok)


@pytest.mark.parametrize("B", [1, 7, 64])
def test_pointnet_baseline_shape_vs_oracle(pcoe, cuda, B):
    """BASELINE configs[4]: B x 1024 points, eval; oracle = torch-CPU restatement on the same checkpoint."""
    torch.manual_seed(77)
    model = pcoe.PointNet(feature_transform=True)
    g = torch.Generator().manual_seed(B)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.2, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(-1.0, 1.5, generator=g)
                m.bias.uniform_(-0.3, 0.3, generator=g)
    sd = sa_torch.clone_state(model.state_dict(), dtype=torch.float64)
    model = model.to(cuda).eval()
    x = pcoe.synthetic.clouds(3, B, 1024)
    with torch.no_grad():
        y = model(x.to(cuda))
    oy, _, _, _ = pointnet_torch.pointnet_forward(sd, x.double(), feature_transform=True)
    assert y.shape == (B, 3)
    assert _close(y.cpu(), oy)
    # size-independent property: the global max-pool makes the output invariant to a permutation of the points
    perm = torch.randperm(1024, generator=g)
    with torch.no_grad():
        y2 = model(x[:, perm].to(cuda))
    assert _close(y2.cpu(), y.cpu(), 1e-4)


def test_pointnet_rejects_train_mode_and_cpu(pcoe, cuda):
    model = pcoe.PointNet().to(cuda)
    with pytest.raises(NotImplementedError):
        model(torch.zeros(2, 64, 3, device=cuda))
    model.eval()
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 64, 3))
