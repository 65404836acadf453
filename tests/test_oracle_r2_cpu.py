"""Round-2 oracle pieces (square_distance, SSG / MSG set abstraction, classifier, vanilla PointNet) against the golden
vectors recorded from the unmodified reference (oracle/make_golden_r2.py), plus host-side checks of the new modules
(state_dict layout, argument errors) that need no GPU."""
import numpy as np
import pytest
import torch

from oracle import pointnet_torch, sa_torch, sampling


def test_square_distance_oracle_matches_reference(golden):
    g = golden("ssg_msg")
    assert np.allclose(sampling.square_distance(g["sqd_src"], g["sqd_dst"]), g["sqd_out"], rtol=1e-5, atol=2e-6)
    z = sampling.square_distance(np.zeros((1, 2, 3), np.float32), np.zeros((1, 0, 3), np.float32))
    assert z.shape == (1, 2, 0)                                             # empty destination set


def test_msg_oracle_matches_reference_branches(pcoe, golden):
    g = golden("ssg_msg")
    radii, ks, mlps = [0.15, 0.3, 0.6], [16, 32, 64], [[16, 16, 32], [32, 32, 64], [32, 48, 64]]
    torch.manual_seed(99)
    layer = pcoe.PointNetSetAbstractionMsg(40, radii, ks, 0, mlps)        # same init as the reference's three branches
    sd = sa_torch.clone_state({f"m.{k}": v for k, v in layer.state_dict().items()})
    xyz, fps = torch.from_numpy(g["msg_xyz"]), torch.from_numpy(g["msg_fps"])
    new_xyz = sa_torch.gather(xyz, fps)
    for r, k in zip(radii, ks):
        assert np.array_equal(sampling.ball_query(r, k, xyz.numpy(), new_xyz.numpy()), g[f"msg_group_{k}"])
    _, y = sa_torch.msg_layer(sd, "m", xyz, None, fps, radii, ks)
    assert torch.allclose(y.transpose(1, 2), torch.from_numpy(g["msg_out"]), rtol=1e-4, atol=1e-4)


def test_ssg_cls_oracle_matches_reference_run(pcoe, golden):
    g = golden("ssg_msg")
    torch.manual_seed(2718)
    model = pcoe.PointNetPlusPlusCls(num_classes=40, normal_channel=True)
    sd = sa_torch.clone_state(model.state_dict())
    logp = sa_torch.ssg_cls_forward(sd, torch.from_numpy(g["cls_x"]), torch.from_numpy(g["cls_fps1"]), torch.from_numpy(g["cls_fps2"]))
    assert torch.allclose(logp, torch.from_numpy(g["cls_logp"]), rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("tag,ft", [("ft", True), ("noft", False)])
def test_pointnet_oracle_matches_reference_run(pcoe, golden, tag, ft):
    g = golden("pointnet")
    torch.manual_seed(1234)
    model = pcoe.PointNet(feature_transform=ft)
    sd = model.state_dict()
    for k in list(sd):
        if f"{tag}_sd.{k}" in g.files:
            sd[k] = torch.from_numpy(g[f"{tag}_sd.{k}"])
    y, gfeat, trans, tf = pointnet_torch.pointnet_forward(sa_torch.clone_state(sd), torch.from_numpy(g[f"{tag}_x"]), feature_transform=ft)
    assert torch.allclose(y, torch.from_numpy(g[f"{tag}_y"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(gfeat, torch.from_numpy(g[f"{tag}_gfeat"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(trans, torch.from_numpy(g[f"{tag}_trans"]), rtol=1e-4, atol=1e-5)
    if ft:
        assert torch.allclose(tf, torch.from_numpy(g[f"{tag}_trans_feat"]), rtol=1e-4, atol=1e-4)


def test_new_modules_state_dict_layout_and_errors(pcoe):
    m = pcoe.PointNet(feature_transform=True)
    keys = list(m.state_dict())
    assert keys[0] == "encoder.stn.conv1.weight" and "encoder.fstn.fc3.bias" in keys and keys[-1] == "fc3.bias"
    assert m.encoder.fstn.fc3.weight.shape == (64 * 64, 256)
    cls = pcoe.PointNetPlusPlusCls()
    assert [k for k in cls.state_dict() if k.startswith("sa1.")][0] == "sa1.mlp_convs.0.weight"
    assert cls.sa1.mlp_convs[0].weight.shape == (64, 6, 1, 1)
    with pytest.raises(ValueError):
        pcoe.PointNetSetAbstractionMsg(16, [0.1, 0.2], [16], 0, [[16, 16, 32]])
    with pytest.raises(NotImplementedError):
        pcoe.STNkd(48)
    with pytest.raises(RuntimeError):                                   # no CPU path
        pcoe.PointNet().eval()(torch.zeros(1, 32, 3))
    with pytest.raises(RuntimeError):
        pcoe.SimpleSetAbstraction(8, 0.2, 16, 0, [16, 16, 32])(torch.zeros(1, 3, 32), None)
