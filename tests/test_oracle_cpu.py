"""The oracle against the golden vectors recorded from the unmodified reference
(oracle/make_golden.py) and against the reference's own committed known answers (debug_log.txt)."""
import math

import numpy as np
import pytest
import torch

from oracle import losses as olosses, sa_torch, sampling


def test_fps_ball_knn_oracle_matches_reference(golden):
    g = golden("sampling")
    for tag in ("a", "b"):
        xyz = g[f"{tag}_xyz"]
        fps = sampling.farthest_point_sample(xyz, 64, g[f"{tag}_start"])
        assert np.array_equal(fps, g[f"{tag}_fps"])
        new_xyz = np.take_along_axis(xyz, fps[..., None].repeat(3, -1), axis=1)
        for r, ns in ((0.2, 16), (0.4, 32), (0.05, 8)):
            assert np.array_equal(sampling.ball_query(r, ns, xyz, new_xyz), g[f"{tag}_ball_{r}_{ns}"])
        idx, margin = sampling.knn(new_xyz, xyz, 32)
        n, eq, tie, bad = sampling.knn_rows_match(g[f"{tag}_knn_sorted"], idx, margin)
        assert bad == 0 and eq >= n - tie


def test_fps_edge_cases():
    xyz = np.zeros((1, 5, 3), np.float32)                       # all points equal: ties -> index 0
    assert sampling.farthest_point_sample(xyz, 4, np.array([3])).tolist() == [[3, 0, 0, 0]]
    xyz = np.array([[[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 2, 0]]], np.float32)
    assert sampling.farthest_point_sample(xyz, 4, np.array([0])).tolist() == [[0, 3, 1, 2]]  # tie (1 vs 2) -> lowest index


def test_ball_query_padding_and_empty_rows():
    xyz = np.array([[[0, 0, 0], [0.1, 0, 0], [5, 5, 5]]], np.float32)
    q = np.array([[[0, 0, 0], [9, 9, 9]]], np.float32)
    out = sampling.ball_query(0.2, 4, xyz, q)
    assert out[0, 0].tolist() == [0, 1, 0, 0]                   # padded with the first hit
    assert out[0, 1].tolist() == [3, 3, 3, 3]                   # no hit: N, as the reference


def test_randperm_replay_matches_recorded_fps_idx(golden):
    g = golden("models")
    fps1 = sampling.randperm_subset_replay(42, 4, 256, 128)
    assert np.array_equal(fps1, g["vonmises_fps1"])


@pytest.mark.parametrize("tag", ["small", "nofeat", "sa2", "gall"])
def test_sa_oracle_matches_reference(golden, tag):
    g = golden("sa")
    B, N, S, K, D, c1, c2, c3, ga = g[f"{tag}_cfg"].tolist()
    sd = {k[len(tag) + 5:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith(f"{tag}_sd0.")}
    sd = sa_torch.clone_state({f"sa.{k}": v for k, v in sd.items()}, requires_grad=True)
    xyz = torch.from_numpy(g[f"{tag}_xyz"])
    pts = torch.from_numpy(g[f"{tag}_pts"]).requires_grad_(True) if D else None
    fps = torch.from_numpy(g[f"{tag}_fps"]) if not ga else None
    grp = torch.from_numpy(g[f"{tag}_grp"]) if not ga else None
    _, y, _ = sa_torch.set_abstraction(sd, "sa", xyz, pts, group_all=bool(ga), nsample=K, fps_idx=fps, group_idx=grp)
    assert torch.allclose(y, torch.from_numpy(g[f"{tag}_out"]), rtol=1e-4, atol=1e-5)
    y.backward(torch.from_numpy(g[f"{tag}_gout"]))
    for k in g.files:
        if k.startswith(f"{tag}_grad.") and not (k.endswith("bias") and ".convs." in k):
            want = torch.from_numpy(g[k])
            got = sd["sa." + k[len(tag) + 6:]].grad
            assert float((got - want).norm() / want.norm().clamp_min(1e-12)) < 2e-3, k
        if k.startswith(f"{tag}_sd1.") and "running" in k:
            assert torch.allclose(sd["sa." + k[len(tag) + 5:]], torch.from_numpy(g[k]), rtol=1e-5, atol=1e-6), k
    if D:
        assert torch.allclose(pts.grad, torch.from_numpy(g[f"{tag}_gpts"]), rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize("kind,cls", [("vonmises", "PointNetPPVonMises"), ("mvm", "PointNetPPMvM"),
                                      ("8dir", "PointNetPP8Dir"), ("xyz", "PointNetPPXYZ")])
def test_model_oracle_and_init_parity(golden, pcoe, kind, cls):
    """Default initialisation of the drop-in modules reproduces the reference's under the same seed
    (same state_dict keys, shapes and values); the oracle then reproduces the recorded outputs."""
    g = golden("models")
    torch.manual_seed(1000)
    model = getattr(pcoe, cls)()
    sd = model.state_dict()
    assert [f"{k}|{'x'.join(map(str, v.shape))}" for k, v in sd.items()] == g[f"{kind}_sd_keys"].tolist()
    if kind == "mvm":
        sd["head_mu.weight"] = torch.from_numpy(g["mvm_head_mu_w"])
        sd["head_pi.weight"] = torch.from_numpy(g["mvm_head_pi_w"])
    chk = float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))
    assert abs(chk - float(g[f"{kind}_sd_checksum"][0])) < 1e-6 * chk
    osd = sa_torch.clone_state(sd)
    xyz = torch.from_numpy(g["xyz"])
    res = sa_torch.model_forward(kind, osd, xyz, torch.from_numpy(g[f"{kind}_fps1"]), torch.from_numpy(g[f"{kind}_fps2"]))
    res = res if isinstance(res, tuple) else (res,)
    for i, r in enumerate(res):
        assert torch.allclose(r, torch.from_numpy(g[f"{kind}_out{i}"]), rtol=1e-3, atol=1e-4)


def test_loss_oracle_reproduces_reference_debug_log(golden):
    """Known answers committed by the reference: results/multi_peak_vonMises_KL_debug/debug_log.txt."""
    g = golden("losses")
    K = g["log_K"]
    assert set(K.tolist()) == {1, 2, 4} and len(K) >= 300
    worst = 0.0
    for i in range(len(K)):
        k = int(K[i])
        f = lambda a: torch.tensor(g[a][i:i + 1], dtype=torch.float32)
        gt = torch.zeros(1, 4, 3)
        gt[0, :, 0], gt[0, :, 1] = f("log_mu_g")[0], f("log_kappa_g")[0]
        _, perm = olosses.match_loss(f("log_mu_p"), f("log_kappa_p"), f("log_w_p"), gt, torch.tensor([k]), return_perm=True)
        cost = olosses.kl_von_mises_multi(f("log_mu_p")[0, :k], f("log_kappa_p")[0, :k], gt[0, perm[0, :k], 0], gt[0, perm[0, :k], 1])
        want = torch.tensor(g["log_cost"][i, :k], dtype=torch.float32)
        worst = max(worst, float(((cost - want).abs() / want.abs().clamp_min(1e-2)).max()))
    assert worst < 5e-4   # inputs are printed with ~8 significant digits


def test_loss_oracle_matches_lifted_reference_functions(golden):
    g = golden("losses")
    t = lambda k: torch.from_numpy(g[k])
    for name, fn in (("single", olosses.kl_von_mises_single), ("multi", olosses.kl_von_mises_multi)):
        mu, ka = t("kl_mu_p").clone().requires_grad_(True), t("kl_kappa_p").clone().requires_grad_(True)
        v = fn(mu, ka, t("kl_mu_q"), t("kl_kappa_q"))
        v.sum().backward()
        assert torch.allclose(v, t(f"kl_{name}_val"), rtol=1e-5, atol=1e-6, equal_nan=True)
        assert torch.allclose(mu.grad, t(f"kl_{name}_dmu"), rtol=1e-4, atol=1e-6, equal_nan=True)
        assert torch.allclose(ka.grad, t(f"kl_{name}_dk"), rtol=1e-4, atol=1e-5, equal_nan=True)
    lv, perm = olosses.match_loss(t("m_mu"), t("m_kappa"), t("m_w"), t("m_gt"), t("m_K"), return_perm=True)
    assert torch.allclose(lv, t("m_loss"), rtol=1e-5, atol=1e-6)
    assert np.array_equal(perm.numpy(), g["m_perm"])
    assert torch.allclose(olosses.soft_ce(t("ce_logits"), t("ce_p")), t("ce_loss"), rtol=1e-6, atol=1e-7)


def test_mu_convention_known_answers():
    """data_process/2d_single_peak_vM_test.ipynb cell 0: yaw of a forward vector, mu = atan2(fx, -fz)."""
    for (fx, fz), want in (((0.0, -1.0), 0.0), ((1.0, 0.0), math.pi / 2), ((-0.749493, -0.662012), -0.847296)):
        assert abs(math.atan2(fx, -fz) - want) < 1e-5
