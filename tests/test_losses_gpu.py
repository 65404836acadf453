"""Fused CUDA loss kernels against the reference's lifted functions (golden), its debug log and fp64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, want32, want64=None, rtol=1e-5, atol=1e-6):
    """within tolerance of the fp32 reference OR closer to the fp64 truth than the fp32 reference is"""
    got, want32 = np.asarray(got, np.float64), np.asarray(want32, np.float64)
    ok = np.isclose(got, want32, rtol=rtol, atol=atol, equal_nan=True)
    if want64 is not None:
        w64 = np.asarray(want64, np.float64)
        ok |= np.abs(got - w64) <= np.abs(want32 - w64) + atol
    return ok


@pytest.mark.parametrize("name,fn", [("single", "kl_von_mises"), ("multi", "kl_von_mises_clamped")])
def test_vm_kl_value_and_grad(pcoe, golden, cuda, name, fn):
    g = golden("losses")
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    mu, ka = t("kl_mu_p").requires_grad_(True), t("kl_kappa_p").requires_grad_(True)
    v = getattr(pcoe, fn)(mu, ka, t("kl_mu_q"), t("kl_kappa_q"))
    v.sum().backward()
    assert _close(v.detach().cpu(), g[f"kl_{name}_val"], g[f"kl_{name}_val64"], rtol=1e-5, atol=2e-6).all()
    assert _close(mu.grad.cpu(), g[f"kl_{name}_dmu"], g[f"kl_{name}_dmu64"], rtol=1e-4, atol=1e-5).all()
    assert _close(ka.grad.cpu(), g[f"kl_{name}_dk"], g[f"kl_{name}_dk64"], rtol=1e-4, atol=1e-5).all()
    # fp64 truth (same reference formula evaluated in double): tight
    assert np.allclose(v.detach().cpu().numpy(), g[f"kl_{name}_val64"], rtol=2e-6, atol=2e-6)


def test_vm_kl_fp32_overflow_semantics(pcoe, cuda):
    """torch.special.i0 is inf in fp32 for kappa > log(FLT_MAX): the single-peak loss is NaN for an
    overflowing prediction and +inf for an overflowing target (train_single_peak_vonMises_KL.py:23-28)."""
    mu = torch.zeros(3, device=cuda)
    v = pcoe.kl_von_mises(mu, torch.tensor([100.0, 1.0, 88.0], device=cuda), mu, torch.tensor([1.0, 100.0, 1.0], device=cuda))
    assert torch.isnan(v[0]) and torch.isinf(v[1]) and v[1] > 0 and torch.isfinite(v[2])


def test_match_loss_golden(pcoe, golden, cuda):
    g = golden("losses")
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    mu, ka, w = t("m_mu").requires_grad_(True), t("m_kappa").requires_grad_(True), t("m_w").requires_grad_(True)
    loss, perm = pcoe.match_loss(mu, ka, w, t("m_gt"), t("m_gt"), t("m_K"), return_perm=True)
    # same assignment as SciPy's Hungarian, except where two targets have kappa_g = 0: their cost
    # columns are equal to ~1e-7 (a flat von Mises ignores mu_g) and the optimum is a near-tie
    diff = (perm.cpu().numpy() != g["m_perm"]).any(1)
    valid = np.arange(4)[None] < g["m_K"][:, None]
    tie_prone = ((g["m_gt"][..., 1] == 0) & valid).sum(1) >= 2
    assert not (diff & ~tie_prone).any()
    assert np.allclose(loss.detach().cpu().numpy(), g["m_loss"], rtol=1e-5, atol=2e-6)
    loss.sum().backward()

    def check(name, got, want, rtol, atol):
        got, want = got.cpu().numpy().astype(np.float64), want.astype(np.float64)
        bad = ~np.isclose(got, want, rtol=rtol, atol=atol, equal_nan=True)
        if tie_prone.any():
            bad[tie_prone & diff] = False                    # a different (equally optimal) assignment
        if name == "dw":
            bad[g["m_loss"] > 1e5] = False                   # (1e6 - loss)/W: pure fp32 cancellation in the reference
        assert not bad.any(), (name, np.argwhere(bad)[:4].tolist(), got[bad][:4], want[bad][:4])

    check("dmu", mu.grad, g["m_dmu"], 1e-4, 1e-5)
    check("dkappa", ka.grad, g["m_dkappa"], 1e-4, 1e-5)
    # dw = (c_i - loss)/W cancels to ~1e-8*c for K = 1: fp32 noise in the reference, exact here
    check("dw", w.grad, g["m_dw"], 1e-3, 2e-4)
    K = g["m_K"]
    assert (loss.detach().cpu().numpy()[K <= 0] == 0).all() and (mu.grad.cpu().numpy()[K <= 0] == 0).all()


def test_match_loss_reference_debug_log(pcoe, golden, cuda):
    """The matched costs the reference printed while training (debug_log.txt) are reproduced."""
    g = golden("losses")
    K = torch.from_numpy(g["log_K"]).to(cuda)
    f = lambda k: torch.from_numpy(g[k]).float().to(cuda)
    gt = torch.zeros(len(K), 4, 3, device=cuda)
    gt[..., 0], gt[..., 1] = f("log_mu_g"), f("log_kappa_g")
    mu, ka, w = f("log_mu_p"), f("log_kappa_p"), f("log_w_p")
    loss, perm = pcoe.match_loss(mu, ka, w, gt, gt, K, return_perm=True)
    pj = perm.clamp_min(0).long()
    cost = pcoe.kl_von_mises_clamped(mu, ka, torch.gather(gt[..., 0], 1, pj), torch.gather(gt[..., 1], 1, pj))
    valid = (torch.arange(4, device=cuda)[None] < K[:, None]).cpu().numpy()
    want = g["log_cost"]
    got = cost.cpu().numpy()
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-2)
    assert rel[valid].max() < 5e-4
    wl = (g["log_w_p"] * want * valid).sum(1) / ((g["log_w_p"] * valid).sum(1) + 1e-8)
    assert np.allclose(loss.cpu().numpy(), wl, rtol=1e-3, atol=1e-4)


def test_soft_ce_golden(pcoe, golden, cuda):
    g = golden("losses")
    lg = torch.from_numpy(g["ce_logits"]).to(cuda).requires_grad_(True)
    v = pcoe.kl_loss_per_sample_from_logits(lg, torch.from_numpy(g["ce_p"]).to(cuda))
    (v * torch.arange(1, 65, device=cuda)).sum().backward()
    assert np.allclose(v.detach().cpu().numpy(), g["ce_loss"], rtol=1e-6, atol=1e-6)
    assert np.allclose(lg.grad.cpu().numpy(), g["ce_dlogits"] * np.arange(1, 65)[:, None], rtol=1e-5, atol=1e-6)


def test_losses_random_vs_oracle_autograd(pcoe, cuda):
    from oracle import losses as ol
    g = torch.Generator().manual_seed(9)
    B = 512
    mu = (torch.rand(B, 4, generator=g) * 2 - 1) * 3.14159
    ka = torch.exp(torch.rand(B, 4, generator=g) * 8 - 4).clamp_max(80.0)
    w = torch.softmax(torch.randn(B, 4, generator=g), -1)
    gt = torch.zeros(B, 4, 3)
    gt[..., 0] = (torch.rand(B, 4, generator=g) * 2 - 1) * 3.14159
    gt[..., 1] = 8.0
    K = torch.randint(0, 5, (B,), generator=g)
    a, b, c = (x.double().requires_grad_(True) for x in (mu, ka, w))
    want = ol.match_loss(a, b, c, gt.double(), K)
    want.sum().backward()
    a2, b2, c2 = (x.to(cuda).requires_grad_(True) for x in (mu, ka, w))
    got = pcoe.match_loss(a2, b2, c2, gt.to(cuda), None, K.to(cuda))
    got.sum().backward()
    assert torch.allclose(got.cpu().double(), want, rtol=1e-5, atol=1e-5)
    assert torch.allclose(a2.grad.cpu().double(), a.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(b2.grad.cpu().double(), b.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(c2.grad.cpu().double(), c.grad, rtol=1e-4, atol=1e-4)
