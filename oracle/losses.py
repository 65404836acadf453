"""torch-CPU oracle of the three losses (autograd gives the gradient oracle).  Any float dtype."""
from __future__ import annotations

import itertools
import math

import torch
import torch.nn.functional as F


def kl_von_mises_single(mu_p, kappa_p, mu_q, kappa_q):
    """Follows kl_von_mises, train_single_peak_vonMises_KL.py:23-28: no clamp, no wrap,
    A(kappa_p) := 0 where kappa_p <= 1e-6."""
    logi0_p = torch.log(torch.special.i0(kappa_p))
    logi0_q = torch.log(torch.special.i0(kappa_q))
    ratio = torch.special.i1(kappa_p) / torch.special.i0(kappa_p)
    a = torch.where(kappa_p <= 1e-6, torch.zeros_like(kappa_p), ratio)
    return logi0_q - logi0_p + kappa_p * a - kappa_q * a * torch.cos(mu_p - mu_q)


def kl_von_mises_multi(mu_p, kappa_p, mu_q, kappa_q):
    """Follows kl_von_mises, train_multi_peaks_vonMises_KL.py:38-52: kappa clamped to [1e-6,500]
    (:39-40), delta wrapped to [-pi,pi) (:48-49), log(I0q/I0p) + A_p (kappa_p - kappa_q cos delta) (:51)."""
    kp = kappa_p.clamp(1e-6, 500.0)
    kq = kappa_q.clamp(1e-6, 500.0)
    i0p, i0q = torch.special.i0(kp), torch.special.i0(kq)
    a = torch.special.i1(kp) / i0p
    delta = torch.remainder(mu_p - mu_q + math.pi, 2 * math.pi) - math.pi
    return torch.log(i0q / i0p) + a * (kp - kq * torch.cos(delta))


def match_loss(mu, kappa, w, vm_gt, K_gt, return_perm: bool = False):
    """Follows match_loss, train_multi_peaks_vonMises_KL.py:54-81.  The host Hungarian step
    (scipy.optimize.linear_sum_assignment, scipy==1.16.2 in the reference's requirements.txt:115, :75)
    is restated as the arg-min over the K! <= 24 permutations of the summed cost, which is the
    published definition of the assignment optimum; gradients flow through cost and w only."""
    B, Kmax = mu.shape
    out = []
    perms = torch.full((B, Kmax), -1, dtype=torch.long)
    for b in range(B):
        K = int(K_gt[b])
        if K <= 0:
            out.append(mu.new_zeros(()))
            continue
        cost = kl_von_mises_multi(mu[b, :K, None], kappa[b, :K, None], vm_gt[b, None, :K, 0], vm_gt[b, None, :K, 1])
        cost = torch.nan_to_num(cost, nan=1e6, posinf=1e6, neginf=1e6)             # :73
        c = cost.detach()
        best = min(itertools.permutations(range(K)), key=lambda p: float(sum(c[i, p[i]] for i in range(K))))
        col = torch.tensor(best)
        perms[b, :K] = col
        ws = w[b, :K]
        out.append((ws * cost[torch.arange(K), col]).sum() / (ws.sum() + 1e-8))    # :77-80
    loss = torch.stack(out)
    return (loss, perms) if return_perm else loss


def soft_ce(logits, p):
    """Follows kl_loss_per_sample_from_logits, train_8dir_KL.py:60-68."""
    return -(p * F.log_softmax(logits, dim=1)).sum(dim=1)
