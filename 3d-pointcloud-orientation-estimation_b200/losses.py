"""Loss callables with the reference's names and signatures, backed by the fused CUDA kernels.

    kl_von_mises(mu_p, kappa_p, mu_q, kappa_q) -> (B,)        train_single_peak_vonMises_KL.py:23-28
    kl_von_mises_clamped(...)                  -> (B,)        train_multi_peaks_vonMises_KL.py:38-52
    match_loss(mu, kappa, w, vm_gt, _, K_gt)   -> (B,)        train_multi_peaks_vonMises_KL.py:54-81
    kl_loss_per_sample_from_logits(logits, p)  -> (B,)        train_8dir_KL.py:60-68

Each is a torch.autograd.Function: the kernel computes the value and d loss_b / d input in one
launch; backward only scales by the upstream gradient.  No host synchronisation (the reference's
match_loss does two D2H syncs per sample and runs SciPy's Hungarian solver on the host).
"""
from __future__ import annotations

import torch

from . import _lib, ops


class _VmKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu_p, kappa_p, mu_q, kappa_q, variant):
        loss, dmu, dk = ops.vm_kl_fwd_bwd(mu_p, kappa_p, mu_q, kappa_q, variant)
        ctx.save_for_backward(dmu, dk)
        ctx.shape = mu_p.shape
        return loss.view(mu_p.shape)

    @staticmethod
    def backward(ctx, g):
        dmu, dk = ctx.saved_tensors
        g = g.contiguous().view(-1)
        return (g * dmu.view(-1)).view(ctx.shape), (g * dk.view(-1)).view(ctx.shape), None, None, None


def kl_von_mises(mu_p, kappa_p, mu_q, kappa_q):
    """Single-peak closed-form KL(vM_p || vM_q); gradients flow to (mu_p, kappa_p) only."""
    return _VmKL.apply(mu_p, kappa_p, mu_q, kappa_q, _lib.VM_SINGLE)


def kl_von_mises_clamped(mu_p, kappa_p, mu_q, kappa_q):
    """Multi-peak variant: kappa clamped to [1e-6, 500], delta wrapped to [-pi, pi)."""
    return _VmKL.apply(mu_p, kappa_p, mu_q, kappa_q, _lib.VM_MULTI)


class _MatchLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, kappa, w, vm_gt, K_gt):
        loss, d3, perm = ops.mvm_match_fwd_bwd(mu, kappa, w, vm_gt, K_gt)
        ctx.save_for_backward(d3)
        ctx.mark_non_differentiable(perm)
        return loss, perm

    @staticmethod
    def backward(ctx, g, _gperm):
        (d3,) = ctx.saved_tensors              # (dmu, dk, dw) stacked [3,B,K]: one multiply instead of three
        gd = g.contiguous().view(1, -1, 1) * d3
        return gd[0], gd[1], gd[2], None, None


def match_loss(mu_pred, kappa_pred, w_pred, vm_gt, _, K_gt, return_perm: bool = False):
    """Per-sample matched mixture KL (B,).  The fifth argument is ignored, as in the reference."""
    loss, perm = _MatchLoss.apply(mu_pred, kappa_pred, w_pred, vm_gt, K_gt)
    return (loss, perm) if return_perm else loss


class _SoftCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, p):
        loss, dl = ops.soft_ce_fwd_bwd(logits, p)
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return g.contiguous().view(-1, 1) * dl, None


def kl_loss_per_sample_from_logits(logits, p_target):
    """-(p_target * log_softmax(logits)).sum(1); gradient flows to logits only."""
    return _SoftCE.apply(logits, p_target)
