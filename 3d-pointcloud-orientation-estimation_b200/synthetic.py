"""Synthetic ModelNet40-shaped inputs and targets shared by bench.py, smoke() and tests."""
from __future__ import annotations

import math

import torch


def clouds(cfg_id: int, B: int, N: int, batch: int = 0) -> torch.Tensor:
    """Centred, unit-ball, duplicate-free clouds (B,N,3) fp32 on the CPU."""
    g = torch.Generator("cpu").manual_seed(1234 + cfg_id + 1000 * batch)
    x = torch.randn(B, N, 3, generator=g)
    x = x - x.mean(1, keepdim=True)
    return (x / x.norm(dim=-1).amax(1).view(B, 1, 1)).contiguous()


def vm_targets(B: int, seed: int = 0):
    """single-peak: mu ~ U(-pi,pi), kappa in {8,0} w.p. 1/2 (data_process/2d_single_peak_vM_gt.py:6-8,43-46)."""
    g = torch.Generator("cpu").manual_seed(77 + seed)
    mu = (torch.rand(B, generator=g) * 2 - 1) * math.pi
    kappa = torch.where(torch.rand(B, generator=g) < 0.5, torch.full((B,), 8.0), torch.zeros(B))
    return mu, kappa


def mvm_targets(B: int, seed: int = 0):
    """multi-peak: K in {1,2,4} w.p. {0.82,0.07,0.11} (debug-log frequencies), peaks at yaw + 2*pi*j/K
    wrapped to (-pi,pi], kappa 8, weights 1/K, zero padded to 4 (2d_multi_peak_MvM_gt_1.py:27,66-72;
    dataloader_multi_peak_vonMises.py:59-64).  Returns vm_gt (B,4,3), K_gt (B,) int64."""
    g = torch.Generator("cpu").manual_seed(99 + seed)
    u = torch.rand(B, generator=g)
    K = torch.where(u < 0.82, 1, torch.where(u < 0.89, 2, 4))
    yaw = (torch.rand(B, generator=g) * 2 - 1) * math.pi
    j = torch.arange(4).view(1, 4)
    ang = yaw.view(B, 1) + 2 * math.pi * j / K.view(B, 1)
    ang = torch.remainder(ang + math.pi, 2 * math.pi) - math.pi
    valid = (j < K.view(B, 1)).float()
    gt = torch.stack([ang * valid, 8.0 * valid, valid / K.view(B, 1)], dim=-1)
    return gt.contiguous(), K


def dir8_targets(B: int, dirs8: torch.Tensor, seed: int = 0):
    """8-direction soft labels: relu(DIRS_8 . v) normalised for a random unit yaw v, uniform 0.125 for
    half of the clouds (data_process/2d_8dir_sample.py:29-39)."""
    g = torch.Generator("cpu").manual_seed(55 + seed)
    th = (torch.rand(B, generator=g) * 2 - 1) * math.pi
    v = torch.stack([torch.sin(th), torch.zeros(B), -torch.cos(th)], -1)
    p = torch.relu(v @ dirs8.t())
    p = p / p.sum(1, keepdim=True)
    uni = torch.rand(B, generator=g) < 0.5
    p[uni] = 0.125
    return p.contiguous()
