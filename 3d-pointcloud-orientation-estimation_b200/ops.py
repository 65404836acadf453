"""Tensor-level wrappers of the C ABI (include/pcoe.h): sampling, grouping, losses.

Every function takes CUDA tensors, enqueues on torch's current stream and returns new tensors.
Signatures follow the reference helpers they replace (models/base.py, PointNet++Demo.py).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"pcoe: {name} must be a CUDA tensor (there is no CPU path); got {t.device}")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# sampling
# ------------------------------------------------------------------------------------------------
def farthest_point_sample(xyz: torch.Tensor, npoint: int, start_idx: torch.Tensor | None = None,
                          return_xyz: bool = False, int32: bool = False):
    """FPS indices (B,npoint) int64.  Reference: PointNet++Demo.py:8-29.

    ``start_idx`` (B,) is the first centroid of every cloud; the reference draws it with
    ``torch.randint(0, N, (B,))`` (:20).  ``None`` draws it the same way (torch global generator).
    """
    if xyz.dim() != 3 or xyz.size(-1) != 3:
        raise ValueError(f"xyz must be (B,N,3), got {tuple(xyz.shape)}")
    xyz = _req(xyz, torch.float32, "xyz")
    B, N, _ = xyz.shape
    if start_idx is None:
        start_idx = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    start = _req(start_idx, torch.int32, "start_idx")
    out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    oxyz = torch.empty(B, npoint, 3, dtype=torch.float32, device=xyz.device) if return_xyz else None
    _lib.check(_lib.load().pcoe_fps_f32(xyz.data_ptr(), B, N, npoint, start.data_ptr(), out.data_ptr(),
                                        _ptr(oxyz), _stream()))
    idx = out if int32 else out.long()
    return (idx, oxyz) if return_xyz else idx


def host_randperm_subsets(B: int, N: int, S: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """``torch.stack([torch.randperm(N)[:S] for _ in range(B)])`` as (B,S) int32 on the HOST, bit-identical to torch and
    consuming torch's global CPU generator exactly like those B calls (reference: models/pointnet_pp_8dir.py:28), but in
    one C call (pcoe_host_randperm_subsets) instead of B Python-level randperm launches.  ``out``: optional (pinned)
    int32 CPU tensor to fill."""
    if out is None:
        out = torch.empty(B, S, dtype=torch.int32)
    if out.is_cuda or out.dtype != torch.int32 or not out.is_contiguous() or out.numel() != B * S:
        raise ValueError("host_randperm_subsets: `out` must be a contiguous int32 CPU tensor of B*S elements")
    state = torch.get_rng_state()
    _lib.check(_lib.load().pcoe_host_randperm_subsets(state.data_ptr(), state.numel(), B, N, S, out.data_ptr()))
    torch.set_rng_state(state)
    return out


def gather_points(points: torch.Tensor, idx32: torch.Tensor) -> torch.Tensor:
    """points (B,N,C) f32, idx (B,S) int32 -> (B,S,C).  Reference: models/base.py:4-14."""
    points = _req(points, torch.float32, "points")
    idx32 = _req(idx32, torch.int32, "idx")
    B, N, Cc = points.shape
    S = idx32.size(1)
    out = torch.empty(B, S, Cc, dtype=torch.float32, device=points.device)
    _lib.check(_lib.load().pcoe_gather_points_f32(points.data_ptr(), B, N, Cc, idx32.data_ptr(), S,
                                                  out.data_ptr(), _stream()))
    return out


def random_subset(B: int, N: int, S: int, seed: int, offset: int, device,
                  counter: torch.Tensor | None = None, xyz: torch.Tensor | None = None):
    """(B,S) int32 uniform subset without replacement, drawn on the device.  ``counter`` (1-element
    int64 CUDA tensor) is added to ``offset`` on the device (CUDA-graph friendly).  With ``xyz`` (B,N,3) the
    selected points are gathered in the same launch and ``(idx, new_xyz)`` is returned."""
    out = torch.empty(B, S, dtype=torch.int32, device=device)
    if xyz is None:
        _lib.check(_lib.load().pcoe_random_subset(B, N, S, seed & (2**64 - 1), offset & (2**64 - 1), _ptr(counter),
                                                  out.data_ptr(), _stream()))
        return out
    xyz = _req(xyz, torch.float32, "xyz")
    new_xyz = torch.empty(B, S, 3, dtype=torch.float32, device=device)
    _lib.check(_lib.load().pcoe_random_subset_xyz(B, N, S, seed & (2**64 - 1), offset & (2**64 - 1), _ptr(counter),
                                                  out.data_ptr(), xyz.data_ptr(), new_xyz.data_ptr(), _stream()))
    return out, new_xyz


# ------------------------------------------------------------------------------------------------
# grouping
# ------------------------------------------------------------------------------------------------
def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """(B,N,C),(B,M,C) -> (B,N,M) f32.  Reference: models/base.py:20-27."""
    src = _req(src, torch.float32, "src")
    dst = _req(dst, torch.float32, "dst")
    if src.dim() != 3 or dst.dim() != 3 or src.size(0) != dst.size(0) or src.size(2) != dst.size(2):
        raise ValueError(f"square_distance: src {tuple(src.shape)} / dst {tuple(dst.shape)}")
    B, N, Cc = src.shape
    M = dst.size(1)
    out = torch.empty(B, N, M, dtype=torch.float32, device=src.device)
    _lib.check(_lib.load().pcoe_square_distance_f32(src.data_ptr(), dst.data_ptr(), B, N, M, Cc, out.data_ptr(), _stream()))
    return out


def knn_int32(new_xyz: torch.Tensor, xyz: torch.Tensor, nsample: int) -> torch.Tensor:
    new_xyz = _req(new_xyz, torch.float32, "new_xyz")
    xyz = _req(xyz, torch.float32, "xyz")
    if xyz.dim() != 3 or new_xyz.dim() != 3 or xyz.size(0) != new_xyz.size(0):
        raise ValueError(f"knn: xyz {tuple(xyz.shape)} / new_xyz {tuple(new_xyz.shape)}")
    B, N, _ = xyz.shape
    S = new_xyz.size(1)
    out = torch.empty(B, S, nsample, dtype=torch.int32, device=xyz.device)
    _lib.check(_lib.load().pcoe_knn_f32(xyz.data_ptr(), new_xyz.data_ptr(), B, N, S, nsample,
                                        out.data_ptr(), _stream()))
    return out


def ball_query_int32(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    new_xyz = _req(new_xyz, torch.float32, "new_xyz")
    xyz = _req(xyz, torch.float32, "xyz")
    B, N, _ = xyz.shape
    S = new_xyz.size(1)
    out = torch.empty(B, S, nsample, dtype=torch.int32, device=xyz.device)
    _lib.check(_lib.load().pcoe_ball_query_f32(xyz.data_ptr(), new_xyz.data_ptr(), B, N, S, nsample,
                                               float(radius), out.data_ptr(), _stream()))
    return out


def ball_query_multi_int32(radii, nsamples, xyz: torch.Tensor, new_xyz: torch.Tensor):
    """Multi-scale radius grouping in one launch: list of (B,S,nsample_r) int32, row r == ball_query_int32(radii[r],
    nsamples[r], ...) bit for bit (query_ball_point, PointNet++Demo.py:49-70, per scale)."""
    new_xyz = _req(new_xyz, torch.float32, "new_xyz")
    xyz = _req(xyz, torch.float32, "xyz")
    if len(radii) != len(nsamples) or not 1 <= len(radii) <= 4:
        raise ValueError("ball_query_multi: 1..4 (radius, nsample) pairs")
    B, N, _ = xyz.shape
    S = new_xyz.size(1)
    n = len(radii)
    outs = [torch.empty(B, S, int(k), dtype=torch.int32, device=xyz.device) for k in nsamples]
    r = (C.c_double * n)(*[float(x) for x in radii])
    k = (C.c_int * n)(*[int(x) for x in nsamples])
    o = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
    _lib.check(_lib.load().pcoe_ball_query_multi_f32(xyz.data_ptr(), new_xyz.data_ptr(), B, N, S, n, r, k, o, _stream()))
    return outs


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """(B,S,nsample) int64, slot-exact.  Reference: query_ball_point, PointNet++Demo.py:49-70."""
    return ball_query_int32(radius, nsample, xyz, new_xyz).long()


# ------------------------------------------------------------------------------------------------
# losses (value + gradient in one launch)
# ------------------------------------------------------------------------------------------------
def vm_kl_fwd_bwd(mu_p, kappa_p, mu_q, kappa_q, variant: int):
    mu_p = _req(mu_p, torch.float32, "mu_p"); kappa_p = _req(kappa_p, torch.float32, "kappa_p")
    mu_q = _req(mu_q, torch.float32, "mu_q"); kappa_q = _req(kappa_q, torch.float32, "kappa_q")
    n = mu_p.numel()
    for t in (kappa_p, mu_q, kappa_q):
        if t.numel() != n:
            raise ValueError("kl_von_mises: all four arguments must have the same number of elements")
    loss = torch.empty_like(mu_p); dmu = torch.empty_like(mu_p); dk = torch.empty_like(mu_p)
    _lib.check(_lib.load().pcoe_vm_kl_fwd_bwd(mu_p.data_ptr(), kappa_p.data_ptr(), mu_q.data_ptr(),
                                              kappa_q.data_ptr(), n, variant, loss.data_ptr(),
                                              dmu.data_ptr(), dk.data_ptr(), _stream()))
    return loss, dmu, dk


def mvm_match_fwd_bwd(mu, kappa, w, vm_gt, K_gt):
    mu = _req(mu, torch.float32, "mu"); kappa = _req(kappa, torch.float32, "kappa")
    w = _req(w, torch.float32, "w"); vm_gt = _req(vm_gt, torch.float32, "vm_gt")
    K_gt = _req(K_gt, torch.int32, "K_gt")
    if mu.dim() != 2 or vm_gt.dim() != 3 or vm_gt.size(1) < mu.size(1):
        raise ValueError(f"match_loss: mu {tuple(mu.shape)} vm_gt {tuple(vm_gt.shape)}")
    B, Kmax = mu.shape
    if vm_gt.size(1) != Kmax:
        vm_gt = vm_gt[:, :Kmax].contiguous()
    loss = torch.empty(B, dtype=torch.float32, device=mu.device)
    d3 = torch.empty(3, B, Kmax, dtype=torch.float32, device=mu.device)   # one buffer: backward scales it in one launch
    dmu, dk, dw = d3[0], d3[1], d3[2]
    perm = torch.empty(B, Kmax, dtype=torch.int32, device=mu.device)
    _lib.check(_lib.load().pcoe_mvm_match_fwd_bwd(mu.data_ptr(), kappa.data_ptr(), w.data_ptr(),
                                                  vm_gt.data_ptr(), vm_gt.size(2), K_gt.data_ptr(), B, Kmax,
                                                  loss.data_ptr(), dmu.data_ptr(), dk.data_ptr(),
                                                  dw.data_ptr(), perm.data_ptr(), _stream()))
    return loss, d3, perm


def soft_ce_fwd_bwd(logits, p):
    logits = _req(logits, torch.float32, "logits"); p = _req(p, torch.float32, "p_target")
    if logits.shape != p.shape or logits.dim() != 2:
        raise ValueError(f"soft CE: logits {tuple(logits.shape)} p {tuple(p.shape)}")
    B, Cc = logits.shape
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits)
    _lib.check(_lib.load().pcoe_soft_ce_fwd_bwd(logits.data_ptr(), p.data_ptr(), B, Cc, loss.data_ptr(),
                                                dl.data_ptr(), _stream()))
    return loss, dl
