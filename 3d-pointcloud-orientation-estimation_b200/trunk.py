"""Fused trunk + mixture head of PointNetPPMvM (models/pointnet_pp_mvM.py:56-66,79-84,91-125) on libpcoe's fp32
building blocks (csrc/trunk.cu): 7 launches forward and 10 backward instead of ~45 + ~95 torch launches.

    fc1 -> LayerNorm -> ReLU -> dropout -> fc2 -> LayerNorm -> ReLU -> dropout -> head_pi | head_mu | head_kappa
        -> (mu, kappa, weight)

One autograd node.  Parameter gradients are added straight into ``p.grad`` when the module's gradients live in a
``pcoe.dp.FlatGradBuffer`` (same contract as the set-abstraction layers); otherwise they are returned to autograd.
Dropout draws its keep mask from a Philox stream keyed by (torch.initial_seed(), per-module device counter): the same
distribution as ``nn.Dropout`` but not torch's random stream (the reference's CPU masks cannot be reproduced on a GPU
either).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_PP = C.c_void_p * 3
_NN = C.c_int * 3


def _ptrs(ts):
    a = _PP()
    for i, t in enumerate(ts):
        a[i] = None if t is None else t.data_ptr()
    return a


def _ints(ns):
    a = _NN()
    for i, n in enumerate(ns):
        a[i] = int(n)
    return a


_MAX_PARTS = 8


def _linear_fwd(lib, st, x, Ws, bs, out, max_parts=1):
    """Returns the number of partial results written to `out` (1 = final, bias included)."""
    n = C.c_int(1)
    _lib.check(lib.pcoe_linear_fwd(x.data_ptr(), x.size(0), x.size(1), len(Ws), _ptrs(Ws), _ptrs(bs),
                                   _ints([w.size(0) for w in Ws]), out.data_ptr(), max_parts, C.byref(n), st))
    return n.value


_side_streams: dict = {}


def _side_stream(dev):
    """Auxiliary stream for the weight-gradient kernels of the trunk backward: only the optimizer consumes them, so
    they run beside the dx chain (forked per linear layer, joined at the end of the node).  None while per-kernel
    profiling is on (kernels are then timed alone)."""
    if _lib.profiling:
        return None
    s = _side_streams.get(dev.index)
    if s is None:
        s = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
    return s


def _linear_bwd(lib, st, dy, x, Ws, dWs, dbs, dx, accumulate, side=None):
    Ns = _ints([w.size(0) for w in Ws])
    st_w = st
    if side is not None:                       # fork: dy is ready on the current stream
        side.wait_stream(torch.cuda.current_stream())
        st_w = side.cuda_stream
    _lib.check(lib.pcoe_linear_bwd_dw(dy.data_ptr(), x.data_ptr(), x.size(0), x.size(1), len(Ws), Ns, _ptrs(dWs),
                                      _ptrs(dbs), int(accumulate), st_w))
    if dx is not None:
        _lib.check(lib.pcoe_linear_bwd_dx(dy.data_ptr(), x.size(0), x.size(1), len(Ws), _ptrs(Ws), Ns, dx.data_ptr(), st))


class MvMTrunkHead(torch.autograd.Function):
    """x (B,1024) + the 14 trunk / head parameters -> (mu, kappa, weight)."""

    @staticmethod
    def forward(ctx, x, cfg, fc1w, fc1b, g1, b1, fc2w, fc2b, g2, b2, piw, pib, muw, mub, kw, kb):
        lib = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        x = x.contiguous().float()
        B, dev = x.size(0), x.device
        f32 = dict(dtype=torch.float32, device=dev)
        train, p, seed, counter, eps1, eps2, temp, kmax, direct = cfg
        K = piw.size(0)
        h1, a1 = torch.empty(B, fc1w.size(0), **f32), torch.empty(B, fc1w.size(0), **f32)
        h2, a2 = torch.empty(B, fc2w.size(0), **f32), torch.empty(B, fc2w.size(0), **f32)
        st1, st2 = torch.empty(2, B, **f32), torch.empty(2, B, **f32)
        m1 = torch.empty(B, fc1w.size(0), dtype=torch.uint8, device=dev)
        m2 = torch.empty(B, fc2w.size(0), dtype=torch.uint8, device=dev)
        raw = torch.empty(B * 4 * K, **f32)                         # segment-major: pi [B,K] | mu_raw [B,2K] | kappa_raw [B,K]
        out = torch.empty(3, B, K, **f32)                           # mu, kappa, weight
        cptr = None if counter is None else counter.data_ptr()
        parts = torch.empty(_MAX_PARTS * B * fc1w.size(0), **f32)   # split-contraction partials (deterministic sum in LN)
        n1 = _linear_fwd(lib, st, x, [fc1w], [fc1b], parts, _MAX_PARTS)
        _lib.check(lib.pcoe_ln_relu_dropout_fwd(parts.data_ptr(), n1, fc1b.data_ptr(), h1.data_ptr(), g1.data_ptr(),
                                                b1.data_ptr(), B, h1.size(1), eps1, p, int(train), seed, cptr,
                                                a1.data_ptr(), st1[0].data_ptr(), st1[1].data_ptr(), m1.data_ptr(), st))
        n2 = _linear_fwd(lib, st, a1, [fc2w], [fc2b], parts, _MAX_PARTS)
        clamp = kmax is not None
        # LayerNorm 2 + ReLU + dropout, the three heads and the head transform are all per-row work: one launch
        heads_w, heads_b = [piw, muw, kw], [pib, mub, kb]
        _lib.check(lib.pcoe_ln_relu_dropout_heads_fwd(
            parts.data_ptr(), n2, fc2b.data_ptr(), h2.data_ptr(), g2.data_ptr(), b2.data_ptr(), B, h2.size(1), eps2, p,
            int(train), seed ^ 0x9E3779B97F4A7C15, cptr, a2.data_ptr(), st2[0].data_ptr(), st2[1].data_ptr(),
            m2.data_ptr(), 3, _ptrs(heads_w), _ptrs(heads_b), _ints([w.size(0) for w in heads_w]), raw.data_ptr(), K,
            float(temp), float(kmax) if clamp else 0.0, int(clamp), out[2].data_ptr(), out[0].data_ptr(),
            out[1].data_ptr(), st))
        ctx.save_for_backward(x, h1, a1, h2, a2, st1, st2, m1, m2, raw, fc1w, fc2w, piw, muw, kw, g1, g2)
        ctx.cfg = cfg
        ctx.params = (fc1w, fc1b, g1, b1, fc2w, fc2b, g2, b2, piw, pib, muw, mub, kw, kb) if direct else None
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_mu, g_k, g_w):
        lib = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        x, h1, a1, h2, a2, st1, st2, m1, m2, raw, fc1w, fc2w, piw, muw, kw, g1, g2 = ctx.saved_tensors
        train, p, seed, counter, eps1, eps2, temp, kmax, _ = ctx.cfg
        B, K, dev = x.size(0), piw.size(0), x.device
        f32 = dict(dtype=torch.float32, device=dev)
        direct = ctx.params is not None and all(q.grad is not None and q.grad.is_contiguous() for q in ctx.params)
        if direct:
            grads = [q.grad for q in ctx.params]
        else:
            shapes = [fc1w.shape, (fc1w.size(0),), g1.shape, g1.shape, fc2w.shape, (fc2w.size(0),), g2.shape, g2.shape,
                      piw.shape, (K,), muw.shape, (2 * K,), kw.shape, (K,)]
            grads = [torch.zeros(s, **f32) for s in shapes]          # LayerNorm gradients are accumulated with atomics
        d1w, d1b, dg1, db1, d2w, d2b, dg2, db2, dpw, dpb, dmw, dmb, dkw, dkb = grads
        acc = int(direct)
        ptr = lambda t: None if t is None else t.contiguous().data_ptr()
        draw = torch.empty_like(raw)
        clamp = kmax is not None
        dh2 = torch.empty_like(h2)
        side = _side_stream(dev)
        heads_w = [piw, muw, kw]
        # head-transform backward -> heads' data gradient -> LayerNorm 2 backward, per row, in one launch (the
        # LayerNorm backward also adds the column sums of its dx into the bias gradient of the linear below it)
        _lib.check(lib.pcoe_heads_ln_relu_dropout_bwd(
            raw.data_ptr(), K, float(temp), float(kmax) if clamp else 0.0, int(clamp), ptr(g_w), ptr(g_mu), ptr(g_k),
            draw.data_ptr(), 3, _ptrs(heads_w), _ints([w.size(0) for w in heads_w]), h2.data_ptr(), a2.data_ptr(),
            g2.data_ptr(), st2[0].data_ptr(), st2[1].data_ptr(), m2.data_ptr(), B, h2.size(1), p, int(train),
            dh2.data_ptr(), dg2.data_ptr(), db2.data_ptr(), d2b.data_ptr(), st))
        # the heads' weight / bias gradients (side stream)
        _linear_bwd(lib, st, draw, a2, heads_w, [dpw, dmw, dkw], [dpb, dmb, dkb], None, acc, side)
        da1, dh1 = torch.empty_like(a1), torch.empty_like(h1)
        _linear_bwd(lib, st, dh2, a1, [fc2w], [d2w], [None], da1, acc, side)
        _lib.check(lib.pcoe_ln_relu_dropout_bwd(da1.data_ptr(), h1.data_ptr(), a1.data_ptr(), g1.data_ptr(), st1[0].data_ptr(),
                                                st1[1].data_ptr(), m1.data_ptr(), B, h1.size(1), p, int(train),
                                                dh1.data_ptr(), dg1.data_ptr(), db1.data_ptr(), d1b.data_ptr(), st))
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        _linear_bwd(lib, st, dh1, x, [fc1w], [d1w], [None], dx, acc, side)
        if side is not None:                   # join before the node returns (buffers above are freed on this stream)
            torch.cuda.current_stream().wait_stream(side)
        if direct:
            return (dx, None) + (None,) * 14
        return (dx, None, *grads)
