"""Drop-in ``PointNetSetAbstraction`` and the ``models/base.py`` helpers, backed by libpcoe.

Reference: models/pointnet_pp_8dir.py:6-43 (the canonical SA layer; byte-identical copies live in
Pointnet_pp_xyz.py, pointnet_pp.py, Pointnet_pp_xyz_Schedmit.py, pointnet_pp_Fwd.py) and
models/base.py:4-35.  Constructor signature, attribute names (npoint, nsample, group_all, convs,
bns) and therefore the ``state_dict`` layout are the reference's; the extra keyword-only arguments
select what the reference hard-codes.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, ops

_DEFAULT_PRECISION = "bf16x3"
PRECISIONS = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16, "bf16x3": _lib.PRECISION_BF16X3}
_layer_seq = 0          # construction index of the SA layers of this process: deterministic Philox stream ids


def _dp_rank() -> int:
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def set_default_precision(p: str) -> None:
    """'bf16x3' (the default: tcgen05 with every operand split into bf16 planes - three planes / six MMAs per product in
    the forward, two / three in the backward -, fp32 stored activations: fp32-class accuracy on the tensor pipe, the
    mode the parity tests gate; layer shapes it does not cover run on the fp32 kernels), 'fp32' (CUDA-core fp32 GEMMs)
    or 'bf16' (tcgen05, plain bf16 operands and stored activations: fastest, stated tolerance, not parity-gated)."""
    global _DEFAULT_PRECISION
    if p not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
    _DEFAULT_PRECISION = p


def get_default_precision() -> str:
    return _DEFAULT_PRECISION


# ------------------------------------------------------------------------------------------------
# models/base.py helpers (same names, same argument meaning)
# ------------------------------------------------------------------------------------------------
def index_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """points (B,N,C), idx (B,S) or (B,S,K) -> gathered points.  Reference: models/base.py:4-18."""
    if idx.dim() == 2:
        return ops.gather_points(points, idx.to(torch.int32))
    B, S, K = idx.shape
    return ops.gather_points(points, idx.reshape(B, S * K).to(torch.int32)).view(B, S, K, -1)


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """(B,N,M) squared distances, ``-2 src.dst + |src|^2 + |dst|^2``.  Reference: models/base.py:20-27.
    One libpcoe kernel (the matrix is written once); the hot path itself never materialises it."""
    return ops.square_distance(src, dst)


def query_ball_point(new_xyz: torch.Tensor, xyz: torch.Tensor, nsample: int) -> torch.Tensor:
    """kNN indices (B,S,nsample) int64 (the reference's misnamed helper, models/base.py:29-35)."""
    return ops.knn_int32(new_xyz, xyz, nsample).long()


# ------------------------------------------------------------------------------------------------
_workspaces: dict = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _fill3(arr, tensors):
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()


class _SAFunction(torch.autograd.Function):
    """pcoe_sa_forward / pcoe_sa_backward as one autograd node."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, nbr, feats, module, *params):
        # params = (W1,b1,g1,be1, W2,b2,g2,be2, W3,b3,g3,be3)
        lib = _lib.load()
        B, N, _ = xyz.shape
        group_all = module.group_all
        S = 1 if group_all else new_xyz.size(1)
        K = N if group_all else nbr.size(2)
        D = 0 if feats is None else feats.size(2)
        Ws, bs, gs, bes = params[0::4], params[1::4], params[2::4], params[3::4]
        train = module.training
        precision = module.precision
        if precision == "bf16x3" and not (K == 32 and D % 64 == 0 and all(w.size(0) % 64 == 0 for w in Ws)):
            # group sizes / widths the split-operand tcgen05 kernels do not cover run on the fp32 CUDA-core kernels
            # (same accuracy class, still libpcoe on the GPU)
            precision = "fp32"
        desc = _lib.SADesc(B=B, N=N, S=S, K=K, D=D, C1=Ws[0].size(0), C2=Ws[1].size(0), C3=Ws[2].size(0),
                           group_all=int(group_all), train=int(train),
                           precision=PRECISIONS[precision],
                           eps=module.bns[0].eps, momentum=module.bns[0].momentum or 0.1)
        P = _lib.SAParams()
        _fill3(P.W, Ws); _fill3(P.bias, bs); _fill3(P.gamma, gs); _fill3(P.beta, bes)
        _fill3(P.running_mean, [bn.running_mean for bn in module.bns])
        _fill3(P.running_var, [bn.running_var for bn in module.bns])
        # incremented inside pcoe_sa_forward (one thread of its last kernel) instead of three torch add_ launches
        _fill3(P.num_batches_tracked, [bn.num_batches_tracked if train else None for bn in module.bns])
        sv_bytes = lib.pcoe_sa_saved_bytes(C.byref(desc))
        ws_bytes = lib.pcoe_sa_workspace_bytes(C.byref(desc))
        if ws_bytes == 0:
            _lib.check(lib.pcoe_sa_forward(C.byref(desc), None, None, None, None, C.byref(P), None, None, 0, None, 0, None))
        saved = torch.empty(sv_bytes, dtype=torch.uint8, device=xyz.device) if train else None
        ws = _workspace(xyz.device, ws_bytes)
        out = torch.empty(B, S, Ws[2].size(0), dtype=torch.float32, device=xyz.device)
        _lib.check(lib.pcoe_sa_forward(C.byref(desc), xyz.data_ptr(), ops._ptr(new_xyz), ops._ptr(nbr),
                                       ops._ptr(feats), C.byref(P), out.data_ptr(), ops._ptr(saved), sv_bytes,
                                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
        ctx.desc, ctx.P, ctx.train = desc, P, train
        ctx.params = params if module.direct_grad_accumulation else None
        ctx.after_backward = getattr(module, "_after_backward", None)
        ctx.has_feats = feats is not None
        ctx.param_shapes = [p.shape for p in params]
        ctx.save_for_backward(xyz, new_xyz, nbr, feats, out, saved, *Ws)
        ctx.keep = (bs, gs, bes, [bn.running_mean for bn in module.bns], [bn.running_var for bn in module.bns])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.train:
            raise NotImplementedError("pcoe: backward through an eval-mode set-abstraction layer is not implemented "
                                      "(the reference never trains with BatchNorm in eval mode)")
        lib = _lib.load()
        xyz, new_xyz, nbr, feats, out, saved, W1, W2, W3 = ctx.saved_tensors
        desc, P = ctx.desc, ctx.P
        dev = xyz.device
        G = _lib.SAGrads()
        # direct accumulation (pcoe.dp.FlatGradBuffer): the kernels add into the p.grad views of the flat
        # gradient buffer and autograd receives None - no per-parameter `grad += g` kernel
        direct = ctx.params is not None and all(
            p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.device == dev
            for p in ctx.params)
        if direct:
            grads = [p.grad for p in ctx.params]
            G.accumulate = 1
        else:
            grads = [torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.param_shapes]
        _fill3(G.dW, grads[0::4]); _fill3(G.dbias, grads[1::4]); _fill3(G.dgamma, grads[2::4]); _fill3(G.dbeta, grads[3::4])
        gfeats = torch.empty_like(feats) if (ctx.has_feats and ctx.needs_input_grad[3]) else None
        ws = _workspace(dev, lib.pcoe_sa_workspace_bytes(C.byref(desc)))
        _lib.check(lib.pcoe_sa_backward(C.byref(desc), xyz.data_ptr(), ops._ptr(new_xyz), ops._ptr(nbr),
                                        ops._ptr(feats), C.byref(P), out.data_ptr(),
                                        grad_out.contiguous().data_ptr(), saved.data_ptr(), saved.numel(),
                                        ops._ptr(gfeats), C.byref(G), ws.data_ptr(), ws.numel(),
                                        torch.cuda.current_stream().cuda_stream))
        if direct:
            if ctx.after_backward is not None:
                ctx.after_backward()               # pcoe.dp.DataParallel: start the all-reduce of the finished bucket
            return (None, None, None, gfeats, None) + (None,) * len(grads)
        return (None, None, None, gfeats, None, *grads)


class PointNetSetAbstraction(nn.Module):
    """Sample -> group -> centre -> 3 x (1x1 conv, BatchNorm, ReLU) -> max over neighbours.

    Positional arguments are the reference's (models/pointnet_pp_8dir.py:7).  Keyword-only extras:

    sampler   'randperm_host' (default; the reference: ``torch.randperm(N)[:npoint]`` per cloud on the
              CPU generator, :28 - index-exact under the same ``torch.manual_seed``), 'randperm_device'
              (same distribution drawn by a CUDA kernel, no host work; pointnet_pp_Fwd.py:44-47),
              or 'fps' (true farthest-point sampling, PointNet++Demo.py:8-29).
    grouper   'knn' (default; what the reference's ``query_ball_point`` computes, base.py:29-35) or
              'ball' (radius query of PointNet++Demo.py:49-70; needs ``radius``).
    precision 'fp32' | 'bf16x3' | 'bf16' | None (= pcoe default at call time); see set_default_precision.
    """

    def __init__(self, npoint, nsample, in_channel, mlp_channels, group_all=False, *,
                 sampler: str = "randperm_host", grouper: str = "knn", radius: float | None = None,
                 precision: str | None = None):
        super().__init__()
        if len(mlp_channels) != 3:
            raise NotImplementedError("pcoe: the fused set-abstraction kernels implement the reference's 3-layer MLP")
        if sampler not in ("randperm_host", "randperm_device", "fps"):
            raise ValueError(f"unknown sampler {sampler!r}")
        if grouper not in ("knn", "ball") or (grouper == "ball" and radius is None):
            raise ValueError("grouper must be 'knn' or 'ball' (with radius)")
        self.npoint = npoint
        self.nsample = nsample
        self.group_all = group_all
        self.sampler, self.grouper, self.radius = sampler, grouper, radius
        if precision is not None and precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self._precision = precision
        self._rng_counter = None
        self._static_idx = None                    # set by pcoe.GraphedTrainStep (host-replayed subsets as a graph input)
        self._last_draw_shape = None
        global _layer_seq
        _layer_seq += 1
        self._stream_id = _layer_seq              # reproducible across runs (unlike id(self)), distinct per layer
        # set by pcoe.dp.FlatGradBuffer: backward adds parameter gradients straight into p.grad
        self.direct_grad_accumulation = False
        self._after_backward = None                # pcoe.dp.DataParallel hook, called at the end of this layer's backward
        self.last_fps_idx = None
        self.last_group_idx = None

        last_ch = in_channel + 3
        self.convs = nn.ModuleList()
        self.bns = nn.ModuleList()
        for out_ch in mlp_channels:
            self.convs.append(nn.Conv2d(last_ch, out_ch, 1))
            self.bns.append(nn.BatchNorm2d(out_ch))
            last_ch = out_ch

    @property
    def precision(self) -> str:
        return self._precision or _DEFAULT_PRECISION

    def _sample(self, xyz: torch.Tensor):
        """-> (idx32 (B,S), new_xyz (B,S,3) or None when the sampler does not gather)."""
        B, N, _ = xyz.shape
        if self.sampler == "randperm_host":
            self._last_draw_shape = (B, N, self.npoint)
            if self._static_idx is not None:
                # pcoe.GraphedTrainStep owns the draw: it replays the reference's host generator before every graph
                # replay and uploads the subsets into this static buffer (a graph input)
                return self._static_idx, None
            # the reference's torch.stack([torch.randperm(N)[:npoint] ...]) on torch's CPU generator, bit-identical
            # (same subsets, same generator state afterwards), in one C call instead of B Python-level launches
            idx = ops.host_randperm_subsets(B, N, self.npoint)
            return idx.to(xyz.device, non_blocking=True), None
        if self.sampler == "randperm_device":
            # the call counter lives on the device so that a CUDA-graph replay draws new subsets
            if self._rng_counter is None or self._rng_counter.device != xyz.device:
                self._rng_counter = torch.zeros(1, dtype=torch.int64, device=xyz.device)
            self._rng_counter.add_(1)
            # Philox stream = (layer construction index, data-parallel rank): same subsets for the same seed on every
            # run, different subsets on different ranks and layers
            offset = ((self._stream_id & 0xFFFF) << 48) | ((_dp_rank() & 0xFFFF) << 32)
            return ops.random_subset(B, N, self.npoint, torch.initial_seed(), offset, xyz.device,
                                     self._rng_counter, xyz=xyz)
        # 'fps': the first centroid of every cloud is drawn ON THE DEVICE (torch's CUDA generator: no host sync, and a
        # CUDA-graph replay draws fresh start points); ops.farthest_point_sample keeps the reference's host draw
        start = torch.randint(0, N, (B,), device=xyz.device, dtype=torch.int32)
        idx, new_xyz = ops.farthest_point_sample(xyz, self.npoint, start, return_xyz=True, int32=True)
        return idx, new_xyz

    @staticmethod
    def _check_xyz(xyz):
        if xyz.dim() != 3 or xyz.size(-1) != 3:
            raise ValueError(f"xyz must be (B,N,3), got {tuple(xyz.shape)}")
        if not xyz.is_cuda:
            raise RuntimeError("pcoe: PointNetSetAbstraction runs on CUDA only (no CPU fallback); move the model and "
                               "inputs to a B200")
        return xyz.contiguous().float()

    def sample(self, xyz, fps_idx=None):
        """Sampling half of forward: (idx32 (B,S), new_xyz (B,S,3)).  Depends on the coordinates only, so a caller can
        run it (and ``group``) for the NEXT layer beside this layer's MLP (pcoe.models._Backbone does)."""
        xyz = self._check_xyz(xyz)
        idx32, new_xyz = self._sample(xyz) if fps_idx is None else (fps_idx.to(torch.int32).to(xyz.device), None)
        if new_xyz is None:
            new_xyz = ops.gather_points(xyz, idx32)
        return idx32, new_xyz

    def group(self, xyz, new_xyz):
        """Grouping half of forward: neighbour indices (B,S,nsample) int32."""
        if self.grouper == "knn":
            return ops.knn_int32(new_xyz, xyz, self.nsample)
        return ops.ball_query_int32(self.radius, self.nsample, xyz, new_xyz)

    def forward(self, xyz, points, fps_idx=None, pre=None):
        """``pre`` = (idx32, new_xyz, nbr) computed earlier by ``sample`` / ``group`` (same results, other stream)."""
        xyz = self._check_xyz(xyz)
        feats = None if points is None else points.contiguous().float()
        params = []
        for conv, bn in zip(self.convs, self.bns):
            params += [conv.weight, conv.bias, bn.weight, bn.bias]
        if self.group_all:
            new_xyz = torch.zeros(xyz.size(0), 1, 3, device=xyz.device)
            out = _SAFunction.apply(xyz, None, None, feats, self, *params)
            return new_xyz, out
        if pre is None:
            idx32, new_xyz = self.sample(xyz, fps_idx)
            nbr = self.group(xyz, new_xyz)
        else:
            idx32, new_xyz, nbr = pre
        self.last_fps_idx, self.last_group_idx = idx32, nbr
        out = _SAFunction.apply(xyz, new_xyz, nbr, feats, self, *params)
        return new_xyz, out
