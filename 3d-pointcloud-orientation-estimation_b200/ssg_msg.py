"""True-FPS + radius-ball-query set abstraction (SSG), its classification model, and multi-scale grouping (MSG).

Reference: PointNet++Demo.py - ``SimpleSetAbstraction`` (:74-129), ``SimpleSetAbstractionGroupAll`` (:131-174),
``PointNetPlusPlusCls`` (:178-240; 512/0.2/32 -> 128/0.4/64 -> group-all, :189-191).  Same constructor arguments,
attribute names (``mlp_convs``, ``mlp_bns``, ``sa1..3``, ``fc1..3``, ``bn1/2``, ``dropout1/2``) and therefore
``state_dict`` layout, same ``(B, C, N)`` channel-first tensors at the module boundary.

``PointNetSetAbstractionMsg`` is the multi-radius extension SURVEY 8(f2) asks for: one FPS draw, ONE pass of the
multi-scale ball-query kernel (``pcoe_ball_query_multi_f32``) for all radii, one fused SA-MLP launch chain per scale on
the shared centroids, outputs concatenated along channels.  The reference has no MSG class; every scale computes
exactly what ``SimpleSetAbstraction`` computes for that (radius, nsample, mlp) on the same centroids, which is how the
oracle (oracle/sa_torch.py: ``msg_forward``) and the golden vectors are built.

Everything runs through libpcoe (FPS kernel, ball-query kernels, fused SA forward/backward); there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .sa import PRECISIONS, _SAFunction, get_default_precision


class _SAView:
    """What ``_SAFunction`` reads from a set-abstraction module, for layers whose parameters live under other
    attribute names (``mlp_convs`` / ``mlp_bns``, ``conv_blocks[i]`` / ``bn_blocks[i]``)."""

    def __init__(self, owner: nn.Module, convs, bns, group_all: bool = False):
        self._owner = owner
        self.convs, self.bns = convs, bns
        self.group_all = group_all
        self.direct_grad_accumulation = False
        self._after_backward = None

    @property
    def training(self) -> bool:
        return self._owner.training

    @property
    def precision(self) -> str:
        return self._owner.precision

    def params(self):
        out = []
        for conv, bn in zip(self.convs, self.bns):
            out += [conv.weight, conv.bias, bn.weight, bn.bias]
        return out


def _mlp(in_channel: int, mlp):
    if len(mlp) != 3:
        raise NotImplementedError("pcoe: the fused set-abstraction kernels implement 3-layer MLPs")
    convs, bns = nn.ModuleList(), nn.ModuleList()
    last = in_channel + 3
    for c in mlp:
        convs.append(nn.Conv2d(last, c, 1))
        bns.append(nn.BatchNorm2d(c))
        last = c
    return convs, bns


def _check_cf(xyz, points):
    if xyz.dim() != 3 or xyz.size(1) != 3:
        raise ValueError(f"xyz must be (B,3,N), got {tuple(xyz.shape)}")
    if not xyz.is_cuda:
        raise RuntimeError("pcoe: set abstraction runs on CUDA only (no CPU fallback)")
    x = xyz.transpose(2, 1).contiguous().float()
    p = None if points is None else points.transpose(2, 1).contiguous().float()
    return x, p


class _PrecisionMixin:
    _precision = None

    @property
    def precision(self) -> str:
        return self._precision or get_default_precision()


class SimpleSetAbstraction(_PrecisionMixin, nn.Module):
    """FPS -> radius ball query -> centre -> 3 x (conv1x1, BN, ReLU) -> max.  PointNet++Demo.py:74-129."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, *, precision: str | None = None):
        super().__init__()
        if precision is not None and precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self._precision = precision
        self.mlp_convs, self.mlp_bns = _mlp(in_channel, mlp)
        self._view = _SAView(self, self.mlp_convs, self.mlp_bns)
        self.last_fps_idx = None
        self.last_group_idx = None

    def forward(self, xyz, points, fps_idx=None):
        """xyz (B,3,N), points (B,D,N) or None -> new_xyz (B,3,npoint), new_points (B,mlp[-1],npoint).
        ``fps_idx`` (B,npoint) overrides the FPS draw (the reference draws the start index with the host generator,
        :20; ops.farthest_point_sample does the same when it is not given)."""
        x, p = _check_cf(xyz, points)
        if fps_idx is None:
            idx32, new_xyz = ops.farthest_point_sample(x, self.npoint, None, return_xyz=True, int32=True)
        else:
            idx32 = fps_idx.to(torch.int32).to(x.device)
            new_xyz = ops.gather_points(x, idx32)
        nbr = ops.ball_query_int32(self.radius, self.nsample, x, new_xyz)
        self.last_fps_idx, self.last_group_idx = idx32, nbr
        out = _SAFunction.apply(x, new_xyz, nbr, p, self._view, *self._view.params())
        return new_xyz.transpose(2, 1).contiguous(), out.transpose(2, 1).contiguous()


class SimpleSetAbstractionGroupAll(_PrecisionMixin, nn.Module):
    """One group of all points, absolute coordinates; new_xyz = the cloud's mean.  PointNet++Demo.py:131-174."""

    def __init__(self, in_channel, mlp, *, precision: str | None = None):
        super().__init__()
        self._precision = precision
        self.mlp_convs, self.mlp_bns = _mlp(in_channel, mlp)
        self._view = _SAView(self, self.mlp_convs, self.mlp_bns, group_all=True)

    def forward(self, xyz, points):
        x, p = _check_cf(xyz, points)
        out = _SAFunction.apply(x, None, None, p, self._view, *self._view.params())      # (B,1,C)
        new_xyz = x.mean(dim=1, keepdim=True).transpose(2, 1).contiguous()
        return new_xyz, out.transpose(2, 1).contiguous()


class PointNetSetAbstractionMsg(_PrecisionMixin, nn.Module):
    """Multi-scale grouping: shared FPS centroids, one (radius, nsample, 3-layer mlp) branch per scale, channel concat.

    npoint, radius_list, nsample_list, in_channel, mlp_list as in the usual PointNet++ MSG layer; every nsample must be a
    power of two <= 128 (the SA kernels' group sizes).  Channel order inside a branch is the reference's
    ``[xyz - centroid | features]`` (PointNet++Demo.py:113-118)."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list, *, precision: str | None = None):
        super().__init__()
        if not (len(radius_list) == len(nsample_list) == len(mlp_list)) or not 1 <= len(radius_list) <= 4:
            raise ValueError("MSG: 1..4 scales, one radius / nsample / mlp each")
        self.npoint = npoint
        self.radius_list, self.nsample_list = list(radius_list), list(nsample_list)
        self._precision = precision
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        self._views = []
        for mlp in mlp_list:
            convs, bns = _mlp(in_channel, mlp)
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
            self._views.append(_SAView(self, convs, bns))
        self.last_fps_idx = None
        self.last_group_idx = None

    def forward(self, xyz, points, fps_idx=None):
        """xyz (B,3,N), points (B,D,N) or None -> new_xyz (B,3,npoint), new_points (B,sum mlp[-1],npoint)."""
        x, p = _check_cf(xyz, points)
        if fps_idx is None:
            idx32, new_xyz = ops.farthest_point_sample(x, self.npoint, None, return_xyz=True, int32=True)
        else:
            idx32 = fps_idx.to(torch.int32).to(x.device)
            new_xyz = ops.gather_points(x, idx32)
        nbrs = ops.ball_query_multi_int32(self.radius_list, self.nsample_list, x, new_xyz)
        self.last_fps_idx, self.last_group_idx = idx32, nbrs
        outs = [_SAFunction.apply(x, new_xyz, nbr, p, v, *v.params()) for nbr, v in zip(nbrs, self._views)]
        return new_xyz.transpose(2, 1).contiguous(), torch.cat(outs, dim=2).transpose(2, 1).contiguous()


class PointNetPlusPlusCls(nn.Module):
    """SSG classification network of PointNet++Demo.py:178-240 (log-softmax scores)."""

    def __init__(self, num_classes=40, normal_channel=True, *, precision: str | None = None):
        super().__init__()
        in_channel = 3 if normal_channel else 0
        self.normal_channel = normal_channel
        self.sa1 = SimpleSetAbstraction(512, 0.2, 32, in_channel, [64, 64, 128], precision=precision)
        self.sa2 = SimpleSetAbstraction(128, 0.4, 64, 128, [128, 128, 256], precision=precision)
        self.sa3 = SimpleSetAbstractionGroupAll(256, [256, 512, 1024], precision=precision)
        self.fc1 = nn.Linear(1024, 512)
        self.bn1 = nn.BatchNorm1d(512)
        self.dropout1 = nn.Dropout(p=0.4)
        self.fc2 = nn.Linear(512, 256)
        self.bn2 = nn.BatchNorm1d(256)
        self.dropout2 = nn.Dropout(p=0.4)
        self.fc3 = nn.Linear(256, num_classes)

    def forward(self, x):
        B = x.size(0)
        if self.normal_channel:
            xyz, points = x[:, :3, :], x[:, 3:, :]
        else:
            xyz, points = x, None
        l1_xyz, l1_points = self.sa1(xyz, points)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        _, l3_points = self.sa3(l2_xyz, l2_points)
        x = l3_points.reshape(B, 1024)
        x = self.dropout1(F.relu(self.bn1(self.fc1(x))))
        x = self.dropout2(F.relu(self.bn2(self.fc2(x))))
        return F.log_softmax(self.fc3(x), dim=1)
