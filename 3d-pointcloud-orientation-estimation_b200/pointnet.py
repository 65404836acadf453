"""Vanilla PointNet, inference path (BASELINE configs[4]).

Reference: models/pointnet.py - ``STN3d`` (:6-34), ``STNkd`` (:36-65), ``PointNetEncoder`` (:67-109), ``PointNet``
(:111-129).  Same class names, constructor arguments, sub-module names and therefore ``state_dict`` layout; default
initialisation is bit-identical under the same seed.

What runs where (eval mode):
  * every conv1d(k=1) + BatchNorm + ReLU stack followed by the max over the points - 3->64->128->1024 (input STN and
    encoder), 64->64->128->1024 (feature STN), 64->128->1024 (encoder after the feature transform) - is ONE
    ``pcoe_pointmlp_forward`` call: split-operand tcgen05 kernels over all B*N points, max pooled in the last layer's
    epilogue, nothing but the pooled (B,1024) features leaves the call;
  * the encoder's 3->64 layer, when its per-point output feeds the feature transform, is ``pcoe_pointwise_linear_f32``;
  * the two per-cloud transforms (``torch.bmm`` with 3x3 / 64x64 matrices, :89,97) and the (B,1024)->512->256->k
    fully connected tails are plain library GEMMs on B rows (torch / cuBLAS, fp32) as in the reference.
Training this model is not part of the hot path (the reference trains the PointNet++ heads); ``forward`` in train mode
raises.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .sa import _workspace


def _stack_params(convs, bns):
    P = _lib.SAParams()
    for i, (cv, bn) in enumerate(zip(convs, bns)):
        P.W[i] = cv.weight.data_ptr()
        P.bias[i] = cv.bias.data_ptr() if cv.bias is not None else None
        P.gamma[i] = bn.weight.data_ptr()
        P.beta[i] = bn.bias.data_ptr()
        P.running_mean[i] = bn.running_mean.data_ptr()
        P.running_var[i] = bn.running_var.data_ptr()
    return P


def pointmlp_max(xyz, feats, convs, bns, relu_last: bool, rows_per_cloud: int) -> torch.Tensor:
    """(clouds * rows_per_cloud) point-major rows -> (clouds, C_last): the conv/BN/ReLU stack + max over each cloud.
    xyz (M,3) or None, feats (M,D) or None (D a multiple of 64)."""
    lib = _lib.load()
    src = xyz if xyz is not None else feats
    M = src.size(0)
    D = 0 if feats is None else feats.size(1)
    Cs = [cv.out_channels for cv in convs]
    desc = _lib.PointMlpDesc(M=M, rows_per_cloud=rows_per_cloud, D=D, use_xyz=int(xyz is not None), nlayers=len(convs),
                             C=(C.c_int32 * 3)(*(Cs + [0] * (3 - len(Cs)))), relu_last=int(relu_last), eps=bns[0].eps)
    P = _stack_params(convs, bns)
    nbytes = lib.pcoe_pointmlp_workspace_bytes(C.byref(desc))
    if nbytes == 0:
        _lib.check(lib.pcoe_pointmlp_forward(C.byref(desc), None, None, C.byref(P), None, None, 0, None))
    ws = _workspace(src.device, nbytes)
    out = torch.empty(M // rows_per_cloud, Cs[-1], dtype=torch.float32, device=src.device)
    _lib.check(lib.pcoe_pointmlp_forward(C.byref(desc), ops._ptr(xyz), ops._ptr(feats), C.byref(P), out.data_ptr(),
                                         ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    return out


def _rows(x_cf: torch.Tensor):
    """(B,C,N) -> point-major (B*Np, C) with every cloud padded to a multiple of 32 rows by repeating its last point
    (the max over the cloud is unchanged), and Np."""
    B, Cc, N = x_cf.shape
    x = x_cf.transpose(2, 1)
    pad = (-N) % 32
    if pad:
        x = torch.cat([x, x[:, -1:, :].expand(B, pad, Cc)], dim=1)
    return x.contiguous().float().view(B * (N + pad), Cc), N + pad


def _require_eval(m: nn.Module, x: torch.Tensor):
    if m.training:
        raise NotImplementedError("pcoe: the vanilla PointNet is implemented as an inference path (call .eval()); "
                                  "training it is outside the hot path")
    if not x.is_cuda:
        raise RuntimeError("pcoe: PointNet runs on CUDA only (no CPU fallback)")


class _STN(nn.Module):
    def __init__(self, channel: int, k: int):
        super().__init__()
        self.k = k
        self.conv1 = nn.Conv1d(channel, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k * k)
        self.bn4 = nn.BatchNorm1d(512)
        self.bn5 = nn.BatchNorm1d(256)
        self.relu = nn.ReLU()

    def _tail(self, g: torch.Tensor) -> torch.Tensor:
        B = g.size(0)
        h = self.relu(self.bn4(self.fc1(g)))
        h = self.relu(self.bn5(self.fc2(h)))
        h = self.fc3(h)
        h = h + torch.eye(self.k, dtype=torch.float32, device=h.device).view(1, self.k * self.k)
        return h.view(B, self.k, self.k)

    def forward_rows(self, xyz_rows, feat_rows, rows_per_cloud: int) -> torch.Tensor:
        g = pointmlp_max(xyz_rows, feat_rows, [self.conv1, self.conv2, self.conv3], [self.bn1, self.bn2, self.bn3],
                         True, rows_per_cloud)
        return self._tail(g)


class STN3d(_STN):
    """models/pointnet.py:6-34: x (B,channel,N) -> (B,3,3).  channel must be 3 here."""

    def __init__(self, channel: int):
        if channel != 3:
            raise NotImplementedError("pcoe: STN3d is implemented for 3 input channels")
        super().__init__(channel, 3)

    def forward(self, x):
        _require_eval(self, x)
        rows, npad = _rows(x)
        return self.forward_rows(rows, None, npad)


class STNkd(_STN):
    """models/pointnet.py:36-65: x (B,k,N) -> (B,k,k).  k must be a multiple of 64."""

    def __init__(self, k: int = 64):
        if k % 64 != 0:
            raise NotImplementedError("pcoe: STNkd needs k to be a multiple of 64")
        super().__init__(k, k)

    def forward(self, x):
        _require_eval(self, x)
        rows, npad = _rows(x)
        return self.forward_rows(None, rows, npad)


class PointNetEncoder(nn.Module):
    """models/pointnet.py:67-109."""

    def __init__(self, global_feat: bool = True, feature_transform: bool = False, channel: int = 3):
        super().__init__()
        if channel != 3:
            raise NotImplementedError("pcoe: PointNetEncoder is implemented for xyz-only input (channel=3), as the "
                                      "reference's PointNet constructs it")
        self.stn = STN3d(channel)
        self.conv1 = nn.Conv1d(channel, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.global_feat = global_feat
        self.feature_transform = feature_transform
        if self.feature_transform:
            self.fstn = STNkd(k=64)

    def forward(self, x):
        """x (B,3,N) -> (global feature (B,1024) [or (B,1088,N) when global_feat is False], trans, trans_feat)."""
        _require_eval(self, x)
        B, D, N = x.shape
        if D != 3:
            raise NotImplementedError(f"pcoe: PointNetEncoder takes xyz-only input (B,3,N), got {tuple(x.shape)}")
        rows, npad = _rows(x)                                          # (B*Np,3)
        trans = self.stn.forward_rows(rows, None, npad)                # :84
        xt = torch.bmm(rows.view(B, npad, 3), trans).reshape(B * npad, 3)   # :89
        need_pointfeat = self.feature_transform or not self.global_feat
        if not need_pointfeat:
            g = pointmlp_max(xt, None, [self.conv1, self.conv2, self.conv3], [self.bn1, self.bn2, self.bn3], False, npad)
            return g, trans, None
        # per-point 64-channel features (they feed the feature transform / the concatenated output)
        s1 = self.bn1.weight / torch.sqrt(self.bn1.running_var + self.bn1.eps)
        t1 = self.bn1.bias + (self.conv1.bias - self.bn1.running_mean) * s1
        z1 = torch.empty(B * npad, 64, dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().pcoe_pointwise_linear_f32(xt.data_ptr(), B * npad, 3, self.conv1.weight.data_ptr(),
                                                         s1.contiguous().data_ptr(), t1.contiguous().data_ptr(), 64, 1,
                                                         z1.data_ptr(), torch.cuda.current_stream().cuda_stream))
        trans_feat = None
        if self.feature_transform:
            trans_feat = self.fstn.forward_rows(None, z1, npad)        # :95
            z1 = torch.bmm(z1.view(B, npad, 64), trans_feat).reshape(B * npad, 64)   # :96-98
        g = pointmlp_max(None, z1, [self.conv2, self.conv3], [self.bn2, self.bn3], False, npad)   # :102-105
        if self.global_feat:
            return g, trans, trans_feat
        pointfeat = z1.view(B, npad, 64)[:, :N].transpose(2, 1)
        return torch.cat([g.view(B, 1024, 1).repeat(1, 1, N), pointfeat], 1), trans, trans_feat


class PointNet(nn.Module):
    """models/pointnet.py:111-129: (B,N,3) or (B,3,N) -> (B,3)."""

    def __init__(self, feature_transform: bool = True):
        super().__init__()
        self.encoder = PointNetEncoder(global_feat=True, feature_transform=feature_transform, channel=3)
        self.fc1 = nn.Linear(1024, 512)
        self.bn1 = nn.BatchNorm1d(512)
        self.fc2 = nn.Linear(512, 256)
        self.bn2 = nn.BatchNorm1d(256)
        self.dropout = nn.Dropout(p=0.4)
        self.fc3 = nn.Linear(256, 3)
        self.relu = nn.ReLU()

    def forward(self, x):
        if x.dim() == 3 and x.shape[2] in (3, 6):
            x = x.transpose(1, 2)
        _require_eval(self, x)
        g, _, _ = self.encoder(x)
        h = self.relu(self.bn1(self.fc1(g)))
        h = self.relu(self.bn2(self.dropout(self.fc2(h))))
        return self.fc3(h)
