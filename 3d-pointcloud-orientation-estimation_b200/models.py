"""Drop-in model classes: same constructors, ``forward(xyz)`` contracts and ``state_dict`` keys as
the reference, with the three set-abstraction layers running in libpcoe.

    PointNetPP8Dir          models/pointnet_pp_8dir.py:58-85      logits (B,8)
    PointNetPPVonMises      models/pointnet_pp_vonMises.py:8-38   mu (B,), kappa (B,)
    PointNetPPMvM           models/pointnet_pp_mvM.py:30-127      mu, kappa, weight (B,K)
    PointNetPPXYZ           models/Pointnet_pp_xyz.py:47-90       two unit vectors (B,3)
    PointNetPP              models/pointnet_pp.py:45-68           (B,3)
    PointNetPPXYZ_Schedmit  models/Pointnet_pp_xyz_Schedmit.py:47-92   two unit vectors (B,3)
    PointNetPPFwd           models/pointnet_pp_Fwd.py:77-98       unit vector (B,3) (device randperm)

The 1024->512->256 trunk and the heads (1.3 MFLOP per cloud) keep their torch.nn parameter modules - the
boundary SURVEY.md draws - so optimizers, clipping and checkpoints work unchanged; PointNetPPMvM evaluates them
with libpcoe's fused fp32 trunk kernels (pcoe.trunk), the BatchNorm1d heads with torch.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .sa import PointNetSetAbstraction, _dp_rank
from .trunk import MvMTrunkHead

# 8 horizontal directions, 45 degree steps starting from the original "forward" [0,0,-1]
# (values of models/pointnet_pp_8dir.py:46-55; imported by train_8dir_KL.py:14)
DIRS_8 = torch.tensor([[math.sin(math.radians(45 * k)), 0.0, -math.cos(math.radians(45 * k))] for k in range(8)],
                      dtype=torch.float64).mul(1e4).round().div(1e4).float() + 0.0


class _Backbone(nn.Module):
    """sa1/sa2/sa3 + fc1/fc2 with the reference's attribute names."""

    def __init__(self, norm: str = "bn", p_drop: float = 0.5, sa_kwargs: dict | None = None):
        super().__init__()
        kw = dict(sa_kwargs or {})
        self.sa1 = PointNetSetAbstraction(128, 32, 0, [64, 64, 128], **kw)
        self.sa2 = PointNetSetAbstraction(32, 32, 128, [128, 128, 256], **kw)
        self.sa3 = PointNetSetAbstraction(None, None, 256, [256, 512, 1024], group_all=True,
                                          **{k: v for k, v in kw.items() if k == "precision"})
        # registration order = the reference's, so state_dict() iterates identically
        self.fc1 = nn.Linear(1024, 512)
        if norm == "bn":
            self.bn1 = nn.BatchNorm1d(512)
        else:
            self.ln1 = nn.LayerNorm(512)
        self.fc2 = nn.Linear(512, 256)
        if norm == "bn":
            self.bn2 = nn.BatchNorm1d(256)
        else:
            self.ln2 = nn.LayerNorm(256)
        self.drop = nn.Dropout(p_drop)

    def _sa_features(self, xyz: torch.Tensor) -> torch.Tensor:
        B = xyz.size(0)
        l1_xyz, l1_pts = self.sa1(xyz, None)
        l2_xyz, l2_pts = self.sa2(l1_xyz, l1_pts)
        _, l3_pts = self.sa3(l2_xyz, l2_pts)
        return l3_pts.view(B, -1)

    def _bn_trunk(self, xyz: torch.Tensor) -> torch.Tensor:
        x = self._sa_features(xyz)
        x = F.relu(self.bn1(self.fc1(x)))
        x = F.relu(self.bn2(self.fc2(x)))
        return self.drop(x)


class PointNetPP8Dir(_Backbone):
    def __init__(self, **sa_kwargs):
        super().__init__("bn", 0.5, sa_kwargs)
        self.fc3 = nn.Linear(256, 8)

    def forward(self, xyz):
        return self.fc3(self._bn_trunk(xyz))


class PointNetPPVonMises(_Backbone):
    def __init__(self, **sa_kwargs):
        super().__init__("bn", 0.5, sa_kwargs)
        self.fc3 = nn.Linear(256, 2)

    def forward(self, xyz):
        out = self.fc3(self._bn_trunk(xyz))
        mu = torch.tanh(out[:, 0]) * math.pi
        kappa = F.softplus(out[:, 1])
        return mu, kappa


class PointNetPP(_Backbone):
    def __init__(self, **sa_kwargs):
        super().__init__("bn", 0.5, sa_kwargs)
        self.fc3 = nn.Linear(256, 3)

    def forward(self, x):
        return self.fc3(self._bn_trunk(x))


class PointNetPPFwd(_Backbone):
    """The reference draws its random subset on the device in this variant (pointnet_pp_Fwd.py:44-47)."""

    def __init__(self, **sa_kwargs):
        sa_kwargs.setdefault("sampler", "randperm_device")
        super().__init__("bn", 0.5, sa_kwargs)
        self.fc3 = nn.Linear(256, 3)

    def forward(self, xyz):
        return F.normalize(self.fc3(self._bn_trunk(xyz)), dim=1)


class PointNetPPXYZ(_Backbone):
    def __init__(self, **sa_kwargs):
        super().__init__("bn", 0.5, sa_kwargs)
        self.head_x = nn.Linear(256, 3)
        self.head_y = nn.Linear(256, 3)

    def forward(self, x):
        feat = self._bn_trunk(x)
        return F.normalize(self.head_x(feat), p=2, dim=1), F.normalize(self.head_y(feat), p=2, dim=1)


class PointNetPPXYZ_Schedmit(_Backbone):
    def __init__(self, **sa_kwargs):
        super().__init__("bn", 0.5, sa_kwargs)
        self.head_y = nn.Linear(256, 3)
        self.head_z = nn.Linear(256, 3)

    def forward(self, x):
        feat = self._bn_trunk(x)
        return F.normalize(self.head_y(feat), p=2, dim=1), F.normalize(self.head_z(feat), p=2, dim=1)


def _maybe_transpose_xyz(xyz: torch.Tensor) -> torch.Tensor:
    """Accept (B,N,3) or (B,3,N); return (B,N,3).  Same acceptance rule and errors as
    models/pointnet_pp_mvM.py:15-27 (which goes to (B,3,N) and straight back, :77)."""
    assert xyz.dim() == 3, f"xyz should be 3D tensor, got {xyz.shape}"
    B, A, Cc = xyz.shape
    if Cc == 3:
        return xyz
    if A == 3:
        return xyz.transpose(1, 2).contiguous()
    raise ValueError(f"xyz must be (B,N,3) or (B,3,N), got {xyz.shape}")


class _MvMHead(torch.autograd.Function):
    """(pi, mu_raw, kappa_raw) -> (mu, kappa, weight): pcoe_mvm_head_fwd / _bwd, one launch each instead of the
    ~20 (+ ~35 in backward) elementwise ops of models/pointnet_pp_mvM.py:91-125."""

    @staticmethod
    def forward(ctx, pi, mu_raw, kappa_raw, temp, kappa_max):
        pi, mu_raw, kappa_raw = pi.contiguous().float(), mu_raw.contiguous().float(), kappa_raw.contiguous().float()
        B, K = pi.shape
        w, mu, kp = torch.empty_like(pi), torch.empty_like(pi), torch.empty_like(pi)
        clamp = kappa_max is not None
        _lib.check(_lib.load().pcoe_mvm_head_fwd(pi.data_ptr(), mu_raw.data_ptr(), kappa_raw.data_ptr(), B, K, float(temp),
                                                 float(kappa_max) if clamp else 0.0, int(clamp), w.data_ptr(),
                                                 mu.data_ptr(), kp.data_ptr(), torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(pi, mu_raw, kappa_raw)
        ctx.cfg = (float(temp), float(kappa_max) if clamp else 0.0, int(clamp))
        return mu, kp, w

    @staticmethod
    def backward(ctx, g_mu, g_k, g_w):
        pi, mu_raw, kappa_raw = ctx.saved_tensors
        B, K = pi.shape
        temp, kmax, clamp = ctx.cfg
        d_pi, d_mu, d_k = torch.empty_like(pi), torch.empty_like(mu_raw), torch.empty_like(kappa_raw)
        ptr = lambda t: None if t is None else t.contiguous().data_ptr()
        _lib.check(_lib.load().pcoe_mvm_head_bwd(pi.data_ptr(), mu_raw.data_ptr(), kappa_raw.data_ptr(), B, K, temp, kmax,
                                                 clamp, ptr(g_w), ptr(g_mu), ptr(g_k), d_pi.data_ptr(), d_mu.data_ptr(),
                                                 d_k.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return d_pi, d_mu, d_k, None, None


class PointNetPPMvM(_Backbone):
    """Mixture-of-von-Mises head.  The three host synchronisations of the reference forward
    (isfinite prints and the ``(norm < 1e-3).any()`` branch, :98,108,118) are removed: the fallback
    ``where`` is applied unconditionally, which is value- and gradient-identical."""

    def __init__(self, max_K: int = 4, kappa_max: float = 80.0, p_drop: float = 0.4, temp: float = 0.7, **sa_kwargs):
        super().__init__("ln", p_drop, sa_kwargs)
        self.max_K = max_K
        self.kappa_max = float(kappa_max)
        self.temp = float(temp)
        hidden = 256
        self.head_pi = nn.Linear(hidden, max_K)
        self.head_mu = nn.Linear(hidden, max_K * 2)
        self.head_kappa = nn.Linear(hidden, max_K)
        nn.init.zeros_(self.head_pi.weight)
        nn.init.zeros_(self.head_pi.bias)
        nn.init.zeros_(self.head_mu.weight)
        nn.init.zeros_(self.head_mu.bias)
        nn.init.constant_(self.head_kappa.bias, 0.0)
        self.fused_trunk = True                    # False: fc1..heads as torch.nn modules (the reference's formulation)
        self.direct_grad_accumulation = False      # set by pcoe.dp.FlatGradBuffer
        self._drop_counter = None

    def _global_feat(self, xyz_bn3: torch.Tensor) -> torch.Tensor:
        x = self._sa_features(xyz_bn3)
        x = self.drop(F.relu(self.ln1(self.fc1(x))))
        x = self.drop(F.relu(self.ln2(self.fc2(x))))
        return x

    def forward(self, xyz: torch.Tensor):
        xyz = _maybe_transpose_xyz(xyz)
        if self.fused_trunk and self.max_K <= 8 and xyz.is_cuda:
            return self._fused(self._sa_features(xyz))
        feat = self._global_feat(xyz)
        if self.max_K <= 8:
            return _MvMHead.apply(self.head_pi(feat), self.head_mu(feat), self.head_kappa(feat), self.temp, self.kappa_max)
        return self._head_torch(feat)

    def _fused(self, x: torch.Tensor):
        """fc1 .. heads .. (mu, kappa, weight) as one autograd node on libpcoe's trunk kernels (pcoe.trunk)."""
        train = self.training
        seed, counter = 0, None
        if train and self.drop.p > 0:
            if torch.cuda.is_current_stream_capturing():
                # CUDA-graph capture: the call counter lives on the device so that every replay draws fresh masks
                if self._drop_counter is None or self._drop_counter.device != x.device:
                    raise RuntimeError("pcoe: run one eager training step before capturing (allocates the dropout counter)")
                self._drop_counter.add_(1)
                # data-parallel ranks share torch.initial_seed(): mix the rank in so that they draw different masks
                seed, counter = torch.initial_seed() + 0x5DEECE66D * _dp_rank(), self._drop_counter
            else:
                # eager: consume torch's CUDA generator like nn.Dropout would (reproducible under torch.manual_seed)
                if self._drop_counter is None or self._drop_counter.device != x.device:
                    self._drop_counter = torch.zeros(1, dtype=torch.int64, device=x.device)
                gen = torch.cuda.default_generators[x.device.index]
                off = gen.get_offset()
                gen.set_offset(off + 4)
                seed = gen.initial_seed() + 0x9E3779B97F4A7C15 * (off // 4 + 1)
        cfg = (train, float(self.drop.p), seed & 0x7FFFFFFFFFFFFFFF, counter,
               self.ln1.eps, self.ln2.eps, self.temp, self.kappa_max, self.direct_grad_accumulation)
        return MvMTrunkHead.apply(x, cfg, self.fc1.weight, self.fc1.bias, self.ln1.weight, self.ln1.bias,
                                  self.fc2.weight, self.fc2.bias, self.ln2.weight, self.ln2.bias,
                                  self.head_pi.weight, self.head_pi.bias, self.head_mu.weight, self.head_mu.bias,
                                  self.head_kappa.weight, self.head_kappa.bias)

    def _head_torch(self, feat):
        """The head transform spelled with torch ops (the reference's formulation; used for max_K > 8 and as the
        fp32 reference of the fused kernels in the tests)."""
        weight = F.softmax(self.head_pi(feat) / self.temp, dim=-1)
        mu_raw = self.head_mu(feat).view(-1, self.max_K, 2)
        mu_unit = F.normalize(mu_raw, dim=-1, eps=1e-4)
        c, s = mu_unit[..., 0], mu_unit[..., 1]
        norm = torch.sqrt(c * c + s * s)
        mask = norm < 1e-3
        c = torch.where(mask, torch.ones_like(c), c)
        s = torch.where(mask, torch.zeros_like(s), s)
        mu = torch.atan2(s, c)
        kappa = F.softplus(self.head_kappa(feat)) + 1e-6
        if self.kappa_max is not None:
            kappa = kappa.clamp_max(self.kappa_max)
        return mu, kappa, weight


@torch.no_grad()
def mvm_density_on_grid(mu, kappa, weight, num=360, device=None):
    """Mixture density on [0, 2*pi).  Reference: models/pointnet_pp_mvM.py:130-144."""
    device = device or mu.device
    theta = torch.linspace(0.0, 2 * math.pi, steps=num, device=device, dtype=mu.dtype)[:-1]
    th = theta[None, None, :]
    vm = torch.exp(kappa[..., None] * torch.cos(th - mu[..., None])) / (2 * math.pi * torch.i0(kappa[..., None]))
    p = (weight[..., None] * vm).sum(dim=1)
    p = p / (p.sum(dim=-1, keepdim=True) + 1e-8)
    return theta.squeeze(), p
