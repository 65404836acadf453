"""torch-CPU oracle of the vanilla PointNet path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Functional restatement of models/pointnet.py:6-129 driven by the reference's ``state_dict`` keys
(``encoder.stn.conv1.weight`` ...).  Inference (eval-mode BatchNorm, dropout = identity) is what SURVEY 8(f1) / BASELINE
configs[4] name; ``training=True`` uses batch statistics without touching the buffers (used only to cross-check the
restatement).  Pinned against the unmodified reference by oracle/make_golden_r2.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bn(sd, name, x, training):
    """BatchNorm1d on (B,C,N) or (B,C) input."""
    return F.batch_norm(x, sd[f"{name}.running_mean"].clone(), sd[f"{name}.running_var"].clone(), sd[f"{name}.weight"],
                        sd[f"{name}.bias"], training, 0.1, 1e-5)


def _conv(sd, name, x):
    """Conv1d with kernel 1 on (B,C,N)."""
    W = sd[f"{name}.weight"].squeeze(-1)
    return torch.einsum("oc,bcn->bon", W, x) + sd[f"{name}.bias"][None, :, None]


def _lin(sd, name, x):
    return x @ sd[f"{name}.weight"].t() + sd[f"{name}.bias"]


def stn(sd, prefix: str, x, k: int, training=False):
    """STN3d / STNkd forward, models/pointnet.py:22-34 / 53-65: x (B,k,N) -> (B,k,k)."""
    B = x.shape[0]
    h = F.relu(_bn(sd, f"{prefix}.bn1", _conv(sd, f"{prefix}.conv1", x), training))       # :24 / :55
    h = F.relu(_bn(sd, f"{prefix}.bn2", _conv(sd, f"{prefix}.conv2", h), training))       # :25
    h = F.relu(_bn(sd, f"{prefix}.bn3", _conv(sd, f"{prefix}.conv3", h), training))       # :26
    h = h.max(dim=2).values                                                              # :27-28
    h = F.relu(_bn(sd, f"{prefix}.bn4", _lin(sd, f"{prefix}.fc1", h), training))          # :29
    h = F.relu(_bn(sd, f"{prefix}.bn5", _lin(sd, f"{prefix}.fc2", h), training))          # :30
    h = _lin(sd, f"{prefix}.fc3", h)                                                      # :31
    h = h + torch.eye(k, dtype=h.dtype).reshape(1, k * k)                                 # :32-33
    return h.reshape(B, k, k)


def encoder(sd, x_cf, feature_transform=True, training=False, prefix="encoder"):
    """PointNetEncoder.forward with global_feat=True, models/pointnet.py:82-106: x (B,3,N) ->
    (global feature (B,1024), trans (B,3,3), trans_feat (B,64,64) or None)."""
    trans = stn(sd, f"{prefix}.stn", x_cf, 3, training)                                   # :84
    x = torch.bmm(x_cf.transpose(2, 1), trans).transpose(2, 1)                            # :85-92
    x = F.relu(_bn(sd, f"{prefix}.bn1", _conv(sd, f"{prefix}.conv1", x), training))       # :93
    trans_feat = None
    if feature_transform:
        trans_feat = stn(sd, f"{prefix}.fstn", x, 64, training)                           # :95
        x = torch.bmm(x.transpose(2, 1), trans_feat).transpose(2, 1)                      # :96-98
    x = F.relu(_bn(sd, f"{prefix}.bn2", _conv(sd, f"{prefix}.conv2", x), training))       # :102
    x = _bn(sd, f"{prefix}.bn3", _conv(sd, f"{prefix}.conv3", x), training)               # :103 (no ReLU)
    return x.max(dim=2).values, trans, trans_feat                                        # :104-107


def pointnet_forward(sd, x, feature_transform=True, training=False):
    """PointNet.forward, models/pointnet.py:123-129: x (B,N,3) or (B,3,N) -> (out (B,3), gfeat, trans, trans_feat).
    Dropout (:128) is the identity here (inference)."""
    if x.dim() == 3 and x.shape[2] in (3, 6):
        x = x.transpose(1, 2)                                                            # :124-125
    g, trans, trans_feat = encoder(sd, x, feature_transform, training)
    h = F.relu(_bn(sd, "bn1", _lin(sd, "fc1", g), training))                              # :127
    h = F.relu(_bn(sd, "bn2", _lin(sd, "fc2", h), training))                              # :128
    return _lin(sd, "fc3", h), g, trans, trans_feat
