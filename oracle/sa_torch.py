"""torch-CPU oracle of the set-abstraction layer and the seven model heads.

Functional restatement driven by a ``state_dict`` (the reference's key names), so one checkpoint
feeds the CUDA modules, this oracle and the unmodified reference alike.  Works in fp32 or fp64
(cast the state dict and inputs); autograd supplies the backward oracle.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def gather(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """points (B,N,C), idx (B,...) -> (B,...,C).  Follows index_points, models/base.py:4-18."""
    B = points.shape[0]
    flat = idx.reshape(B, -1)
    out = torch.gather(points, 1, flat.unsqueeze(-1).expand(-1, -1, points.shape[-1]))
    return out.reshape(*idx.shape, points.shape[-1])


def knn_indices(new_xyz: torch.Tensor, xyz: torch.Tensor, k: int) -> torch.Tensor:
    """Follows square_distance + topk, models/base.py:20-35 (same formula: -2ab + a^2 + b^2)."""
    d = -2.0 * (new_xyz @ xyz.transpose(1, 2))
    d = d + (new_xyz ** 2).sum(-1, keepdim=True)
    d = d + (xyz ** 2).sum(-1).unsqueeze(1)
    return d.topk(k, dim=-1, largest=False, sorted=False).indices


def set_abstraction(sd: dict, prefix: str, xyz: torch.Tensor, points, *, group_all: bool, nsample=None,
                    fps_idx=None, group_idx=None, training: bool = True, update_buffers: bool = True,
                    eps: float = 1e-5, momentum: float = 0.1, names=("convs", "bns")):
    """One SA layer.  Follows PointNetSetAbstraction.forward, models/pointnet_pp_8dir.py:21-43.
    ``fps_idx`` (B,S) must be given for a non-global layer (the reference draws it from the CPU
    generator, :28 - see sampling.randperm_subset_replay); ``group_idx`` (B,S,K) overrides the kNN.
    ``names`` = attribute names of the conv / BatchNorm lists (("mlp_convs", "mlp_bns") for PointNet++Demo.py's
    SimpleSetAbstraction, whose MLP / pooling arithmetic :119-127 is the same).
    Returns (new_xyz, new_points, group_idx)."""
    cn, bn_ = names
    B = xyz.shape[0]
    if group_all:
        new_xyz = torch.zeros(B, 1, 3, dtype=xyz.dtype, device=xyz.device)
        rows = xyz.unsqueeze(1)                                                  # :24 absolute xyz
        if points is not None:
            rows = torch.cat([rows, points.unsqueeze(1)], -1)                    # :25
    else:
        new_xyz = gather(xyz, fps_idx)                                           # :29
        if group_idx is None:
            group_idx = knn_indices(new_xyz, xyz, nsample)                       # :30
        rows = gather(xyz, group_idx) - new_xyz.unsqueeze(2)                     # :31-32
        if points is not None:
            rows = torch.cat([rows, gather(points, group_idx)], -1)              # :34-35
    x = rows                                                                     # (B,S,K,C) channels last
    for l in range(3):
        W = sd[f"{prefix}.{cn}.{l}.weight"].reshape(-1, x.shape[-1])            # (Cout,Cin,1,1)
        x = x @ W.t() + sd[f"{prefix}.{cn}.{l}.bias"]                           # 1x1 conv, :41
        rm, rv = sd[f"{prefix}.{bn_}.{l}.running_mean"], sd[f"{prefix}.{bn_}.{l}.running_var"]
        if not update_buffers:
            rm, rv = rm.clone(), rv.clone()
        flat = x.reshape(-1, x.shape[-1])
        flat = F.batch_norm(flat, rm, rv, sd[f"{prefix}.{bn_}.{l}.weight"], sd[f"{prefix}.{bn_}.{l}.bias"],
                            training, momentum, eps)
        if training and update_buffers:
            sd[f"{prefix}.{bn_}.{l}.num_batches_tracked"] += 1
        x = F.relu(flat).reshape(x.shape)
    return new_xyz, x.max(dim=2).values, group_idx                               # :42-43


def ssg_layer(sd: dict, prefix: str, xyz, points, fps_idx, radius: float, nsample: int, training=True,
              update_buffers=True, names=("mlp_convs", "mlp_bns")):
    """SimpleSetAbstraction.forward, PointNet++Demo.py:96-129, channels-last: xyz (B,N,3), points (B,N,D) or None,
    fps_idx (B,S) from the FPS oracle (:106), radius ball query (:109) from oracle.sampling.ball_query.
    Returns (new_xyz (B,S,3), new_points (B,S,C), group_idx)."""
    from . import sampling
    new_xyz = gather(xyz, fps_idx)
    gi = torch.from_numpy(sampling.ball_query(radius, nsample, xyz.detach().float().numpy(), new_xyz.detach().float().numpy()))
    return set_abstraction(sd, prefix, xyz, points, group_all=False, fps_idx=fps_idx, group_idx=gi, training=training,
                           update_buffers=update_buffers, names=names)


def msg_layer(sd: dict, prefix: str, xyz, points, fps_idx, radii, nsamples, training=True, update_buffers=True):
    """Multi-scale grouping: ssg_layer per (radius, nsample) on shared centroids, parameters under
    ``{prefix}.conv_blocks.{i}`` / ``{prefix}.bn_blocks.{i}``, outputs concatenated along channels."""
    outs = []
    for i, (r, k) in enumerate(zip(radii, nsamples)):
        sub = {}
        for key, v in sd.items():
            if key.startswith(f"{prefix}.conv_blocks.{i}."):
                sub["b.mlp_convs." + key[len(f"{prefix}.conv_blocks.{i}."):]] = v
            elif key.startswith(f"{prefix}.bn_blocks.{i}."):
                sub["b.mlp_bns." + key[len(f"{prefix}.bn_blocks.{i}."):]] = v
        new_xyz, o, _ = ssg_layer(sub, "b", xyz, points, fps_idx, r, k, training, update_buffers)
        outs.append(o)
    return new_xyz, torch.cat(outs, dim=-1)


def ssg_cls_forward(sd: dict, x_cf: torch.Tensor, fps1, fps2, training=True, update_buffers=True, normal_channel=True):
    """PointNetPlusPlusCls.forward, PointNet++Demo.py:212-240 (dropout as identity): x (B,C,N) channel-first."""
    xyz = x_cf[:, :3].transpose(1, 2)
    pts = x_cf[:, 3:].transpose(1, 2) if normal_channel else None
    l1_xyz, l1, _ = ssg_layer(sd, "sa1", xyz, pts, fps1, 0.2, 32, training, update_buffers)
    l2_xyz, l2, _ = ssg_layer(sd, "sa2", l1_xyz, l1, fps2, 0.4, 64, training, update_buffers)
    _, l3, _ = set_abstraction(sd, "sa3", l2_xyz, l2, group_all=True, training=training, update_buffers=update_buffers,
                               names=("mlp_convs", "mlp_bns"))
    h = _bn_trunk(sd, l3.reshape(x_cf.shape[0], -1), training, update_buffers)
    return F.log_softmax(h @ sd["fc3.weight"].t() + sd["fc3.bias"], dim=1)


def sa_features(sd: dict, xyz: torch.Tensor, fps1, fps2, training=True, update_buffers=True, record=None):
    """sa1 -> sa2 -> sa3 -> (B,1024) with the constructor constants of pointnet_pp_8dir.py:65-67."""
    l1_xyz, l1, g1 = set_abstraction(sd, "sa1", xyz, None, group_all=False, nsample=32, fps_idx=fps1,
                                     training=training, update_buffers=update_buffers)
    l2_xyz, l2, g2 = set_abstraction(sd, "sa2", l1_xyz, l1, group_all=False, nsample=32, fps_idx=fps2,
                                     training=training, update_buffers=update_buffers)
    _, l3, _ = set_abstraction(sd, "sa3", l2_xyz, l2, group_all=True, training=training,
                               update_buffers=update_buffers)
    if record is not None:
        record.update(l1=l1, l2=l2, l3=l3, g1=g1, g2=g2, l1_xyz=l1_xyz, l2_xyz=l2_xyz)
    return l3.reshape(xyz.shape[0], -1)


def _bn_trunk(sd, feat, training, update_buffers, dropout_p=0.0):
    x = feat
    for fc, bn in (("fc1", "bn1"), ("fc2", "bn2")):
        x = x @ sd[f"{fc}.weight"].t() + sd[f"{fc}.bias"]
        rm, rv = sd[f"{bn}.running_mean"], sd[f"{bn}.running_var"]
        if not update_buffers:
            rm, rv = rm.clone(), rv.clone()
        x = F.relu(F.batch_norm(x, rm, rv, sd[f"{bn}.weight"], sd[f"{bn}.bias"], training, 0.1, 1e-5))
    return F.dropout(x, dropout_p, training) if dropout_p > 0 else x   # parity runs use p=0


def _ln_trunk(sd, feat, training=True, dropout_p=0.0):
    x = feat
    for fc, ln in (("fc1", "ln1"), ("fc2", "ln2")):
        x = x @ sd[f"{fc}.weight"].t() + sd[f"{fc}.bias"]
        x = F.relu(F.layer_norm(x, x.shape[-1:], sd[f"{ln}.weight"], sd[f"{ln}.bias"], 1e-5))
        if dropout_p > 0:
            x = F.dropout(x, dropout_p, training)      # applied after both layers (pointnet_pp_mvM.py:82-83)
    return x


def model_forward(kind: str, sd: dict, xyz: torch.Tensor, fps1, fps2, training=True, update_buffers=True,
                  record=None, kappa_max=80.0, temp=0.7, max_K=4, dropout_p=0.0):
    """Heads (dropout treated as identity):
      'vonmises' pointnet_pp_vonMises.py:26-38 | '8dir' pointnet_pp_8dir.py:76-85 | 'pp' pointnet_pp.py:59-68
      'xyz' Pointnet_pp_xyz.py:68-90 | 'schedmit' Pointnet_pp_xyz_Schedmit.py:68-90 | 'fwd' pointnet_pp_Fwd.py:89-98
      'mvm' pointnet_pp_mvM.py:86-127"""
    feat = sa_features(sd, xyz, fps1, fps2, training, update_buffers, record)
    lin = lambda name, x: x @ sd[f"{name}.weight"].t() + sd[f"{name}.bias"]
    if kind == "mvm":
        h = _ln_trunk(sd, feat, training, dropout_p)
        weight = F.softmax(lin("head_pi", h) / temp, dim=-1)
        mu_raw = lin("head_mu", h).view(-1, max_K, 2)
        unit = F.normalize(mu_raw, dim=-1, eps=1e-4)
        c, s = unit[..., 0], unit[..., 1]
        bad = torch.sqrt(c * c + s * s) < 1e-3
        c = torch.where(bad, torch.ones_like(c), c)
        s = torch.where(bad, torch.zeros_like(s), s)
        kappa = F.softplus(lin("head_kappa", h)) + 1e-6
        if kappa_max is not None:
            kappa = kappa.clamp_max(kappa_max)
        return torch.atan2(s, c), kappa, weight
    h = _bn_trunk(sd, feat, training, update_buffers, dropout_p)
    if kind == "vonmises":
        o = lin("fc3", h)
        return torch.tanh(o[:, 0]) * math.pi, F.softplus(o[:, 1])
    if kind in ("8dir", "pp"):
        return lin("fc3", h)
    if kind == "fwd":
        return F.normalize(lin("fc3", h), dim=1)
    if kind == "xyz":
        return F.normalize(lin("head_x", h), p=2, dim=1), F.normalize(lin("head_y", h), p=2, dim=1)
    if kind == "schedmit":
        return F.normalize(lin("head_y", h), p=2, dim=1), F.normalize(lin("head_z", h), p=2, dim=1)
    raise ValueError(kind)


def clone_state(sd: dict, dtype=None, requires_grad: bool = False, device="cpu") -> dict:
    """Detached copy of a state dict on the CPU - or on `device` for bench.py's eager-CUDA run of the port - (floating
    tensors optionally cast / made leaves)."""
    out = {}
    for k, v in sd.items():
        t = v.detach().cpu().clone().to(device)
        if t.is_floating_point():
            if dtype is not None:
                t = t.to(dtype)
            if requires_grad and "running_" not in k:
                t.requires_grad_(True)
        out[k] = t
    return out
