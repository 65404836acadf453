"""CPU training step of the reference algorithm (oracle port) - the cpu_baseline / --impl reference
leg of bench.py.  zero_grad -> forward -> loss -> backward -> clip_grad_norm_(1.0) -> Adam, i.e. the
loop body of train_multi_peaks_vonMises_KL.py:221-236 (clip only there), train_single_peak_vonMises_KL.py:80-86
and train_8dir_KL.py:92-97, on the torch-CPU oracle (sa_torch / losses)."""
from __future__ import annotations

import torch

from . import losses, sa_torch


class OracleTrainer:
    def __init__(self, kind: str, state_dict: dict, lr: float = 1e-3, dropout_p: float | None = None, device="cpu"):
        """device = "cpu" (the cpu_baseline / --impl reference legs) or a CUDA device: the same torch-eager restatement of
        the reference step run on the GPU, i.e. what the reference's own scripts do when CUDA is available
        (train_multi_peaks_vonMises_KL.py:27: device = cuda if available)."""
        self.kind = kind
        self.device = torch.device(device)
        self.sd = sa_torch.clone_state(state_dict, requires_grad=True, device=self.device)
        self.params = [v for k, v in self.sd.items() if v.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=lr)
        self.dropout_p = {"mvm": 0.4}.get(kind, 0.5) if dropout_p is None else dropout_p

    def step(self, xyz: torch.Tensor, targets: tuple) -> float:
        B, N, _ = xyz.shape
        self.opt.zero_grad(set_to_none=True)
        # the reference draws its random subsets on the host generator (pointnet_pp_8dir.py:28)
        fps1 = torch.stack([torch.randperm(N)[:128] for _ in range(B)]).to(self.device)
        fps2 = torch.stack([torch.randperm(128)[:32] for _ in range(B)]).to(self.device)
        res = sa_torch.model_forward(self.kind, self.sd, xyz, fps1, fps2, training=True, dropout_p=self.dropout_p)
        if self.kind == "mvm":
            loss = losses.match_loss(res[0], res[1], res[2], targets[0], targets[1]).mean()
        elif self.kind == "vonmises":
            loss = losses.kl_von_mises_single(res[0], res[1], targets[0], targets[1]).mean()
        elif self.kind == "8dir":
            loss = losses.soft_ce(res, targets[0]).mean()
        else:
            loss = sum((r ** 2).sum() for r in (res if isinstance(res, tuple) else (res,)))
        loss.backward()
        if self.kind == "mvm":
            torch.nn.utils.clip_grad_norm_(self.params, 1.0)
        self.opt.step()
        return float(loss)
