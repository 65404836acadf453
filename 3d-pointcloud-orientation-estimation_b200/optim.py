"""Fused clip + Adam (+ zero_grad) over flat buffers: two kernel launches per optimisation step.

Replaces, in the reference's training loops (train_multi_peaks_vonMises_KL.py:182,221,235-236;
train_single_peak_vonMises_KL.py:68,81,90; train_8dir_KL.py:72,93,97)::

    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
    ...
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)     # mvM script only
    optimizer.step()

with::

    optimizer = pcoe.optim.FusedAdam(model, lr=1e-3, max_grad_norm=1.0)
    ...
    optimizer.step()

Every parameter's ``.data`` becomes a view into one flat fp32 buffer and every ``.grad`` a view into the flat
gradient buffer of ``pcoe.dp.FlatGradBuffer`` (shared with the data-parallel all-reduce), so the update is
``pcoe_adam_step`` over four flat arrays.  The step counter and the gradient norm live on the device:
the step is CUDA-graph capturable.  ``state_dict()`` / ``load_state_dict()`` use ``torch.optim.Adam``'s
layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so checkpoints are interchangeable.
"""
from __future__ import annotations

import torch

from . import _lib
from .dp import DataParallel, FlatGradBuffer


class FusedAdam:
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_grad_norm: float | None = None, zero_grad_in_step: bool = False):
        """`model`: an nn.Module, a pcoe.dp.DataParallel engine or a pcoe.dp.FlatGradBuffer.
        `max_grad_norm`: fold ``clip_grad_norm_(params, max_grad_norm)`` into the step.
        `zero_grad_in_step`: clear the gradients while they are consumed (then skip ``zero_grad()``)."""
        self.grad_scale = 1.0
        if isinstance(model, DataParallel):
            grads = model.grads
            model.defer_scale = True               # the 1/world factor of the gradient mean is applied in the step kernel
            self.grad_scale = 1.0 / model.world
        elif isinstance(model, FlatGradBuffer):
            grads = model
        else:
            grads = FlatGradBuffer(model)
        self.grads = grads
        self.params = grads.params
        flat_g = grads.flat
        if not flat_g.is_cuda:
            raise RuntimeError("pcoe.optim.FusedAdam runs on CUDA only (no CPU fallback)")
        dev = flat_g.device
        self.flat_p = torch.zeros(flat_g.numel(), dtype=torch.float32, device=dev)   # (the gradient buffer may be padded / symmetric)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                view = self.flat_p[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view                      # parameters now live in the flat buffer
                off += n
        self.exp_avg = torch.zeros(flat_g.numel(), dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(flat_g.numel(), dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)   # ||g|| before clipping, last step
        self._ws = torch.zeros(_lib.load().pcoe_adam_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.defaults = dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps),
                             weight_decay=float(weight_decay))
        self.max_grad_norm = max_grad_norm
        self.zero_grad_in_step = bool(zero_grad_in_step)
        self.fused_clip = max_grad_norm is not None     # GraphedTrainStep: do not clip a second time

    # torch.optim.Optimizer look-alikes --------------------------------------------------------
    @property
    def param_groups(self):
        return [dict(self.defaults, params=self.params)]

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.grads.zero_()

    @torch.no_grad()
    def step(self) -> None:
        d = self.defaults
        _lib.check(_lib.load().pcoe_adam_step(
            self.flat_p.data_ptr(), self.grads.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.flat_p.numel(), d["lr"], d["betas"][0], d["betas"][1], d["eps"], d["weight_decay"],
            float(self.max_grad_norm) if self.max_grad_norm is not None else 0.0, self.grad_scale,
            int(self.zero_grad_in_step),
            self.step_dev.data_ptr(), self.grad_norm.data_ptr(), self._ws.data_ptr(),
            torch.cuda.current_stream().cuda_stream))

    def state_dict(self) -> dict:
        state, off = {}, 0
        step = self.step_dev.to(torch.float32).cpu().reshape(())
        for i, p in enumerate(self.params):
            n = p.numel()
            state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
            off += n
        group = dict(lr=self.defaults["lr"], betas=self.defaults["betas"], eps=self.defaults["eps"],
                     weight_decay=self.defaults["weight_decay"], amsgrad=False, maximize=False, foreach=None,
                     capturable=True, differentiable=False, fused=True, decoupled_weight_decay=False,
                     params=list(range(len(self.params))))
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        g = sd["param_groups"][0]
        self.defaults.update(lr=float(g["lr"]), betas=(float(g["betas"][0]), float(g["betas"][1])),
                             eps=float(g["eps"]), weight_decay=float(g.get("weight_decay", 0.0)))
        off, step = 0, 0
        for i, p in enumerate(self.params):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                step = int(float(st["step"]))
            off += n
        self.step_dev.fill_(step)
