"""CUDA-graph capture of a whole training step ("CUDA-graph-captured steps").

A step of the drop-in models is ~50 libpcoe launches plus a handful of small torch launches; at 64 clouds per
GPU the CPU cannot enqueue them as fast as the B200 executes them.  ``GraphedTrainStep`` captures

    zero_grad -> forward -> loss -> backward -> [gradient all-reduce] -> [clip_grad_norm_] -> optimizer.step

once (after warm-up on a side stream) and replays it per batch; inputs are copied into static device
buffers first.  Requirements: an optimizer constructed with ``capturable=True``, fixed batch shape.

Sampling.  ``sampler="randperm_device"`` / ``"fps"`` draw inside the graph.  The reference's own sampler
(``"randperm_host"``: ``torch.randperm`` on the CPU generator, models/pointnet_pp_8dir.py:28) cannot run inside a
graph, so its subsets become graph INPUTS: before every replay the step replays the reference's draws on the host -
B x randperm(N) for sa1, then B x randperm(128) for sa2, on torch's global CPU generator, bit-identical
(``pcoe.ops.host_randperm_subsets``, 0.14 ms for 64 clouds) - into pinned buffers and uploads them (40 KB) into the
layers' static index buffers.  A seeded run therefore trains on exactly the subsets of the seeded reference run.
"""
from __future__ import annotations

import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, model, loss_fn, optimizer, example_xyz: torch.Tensor, example_targets: tuple,
                 clip_norm: float | None = None, engine=None, warmup: int = 3):
        self.model, self.loss_fn, self.opt, self.clip, self.engine = model, loss_fn, optimizer, clip_norm, engine
        self.xyz = example_xyz.clone()
        self.targets = tuple(t.clone() for t in example_targets)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.loss = None
        # layers whose subsets are host-replayed graph inputs, in forward (= the reference's draw) order
        self._fed = [m for m in model.modules()
                     if getattr(m, "sampler", None) == "randperm_host" and not getattr(m, "group_all", False)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._step()                       # eager: also records every fed layer's (B, N, npoint)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._ring, self._slot = [], 0
        if self._fed:
            dev = self.xyz.device
            for m in self._fed:
                B, N, S = m._last_draw_shape
                m._static_idx = torch.zeros(B, S, dtype=torch.int32, device=dev)
            # ring of pinned host buffers (an async upload may still be reading slot k when slot k+1 is drawn)
            for _ in range(4):
                self._ring.append(([torch.empty(m._last_draw_shape[0], m._last_draw_shape[2], dtype=torch.int32).pin_memory()
                                    for m in self._fed], torch.cuda.Event()))
            self._draw_into([m._static_idx for m in self._fed])
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()

    def _draw_into(self, dst: list) -> None:
        """Replay the reference's subset draws of ONE forward on the host generator and upload them to `dst`."""
        bufs, ev = self._ring[self._slot]
        self._slot = (self._slot + 1) % len(self._ring)
        ev.synchronize()                           # the upload that last used this slot has finished (no-op when fresh)
        for m, h, d in zip(self._fed, bufs, dst):
            B, N, S = m._last_draw_shape
            ops.host_randperm_subsets(B, N, S, out=h)
            d.copy_(h, non_blocking=True)
        ev.record()

    def release(self) -> None:
        """Give the fed layers their own (eager) host draw back."""
        for m in self._fed:
            m._static_idx = None

    def _step(self):
        if getattr(self.opt, "zero_grad_in_step", False):
            pass                                   # pcoe.optim.FusedAdam clears the gradients as it consumes them
        elif self.engine is not None:
            self.engine.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=False)
        out = self.model(self.xyz)
        loss = self.loss_fn(out, *self.targets)
        loss.backward()
        if self.engine is not None:
            self.engine.allreduce_grads()
        if self.clip is not None and not getattr(self.opt, "fused_clip", False):
            torch.nn.utils.clip_grad_norm_(self.params, self.clip, foreach=True)
        self.opt.step()
        return loss.detach()

    def __call__(self, xyz: torch.Tensor, *targets: torch.Tensor) -> torch.Tensor:
        """Copies the batch into the static buffers (H2D if `xyz` is a pinned host tensor), replays the
        step, returns the (static) loss tensor."""
        self.xyz.copy_(xyz, non_blocking=True)
        for dst, src in zip(self.targets, targets):
            dst.copy_(src, non_blocking=True)
        if self._fed:
            self._draw_into([m._static_idx for m in self._fed])
        self.graph.replay()
        return self.loss

    # ---- double-buffered input feed: the H2D copy of batch i+1 overlaps step i ------------------------------
    def prefetch(self, xyz: torch.Tensor, *targets: torch.Tensor) -> None:
        """Start copying the NEXT batch (pinned host tensors) into staging device buffers on a copy stream; the
        copy runs beside the step that is executing.  Consume it with ``step_prefetched()``."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = (torch.empty_like(self.xyz),) + tuple(torch.empty_like(t) for t in self.targets)
            self._stage_idx = [torch.empty_like(m._static_idx) for m in self._fed]
            self._consumed = None                  # event: the staging buffers have been read by the last D2D copy
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)          # only the previous D2D copy, NOT the step that follows it
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stage, (xyz,) + tuple(targets)):
                dst.copy_(src, non_blocking=True)
            if self._fed:                          # this batch's subsets: drawn now (host generator order = step order)
                self._draw_into(self._stage_idx)
        self._staged = torch.cuda.Event()
        self._staged.record(cs)

    def step_prefetched(self) -> torch.Tensor:
        """Replay the step on the batch staged by ``prefetch()`` (device-to-device copy into the static buffers)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for dst, src in zip((self.xyz,) + tuple(self.targets), self._stage):
            dst.copy_(src, non_blocking=True)
        for m, src in zip(self._fed, self._stage_idx):
            m._static_idx.copy_(src, non_blocking=True)
        self._consumed = torch.cuda.Event()
        self._consumed.record(cur)
        self.graph.replay()
        return self.loss
