"""Data-parallel training over the batch of clouds: one process per GPU, one flat fp32 gradient
buffer, ONE all-reduce per step.

Clouds are independent units for sampling, grouping, the MLP GEMMs, the max-pool and the per-sample
losses, so the batch is sharded with no data-path collective; the only exchange is the gradient
mean.  BatchNorm batch statistics and running buffers stay rank-local (the north star: "allreduce
for the MLP gradients only"), i.e. every rank behaves exactly like the reference run on its shard.

Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradBuffer:
    """Re-points every ``p.grad`` into one contiguous fp32 buffer so that autograd accumulates in
    place and the whole gradient is reduced by a single collective (5.86-5.88 MB for these models)."""

    def __init__(self, module: torch.nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, total = self.params[0].device, sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatGradBuffer expects fp32 parameters on one device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        # set-abstraction layers add their parameter gradients into these views inside pcoe_sa_backward
        # (pcoe_sa_grads.accumulate) instead of returning fresh tensors for autograd to add
        for m in module.modules():
            if hasattr(m, "direct_grad_accumulation"):
                m.direct_grad_accumulation = True

    def zero_(self) -> None:
        self.flat.zero_()

    def check_views(self) -> None:
        """Every ``p.grad`` must still be a view into the flat buffer.  ``optimizer.zero_grad()`` with its default
        ``set_to_none=True`` detaches them: autograd would then allocate fresh gradients and the all-reduce would
        silently exchange a stale buffer.  Raises instead (use ``engine.zero_grad()`` / ``set_to_none=False``)."""
        lo = self.flat.data_ptr()
        hi = lo + self.flat.numel() * 4
        off = 0
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != lo + off * 4 or not (lo <= g.data_ptr() < hi):
                raise RuntimeError(
                    "pcoe.dp: a parameter's .grad no longer points into the flat gradient buffer (was "
                    "optimizer.zero_grad(set_to_none=True) called?); call engine.zero_grad() or "
                    "optimizer.zero_grad(set_to_none=False) instead")
            off += p.numel()

    def nbytes(self) -> int:
        return self.flat.numel() * 4


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Even split of `total` clouds: rank r owns [r*total/world, (r+1)*total/world)."""
    if total % world:
        raise ValueError(f"batch of {total} clouds does not split evenly over {world} ranks")
    per = total // world
    return rank * per, (rank + 1) * per


class DataParallel:
    """Minimal DP engine around a drop-in model: broadcast initial parameters from rank 0, keep the
    gradients in a FlatGradBuffer, average them after backward.

    The exchange is split in two so that most of it overlaps the backward pass: as soon as the
    LAST set-abstraction layer (sa3: 727 k of the 1.47 M parameters; the trunk and heads behind it are already
    done by then) has finished its backward, the tail of the flat buffer - sa3 + trunk + heads, 94 % of the bytes -
    is all-reduced asynchronously while sa2 / sa1 run their backward; ``allreduce_grads()`` then reduces the small
    head of the buffer and joins.  Both collectives are captured by ``pcoe.GraphedTrainStep``."""

    def __init__(self, module: torch.nn.Module, process_group=None, broadcast_buffers: bool = True, overlap: bool = True):
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.defer_scale = False       # set by pcoe.optim.FusedAdam: 1/world is applied inside the optimizer kernel
        if self.world > 1:
            with torch.no_grad():
                for p in module.parameters():
                    dist.broadcast(p.data, src=0, group=process_group)
                if broadcast_buffers:
                    for b in module.buffers():
                        dist.broadcast(b.data, src=0, group=process_group)
        self.grads = FlatGradBuffer(module)
        self._late_off, self._late_work = None, None
        self._sync = True              # False inside no_sync(): micro-batch gradients accumulate locally
        if self.world > 1 and overlap:
            sa = [m for m in module.modules() if hasattr(m, "direct_grad_accumulation") and list(m.parameters())]
            if sa:
                last = sa[-1]
                first_param = next(last.parameters())
                off = 0
                for p in self.grads.params:
                    if p is first_param:
                        break
                    off += p.numel()
                if 0 < off < self.grads.flat.numel():
                    self._late_off = off
                    last._after_backward = self._reduce_tail_async

    def _reduce_tail_async(self) -> None:
        """Called by the last SA layer at the end of its backward: everything from its first parameter to the end
        of the flat buffer is final (autograd accumulates the trunk / head gradients before this node runs)."""
        if self._late_work is not None:
            # a second backward() before allreduce_grads() (gradient accumulation without no_sync, or an exception
            # between the two): reducing the already-summed tail again would count the first micro-batch `world` times
            raise RuntimeError("pcoe.dp: backward() ran twice before allreduce_grads(); wrap the extra micro-batches "
                               "in `with engine.no_sync():` (gradient accumulation)")
        if self._sync:
            self._late_work = dist.all_reduce(self.grads.flat[self._late_off:], op=dist.ReduceOp.SUM, group=self.group,
                                              async_op=True)

    def zero_grad(self) -> None:
        self.grads.zero_()

    def no_sync(self):
        """Context manager for gradient accumulation: backward() inside it adds into the flat buffer without
        starting the overlapped exchange; the first backward() outside it (followed by ``allreduce_grads()``)
        reduces the accumulated sum once."""
        engine = self

        class _NoSync:
            def __enter__(self_inner):
                self_inner.prev, engine._sync = engine._sync, False

            def __exit__(self_inner, *exc):
                engine._sync = self_inner.prev
                return False
        return _NoSync()

    def allreduce_grads(self) -> None:
        """grad <- mean over ranks (sum all-reduce of the flat buffer; scaled by 1/world here unless a
        pcoe.optim.FusedAdam built on this engine applies the factor in its step kernel)."""
        if self.world > 1:
            self.grads.check_views()
            if self._late_work is not None:
                dist.all_reduce(self.grads.flat[:self._late_off], op=dist.ReduceOp.SUM, group=self.group)
                self._late_work.wait()
                self._late_work = None
            else:
                dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.group)
            if not self.defer_scale:
                self.grads.flat.mul_(1.0 / self.world)

    def __call__(self, *a, **k):
        return self.module(*a, **k)
