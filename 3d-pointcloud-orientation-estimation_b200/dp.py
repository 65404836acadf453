"""Data-parallel training over the batch of clouds: one process per GPU, one flat fp32 gradient
buffer, ONE all-reduce per step.

Clouds are independent units for sampling, grouping, the MLP GEMMs, the max-pool and the per-sample
losses, so the batch is sharded with no data-path collective; the only exchange is the gradient
mean.  BatchNorm batch statistics and running buffers stay rank-local (the north star: "allreduce
for the MLP gradients only"), i.e. every rank behaves exactly like the reference run on its shard.

The exchange itself: libpcoe's peer-memory kernel on one NVLink node (PeerExchange, csrc/peer.cu: the flat buffer
lives in symmetric memory, two flag exchanges and one pass over the data per call), or any torch.distributed backend
(NCCL across nodes, gloo in the CPU tests) - ``DataParallel(exchange="auto" | "peer" | "nccl")``.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class FlatGradBuffer:
    """Re-points every ``p.grad`` into one contiguous fp32 buffer so that autograd accumulates in
    place and the whole gradient is reduced by a single collective (5.86-5.88 MB for these models)."""

    def __init__(self, module: torch.nn.Module, alloc=None):
        """``alloc(n_floats, device) -> tensor`` overrides the allocation (pcoe.dp.PeerExchange: symmetric memory that
        every rank of the node maps); the buffer is padded to a multiple of 4 floats for 16-byte accesses."""
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, total = self.params[0].device, sum(p.numel() for p in self.params)
        self.total = total
        padded = (total + 3) // 4 * 4
        self.flat = torch.zeros(padded, dtype=torch.float32, device=dev) if alloc is None else alloc(padded, dev)
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
        """Bytes of gradient payload (the ≤ 3 floats of 16-byte padding at the end stay zero and are not counted)."""
        return self.total * 4


class PeerExchange:
    """Sum-all-reduce of (a slice of) the flat gradient buffer over NVLink peer memory: libpcoe's two-shot kernel
    (csrc/peer.cu) on torch's symmetric memory (every rank maps every peer's buffer and a small signal pad).  One node,
    <= 8 ranks, CUDA only; CUDA-graph capturable (flag epochs live on the device)."""

    def __init__(self, group=None, max_ctas: int | None = None, multicast: bool | None = None):
        """max_ctas: grid cap of the exchange kernel (default: 16 with an NVSwitch multicast mapping - one load and one
        store per element saturate early - else 64).  multicast: use the multicast mapping when the fabric offers one
        (default: with more than 2 ranks; PCOE_PEER_MULTICAST=0 / 1 overrides)."""
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self._C, self._symm = C, symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PeerExchange: at most 8 ranks (one NVSwitch node)")
        self.max_ctas = max_ctas
        if multicast is None:
            # with one peer the switch-side reduction saves nothing (one remote load either way) and the plain P2P loop
            # with 64 CTAs is the faster one (23 vs 28-42 us for the whole buffer, profiles/r02_scaling.txt)
            env = os.environ.get("PCOE_PEER_MULTICAST")
            multicast = (env != "0") if env is not None else self.world > 2
        self._want_mc, self.mc_ptr = multicast, 0
        self.flat = None
        self.stream = torch.cuda.Stream()

    def alloc(self, n: int, device) -> torch.Tensor:
        symm = self._symm
        self.flat = symm.empty(n, dtype=torch.float32, device=device)
        self.flat.zero_()
        self._hb = symm.rendezvous(self.flat, self.group)
        self.pad = symm.empty(64, dtype=torch.int32, device=device)
        self.pad.zero_()
        self._hp = symm.rendezvous(self.pad, self.group)
        self.ctl = torch.zeros(16, dtype=torch.int32, device=device)
        C = self._C
        self._bufs = (C.c_void_p * self.world)(*[int(p) for p in self._hb.buffer_ptrs])
        self._pads = (C.c_void_p * self.world)(*[int(p) for p in self._hp.buffer_ptrs])
        if self._want_mc:                              # NVSwitch multicast mapping, when the fabric offers one
            try:
                self.mc_ptr = int(getattr(self._hb, "multicast_ptr", 0) or 0)
            except Exception:
                self.mc_ptr = 0
        if self.max_ctas is None:
            self.max_ctas = 16 if self.mc_ptr else 64
        torch.cuda.synchronize()
        dist.barrier(self.group)                       # every rank's pad / buffer is zeroed before the first flag lands
        return self.flat

    def all_reduce_(self, lo: int, hi: int, stream=None) -> None:
        """flat[lo:hi] <- sum over ranks, enqueued on `stream` (default: the current stream).  lo, hi multiples of 4."""
        from . import _lib
        st = torch.cuda.current_stream() if stream is None else stream
        _lib.check(_lib.load().pcoe_peer_allreduce_f32(self._bufs, self._pads, self.mc_ptr or None, self.rank, self.world, lo, hi - lo,
                                                       self.ctl.data_ptr(), self.max_ctas, st.cuda_stream))


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

    def __init__(self, module: torch.nn.Module, process_group=None, broadcast_buffers: bool = True, overlap: bool = True,
                 exchange: str = "auto"):
        """exchange: 'nccl' (torch.distributed all-reduce: any backend), 'peer' (libpcoe's NVLink peer-memory kernel,
        PeerExchange) or 'auto' ('peer' when it can be set up - CUDA, <= 8 ranks, symmetric memory - else 'nccl')."""
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
        if exchange not in ("nccl", "peer", "auto"):
            raise ValueError(f"pcoe.dp: exchange={exchange!r} (expected 'nccl', 'peer' or 'auto')")
        self.peer = None
        if self.world > 1 and exchange in ("peer", "auto") and next(module.parameters()).is_cuda:
            err = None
            try:
                self.peer = PeerExchange(process_group)
                self.grads = FlatGradBuffer(module, alloc=self.peer.alloc)
            except Exception as e:                                # no symmetric memory on this system / across nodes
                err = e
            # the choice is collective: one rank without a mapping sends every rank to the torch.distributed path
            ok = torch.tensor([0 if err is not None else 1], device=next(module.parameters()).device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
            if int(ok) == 0:
                if exchange == "peer":
                    raise RuntimeError(f"pcoe.dp: peer-memory exchange unavailable on at least one rank ({err})")
                import warnings
                warnings.warn(f"pcoe.dp: peer-memory exchange unavailable ({err}); using the torch.distributed all-reduce")
                self.peer = None
        if self.peer is None:
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
                off = (off + 3) // 4 * 4     # 16-byte aligned boundary, rounded UP: the tail bucket holds only final gradients
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
            if self.peer is not None:
                # side stream: the tail's exchange runs beside the remaining backward kernels
                side = self.peer.stream
                side.wait_stream(torch.cuda.current_stream())
                self.peer.all_reduce_(self._late_off, self.grads.flat.numel(), side)
                self._late_work = side
            else:
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
            if self.peer is not None:
                cur = torch.cuda.current_stream()
                if self._late_work is not None:               # head bucket behind the tail on the side stream, then join
                    side = self._late_work
                    side.wait_stream(cur)
                    self.peer.all_reduce_(0, self._late_off, side)
                    cur.wait_stream(side)
                    self._late_work = None
                else:
                    self.peer.all_reduce_(0, self.grads.flat.numel(), cur)
            elif self._late_work is not None:
                dist.all_reduce(self.grads.flat[:self._late_off], op=dist.ReduceOp.SUM, group=self.group)
                self._late_work.wait()
                self._late_work = None
            else:
                dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.group)
            if not self.defer_scale:
                self.grads.flat.mul_(1.0 / self.world)

    def __call__(self, *a, **k):
        return self.module(*a, **k)
