"""Data-parallel host logic on CPU: world_size-2 gloo processes (the N>1 path of bench.py / pcoe.dp)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Toy(torch.nn.Module):
    """Stands in for a drop-in model on CPU (the SA layers are CUDA-only): BatchNorm makes the
    replicas' statistics rank-local, exactly like the real models."""

    def __init__(self):
        super().__init__()
        self.fc1 = torch.nn.Linear(6, 16)
        self.bn = torch.nn.BatchNorm1d(16)
        self.fc2 = torch.nn.Linear(16, 3)

    def forward(self, x):
        return self.fc2(torch.relu(self.bn(self.fc1(x))))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pcoe
    torch.manual_seed(100 + rank)                       # different initial weights per rank ...
    model = _Toy()
    engine = pcoe.dp.DataParallel(model)                # ... made identical by the rank-0 broadcast
    w0 = [p.detach().clone() for p in model.parameters()]
    gathered = [torch.zeros_like(w0[0]) for _ in range(world)]
    dist.all_gather(gathered, w0[0])
    assert all(torch.equal(g, gathered[0]) for g in gathered)

    g = torch.Generator().manual_seed(5)
    full = torch.randn(8, 6, generator=g)               # global batch of 8 "clouds"
    lo, hi = pcoe.dp.shard_bounds(8, world, rank)
    assert (hi - lo) == 4
    engine.zero_grad()
    model(full[lo:hi]).pow(2).mean().backward()         # autograd accumulates into the flat buffer
    flat_ptr = engine.grads.flat.data_ptr()
    assert all(p.grad.data_ptr() >= flat_ptr for p in model.parameters())
    local = engine.grads.flat.clone()
    engine.allreduce_grads()
    torch.save({"local": local, "reduced": engine.grads.flat.clone(), "rm": model.bn.running_mean.clone()},
               os.path.join(tmp, f"r{rank}.pt"))
    # second step re-uses the same storage (zero_ in place, no re-allocation)
    engine.zero_grad()
    model(full[lo:hi]).sum().backward()
    assert engine.grads.flat.data_ptr() == flat_ptr and float(engine.grads.flat.abs().sum()) > 0
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{i}.pt") for i in range(2))
    # reduced gradient = mean of the shard gradients, identical on both ranks
    assert torch.allclose(r0["reduced"], (r0["local"] + r1["local"]) / 2, rtol=1e-6, atol=1e-7)
    assert torch.equal(r0["reduced"], r1["reduced"])
    # BatchNorm statistics stay rank-local: the shards differ, so do the running means
    assert not torch.allclose(r0["rm"], r1["rm"])


def test_shard_bounds_and_errors():
    import pcoe
    assert [pcoe.dp.shard_bounds(256, 8, r) for r in (0, 7)] == [(0, 32), (224, 256)]
    with pytest.raises(ValueError):
        pcoe.dp.shard_bounds(10, 4, 0)
    with pytest.raises(ValueError):
        pcoe.dp.FlatGradBuffer(torch.nn.ReLU())


def test_single_process_engine_is_a_no_op_allreduce():
    import pcoe
    m = _Toy()
    e = pcoe.dp.DataParallel(m)
    e.zero_grad()
    m(torch.randn(4, 6)).sum().backward()
    before = e.grads.flat.clone()
    e.allreduce_grads()
    assert torch.equal(before, e.grads.flat) and e.grads.nbytes() == 4 * sum(p.numel() for p in m.parameters())


def test_exchange_choice_and_padded_flat_buffer():
    """exchange='auto' on CPU parameters keeps the torch.distributed path; an unknown name is refused; the flat buffer is
    padded to 16 bytes (the peer-memory kernel moves float4) and the padding never counts as payload."""
    import pytest
    import pcoe
    m = _Toy()
    e = pcoe.dp.DataParallel(m, exchange="auto")
    assert e.peer is None
    with pytest.raises(ValueError):
        pcoe.dp.DataParallel(_Toy(), exchange="mpi")
    total = sum(p.numel() for p in m.parameters())
    assert e.grads.flat.numel() == (total + 3) // 4 * 4 and e.grads.nbytes() == 4 * total
    opt = pcoe.optim.FusedAdam if hasattr(pcoe, "optim") else None
    assert opt is not None


def test_overlapped_exchange_is_armed_once_per_step():
    """a second backward() before allreduce_grads() must not all-reduce the already-summed tail again.
    The hook raises unless the extra micro-batches run under no_sync(); FlatGradBuffer notices detached .grad views."""
    import pytest
    import torch
    import pcoe

    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    eng = pcoe.dp.DataParallel(net)                     # world 1: no process group needed
    eng._late_off, eng._late_work = 4, object()         # as if the tail exchange of this step were in flight
    with pytest.raises(RuntimeError, match="no_sync"):
        eng._reduce_tail_async()
    eng._late_work = None
    with eng.no_sync():
        assert eng._sync is False
        eng._reduce_tail_async()                        # accumulation micro-batch: nothing is exchanged
        assert eng._late_work is None
    assert eng._sync is True
    eng.grads.check_views()
    torch.optim.SGD(net.parameters(), lr=0.1).zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError, match="flat gradient buffer"):
        eng.grads.check_views()
