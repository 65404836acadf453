"""Micro-benchmark of the gradient exchange alone: NCCL all_reduce vs libpcoe's peer-memory kernel (csrc/peer.cu), on the
bucket sizes of the c2 step.  Launch with torchrun on N GPUs of one node; rank 0 prints one line per (size, method)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcoe  # noqa: E402


def timed(fn, iters=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    n = 1_480_000
    for ctas, mc in ((16, False), (32, False), (64, False), (128, False), (16, True), (32, True), (64, True)):
        px = pcoe.dp.PeerExchange(max_ctas=ctas, multicast=mc)
        flat = px.alloc(n, torch.device("cuda"))
        ref = torch.zeros(n, device="cuda")
        for size in (80_000, 740_000, n):
            if mc and not px.mc_ptr:
                continue
            flat.fill_(rank + 1.0)
            torch.cuda.synchronize(); dist.barrier()
            px.all_reduce_(0, size)
            torch.cuda.synchronize(); dist.barrier()
            ok = bool((flat[:size] == world * (world + 1) / 2).all()) and bool((flat[size:] == rank + 1.0).all())
            t_peer = timed(lambda: px.all_reduce_(0, size))
            t_nccl = timed(lambda: dist.all_reduce(ref[:size])) if (ctas, mc) == (16, False) else float("nan")
            if rank == 0:
                print(f"world {world} ctas {ctas:3d} multicast {int(px.mc_ptr != 0)} floats {size:8d}: peer {t_peer:7.1f} us  nccl {t_nccl:7.1f} us  correct {ok}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
