"""Worker of test_dp_gpu.py (launched with torch.distributed.run, one process per GPU; world size 1 is allowed):
libpcoe's peer-memory all-reduce against torch.distributed's, on integer-valued floats so that every sum is exact."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcoe  # noqa: E402


def main() -> int:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    bad = 0
    for multicast in (False, True):
        px = pcoe.dp.PeerExchange(max_ctas=16, multicast=multicast)
        n = 100_000
        flat = px.alloc(n, dev)
        if multicast and not px.mc_ptr:
            continue
        g = torch.Generator(device="cpu").manual_seed(7 + rank)
        # (lo, hi) slices: whole buffer, unaligned-to-slice sizes, a 4-float sliver, an empty range
        for it, (lo, hi) in enumerate([(0, n), (4, 99_996), (1000, 1004), (64, 64), (0, 52), (50_000, n)] * 2):
            mine = torch.randint(-1000, 1000, (n,), generator=g).float().to(dev)
            want = mine.clone()
            dist.all_reduce(want[lo:hi])
            flat.copy_(mine)
            torch.cuda.synchronize(); dist.barrier()
            px.all_reduce_(lo, hi)
            torch.cuda.synchronize(); dist.barrier()
            if not torch.equal(flat, want):
                bad += 1
                print(f"rank {rank} multicast {multicast} slice [{lo}, {hi}): mismatch, max |d| = {float((flat - want).abs().max())}", flush=True)
        # CUDA-graph replay: the epoch lives on the device
        mine = torch.randint(-1000, 1000, (n,), generator=g).float().to(dev)
        want = mine.clone()
        dist.all_reduce(want)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            flat.copy_(mine)
            torch.cuda.synchronize(); dist.barrier()
            with torch.cuda.graph(graph, stream=s):
                px.all_reduce_(0, n)
        for _ in range(3):
            flat.copy_(mine)
            torch.cuda.synchronize(); dist.barrier()
            graph.replay()
            torch.cuda.synchronize(); dist.barrier()
            if not torch.equal(flat, want):
                bad += 1
                print(f"rank {rank} multicast {multicast}: graph replay mismatch", flush=True)
        del graph
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("peer_worker:", "PASS" if int(t) == 0 else f"FAIL ({int(t)})", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(t) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
