#!/usr/bin/env python
"""bench.py - training throughput of the PointNet++ set-abstraction hot path on B200.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this repo's CUDA path
    python bench.py --impl reference --steps 5 --warmup 1     # the reference algorithm on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one optimisation step of the reference's training loop on one synthetic batch
(zero_grad -> forward -> loss -> backward -> [grad all-reduce] -> clip_grad_norm_ (mvM only) -> Adam),
train_multi_peaks_vonMises_KL.py:221-236.  Workload at N=1 = BASELINE.json configs[1]
(PointNetPPMvM, 64 clouds x 1024 points per GPU); weak scaling over GPUs.

One JSON line on stdout (rank 0).  Both headline numbers run the SAME captured step (pcoe.GraphedTrainStep) in the
`bf16x3` precision mode (split-operand tcgen05 kernels, fp32-class accuracy: the mode the T3 parity tests gate) with the
REFERENCE's sampler: the random subsets of every step are replayed on torch's CPU generator (bit-identical to
`torch.randperm`, models/pointnet_pp_8dir.py:28) and uploaded as graph inputs (40 KB per step).
`value`: xyz + targets resident in HBM, CUDA-event timed per step, L2 flushed between steps.
`e2e`: xyz + targets in pinned HOST memory, copied H2D inside the timed region every step (next batch's copy
overlapped with the running step), loss read back D2H every step, wall-clock timed.
`modes`: the same step in the other precision modes (plain bf16 = throughput mode with a stated tolerance; fp32 =
CUDA-core kernels), so all three are on record.
N > 1: the gradient exchange inside the captured step is libpcoe's NVLink peer-memory all-reduce (csrc/peer.cu) when the
ranks can map each other's memory, else torch.distributed's (`--exchange auto|peer|nccl`; the `grad_exchange` key says
which one ran).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {  # name -> (kind, model class, clouds per GPU, points)
    "c1": ("vonmises", "PointNetPPVonMises", 16, 1024),
    "c2": ("mvm", "PointNetPPMvM", 64, 1024),
    "c3": ("8dir", "PointNetPP8Dir", 32, 2048),
    "c4": ("xyz", "PointNetPPXYZ", 32, 8192),
    "c5": ("pointnet", "PointNet", 1024, 1024),       # inference sweep B = 1 .. 1024 (BASELINE configs[4])
}
WORKLOAD_NAMES = {
    "c1": "PointNet++ SSG single-peak von Mises KL head, 16 x 1024 pts per GPU, train step",
    "c2": "PointNet++ multi-peak mvM KL head (pointnet_pp_mvM.py), 64 x 1024 pts per GPU, train step",
    "c3": "PointNet++ 8-direction head, 32 x 2048 pts per GPU, train step",
    "c4": "Pointnet_pp_xyz at 8192 pts/cloud, 32 clouds per GPU, train step",
    "c5": "PointNet vanilla (models/pointnet.py) inference sweep batch 1-1024 x 1024 pts, eval forward",
}


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=float(p["hbm_gbs"]), tensor=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json; bf16 sustained figure: kernel timed inside a long step)")
    except Exception:
        return dict(hbm=6650.0, tensor=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if not (t0 - 0.1 <= ts <= t1 + 0.3):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def sa_shapes(B: int, N: int):
    """(rows M, Cin, C1, C2, C3, N_in, S, K) of the three SA layers (pointnet_pp_8dir.py:65-67)."""
    return [(B * 128 * 32, 3, 64, 64, 128, N, 128, 32), (B * 32 * 32, 131, 128, 128, 256, 128, 32, 32),
            (B * 32, 259, 256, 512, 1024, 32, 1, 32)]


def kernel_work(B: int, N: int, precision: str = "bf16") -> dict:
    """Algorithmic work per STEP of every libpcoe kernel name: name -> (FLOPs, bytes).
    FLOPs = 2*M*Cin*Cout per GEMM (no padding / recompute).  Bytes = compulsory HBM traffic of the launch:
    every input it must read once and every output it must write once (bf16 activations, fp32 sources,
    int32 indices; weights and per-channel vectors are negligible and left out).  The roofline that binds
    a kernel is whichever of FLOPs / tensor peak and bytes / HBM peak takes longer."""
    sh = sa_shapes(B, N)
    w = {}
    esz = 2.0 if precision == "bf16" else 4.0                         # bf16x3 / fp32 store fp32 activations
    for li, s in enumerate(sh):
        M, c, Nin, S = s[0], s[1:5], s[5], s[6]
        G = M // 32
        g = lambda a, b: 2.0 * M * c[a] * c[b]
        act = lambda k: esz * M * c[k]                                # one activation tensor [M x C_k]
        src = 4.0 * M + 4.0 * (M // (S * 32)) * Nin * c[0]             # neighbour indices + the gathered fp32 source, once
        pool = 10.0 * G * c[3]                                        # ymax, ymin (f32) + amax, amin (u8) per group
        t = f"sa{li + 1}_"
        dg1 = 2.0 * M * (c[0] - 3) * c[1]                             # no data gradient into xyz
        scat = 4.0 * (M // (S * 32)) * Nin * (c[0] - 3)               # grad_feats written once
        # SA1 / SA2 (v4 kernels): the last layer's pre-activations y3 are never stored - the forward accumulates the
        # Gram matrix of its input and the backward uses the sparse max-pool routing - so
        # act(3) is neither written by fwd_l3 nor read by bwd_l3.  SA3 (v5 kernels) still stores y3.
        y3 = 0.0 if (li < 2 and precision == "bf16") else act(3)
        if precision == "bf16x3":
            y3 = 2.0 * M * c[3]                                       # bf16x3 stores the last layer's pre-activations as bf16
        # SA1 (no input features): dW1 is accumulated by the layer-2 backward epilogue (MaskStatsW1) - that kernel also
        # reads the indices + xyz (src) and does not write dz1; the layer-1 backward kernel does not run
        l2_extra = (src - act(1)) if li == 0 else 0.0
        w.update({
            t + "fwd_l1": (g(0, 1), src + act(1)), t + "fwd_l2": (g(1, 2), act(1) + act(2)),
            t + "fwd_l3": (g(2, 3), act(2) + y3 + pool),
            # v4: one fused wgrad+dgrad kernel per layer
            t + "bwd_l3": (2 * g(2, 3), y3 + act(2) + 5.0 * G * c[3] + act(2)),
            t + "bwd_l2": (2 * g(1, 2) + (g(0, 1) if li == 0 else 0.0), 2 * act(2) + act(1) + act(1) + l2_extra),
            t + "bwd_l1": (g(0, 1) + dg1, 2 * act(1) + src + scat),
            # v5 / first-generation path: separate kernels
            t + "bwd_wgrad3": (g(2, 3), y3 + act(2) + 5.0 * G * c[3]),
            t + "bwd_dgrad3": (g(2, 3), y3 + act(2) + 5.0 * G * c[3] + act(2)),
            t + "bwd_wgrad2": (g(1, 2), 2 * act(2) + act(1)),
            # bf16x3 SA1: the layer-2 dgrad epilogue also accumulates dW1 (MaskStatsW6): it reads the indices + xyz, recomputes
            # y1 = W1 x0 instead of reading it and does not write dz1; the layer-1 wgrad kernel does not run
            t + "bwd_dgrad2": ((g(1, 2) + 2 * g(0, 1), 2 * act(2) + src) if (li == 0 and precision == "bf16x3")
                               else (g(1, 2), 2 * act(2) + act(1) + act(1))),
            t + "bwd_wgrad1": (g(0, 1), 2 * act(1) + src), t + "bwd_dgrad1": (dg1, 2 * act(1) + scat)})
    grp = float(sum(B * (12 * s[5] + 12 * s[6] + 4 * s[6] * s[7]) for s in sh[:2]))
    smp = float(sum(B * (12 * s[5] + 4 * s[6] + 12 * s[6]) for s in sh[:2]))
    w["knn_kernel"] = (0.0, grp)
    w["ball_query_kernel"] = (0.0, grp)
    w["fps_kernel"] = (0.0, smp)
    w["random_subset_kernel"] = (0.0, float(sum(B * 4 * s[6] for s in sh[:2])))
    w["gather_points_kernel"] = (0.0, float(sum(B * (4 * s[6] + 24 * s[6]) for s in sh[:2])))
    return w


def sampling_grouping_leg(pcoe, torch, dev, B, N, peaks, flush, iters=20):
    """Stand-alone timing of the sampling / grouping / gather kernels at this config's shapes (north star: 'achieved HBM
    GB/s for FPS / ball-query / gather').  Each kernel: `iters` launches, CUDA-event pair per launch, L2 flushed between
    launches; bytes = SURVEY 8(d)'s algorithmic bytes per cloud x B (fp32 xyz, int32 indices, no re-reads).  FPS is
    latency-bound by construction (S dependent arg-max steps per cloud), kNN / ball query are instruction-bound
    (S*N distance evaluations + selection per cloud against ~50 KB of compulsory traffic): the HBM fraction is reported as
    the north star asks, next to the instruction-side figure (distance evaluations per second)."""
    xyz = pcoe.synthetic.clouds(7, B, N).to(dev)
    S1, K = 128, 32
    start = torch.zeros(B, dtype=torch.int32, device=dev)
    fps_idx, new_xyz = pcoe.ops.farthest_point_sample(xyz, S1, start, return_xyz=True, int32=True)
    l1_xyz = new_xyz
    feats = torch.randn(B, S1, 128, device=dev)
    idx2 = fps_idx[:, :32].clamp(max=S1 - 1).contiguous()
    nbr = pcoe.ops.knn_int32(new_xyz, xyz, K)
    out = {}

    def timed(name, fn, nbytes, evals=None, note=None):
        fn(); torch.cuda.synchronize()
        ms = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        ms.sort()
        t = ms[len(ms) // 2] * 1e-3
        ent = {"us": t * 1e6, "algorithmic_bytes": nbytes, "gbs": nbytes / t / 1e9, "hbm_frac": nbytes / t / 1e9 / peaks["hbm"]}
        if evals:
            ent["distance_evals_per_s"] = evals / t
        if note:
            ent["note"] = note
        out[name] = ent

    timed(f"fps N={N} S={S1}", lambda: pcoe.ops.farthest_point_sample(xyz, S1, start, return_xyz=True, int32=True),
          B * (12 * N + 16 * S1), evals=B * S1 * N, note="latency-bound: S dependent block-wide arg-max steps per cloud")
    timed("fps N=128 S=32", lambda: pcoe.ops.farthest_point_sample(l1_xyz, 32, start, return_xyz=True, int32=True),
          B * (12 * 128 + 16 * 32), evals=B * 32 * 128)
    timed(f"knn N={N} S={S1} K={K}", lambda: pcoe.ops.knn_int32(new_xyz, xyz, K), B * (12 * N + 12 * S1 + 4 * S1 * K),
          evals=B * S1 * N, note="instruction-bound: S*N distance evaluations + exact K-selection per cloud")
    timed(f"ball_query N={N} S={S1} K={K} r=0.2", lambda: pcoe.ops.ball_query_int32(0.2, K, xyz, new_xyz),
          B * (12 * N + 12 * S1 + 4 * S1 * K), evals=B * S1 * N)
    timed(f"ball_query_multi N={N} S={S1} r=(0.1,0.2,0.4) K=(16,32,128)",
          lambda: pcoe.ops.ball_query_multi_int32([0.1, 0.2, 0.4], [16, 32, 128], xyz, new_xyz),
          B * (12 * N + 12 * S1 + 4 * S1 * (16 + 32 + 128)), evals=B * S1 * N)
    timed(f"gather_points xyz S={S1}", lambda: pcoe.ops.gather_points(xyz, fps_idx), B * (4 * S1 + 24 * S1))
    timed("gather_points feats (B,128,128)->(B,32,128)", lambda: pcoe.ops.gather_points(feats, idx2), B * (4 * 32 + 2 * 512 * 32))
    grouped_idx = nbr.reshape(B, S1 * K)
    timed(f"gather_points grouped xyz (B,{S1}*{K},3)", lambda: pcoe.ops.gather_points(xyz, grouped_idx),
          B * (4 * S1 * K + 12 * N + 12 * S1 * K), note="the stand-alone grouped gather the fused SA kernels avoid")
    timed(f"square_distance ({S1} x {N})", lambda: pcoe.square_distance(new_xyz, xyz), B * (12 * N + 12 * S1 + 4 * S1 * N),
          note="write-bound: the (B,S,N) matrix once; the hot path never materialises it")
    sizes = [N * 4] * B
    cache = pcoe.data.PointCloudCache.from_arrays([pcoe.synthetic.clouds(9, 1, n)[0].numpy() for n in sizes],
                                                 torch.zeros(B, 2).numpy(), kind="vm", device=dev)
    ids = torch.arange(B, device=dev)
    timed(f"resample_clouds n={4 * N}->{N}", lambda: cache.batch(ids, N, seed=1, draw=0), B * (12 * N + 12 * N),
          note="device input pipeline: exact uniform subset per cloud (5 hash passes over n keys + ordered compaction)")
    return out


def load_traffic() -> dict:
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels captured with
    `ncu --set full`, summarised by tools/ncu_summary.py into profiles/traffic.json (kernel name -> bytes)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def make_targets(kind: str, B: int, pcoe, batch: int):
    syn = pcoe.synthetic
    if kind == "mvm":
        gt, K = syn.mvm_targets(B, batch)
        return (gt, K.to(dtype=__import__("torch").int32))
    if kind == "vonmises":
        return syn.vm_targets(B, batch)
    if kind == "8dir":
        return (syn.dir8_targets(B, pcoe.DIRS_8, batch),)
    return ()


def loss_of(kind: str, res, tg, pcoe):
    if kind == "mvm":
        return pcoe.match_loss(res[0], res[1], res[2], tg[0], None, tg[1]).mean()
    if kind == "vonmises":
        return pcoe.kl_von_mises(res[0], res[1], tg[0], tg[1]).mean()
    if kind == "8dir":
        return pcoe.kl_loss_per_sample_from_logits(res, tg[0]).mean()
    return sum((r ** 2).sum() for r in res)


def _leave_multirank(torch, dist, graphs=()):
    """Orderly multi-rank exit: final barrier, drop the captured graphs (they hold references to the NCCL communicator),
    then destroy_process_group().  That call has been observed to hang on this stack (torch 2.11 / NCCL 2.28) while a
    captured graph was alive, so it runs under a watchdog: if it has not returned after 20 s the rank leaves with
    os._exit(0) - after the JSON line has been flushed, so a stuck teardown can never cost the measurement."""
    import gc
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    for g in graphs:
        try:
            if g is not None and getattr(g, "graph", None) is not None:
                g.graph.reset()
        except Exception:
            pass
    gc.collect()
    torch.cuda.synchronize()
    t = threading.Timer(20.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    try:
        dist.destroy_process_group()
    finally:
        t.cancel()


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference algorithm (oracle port, torch CPU) on the box's host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import pcoe
    from oracle.step import OracleTrainer
    kind, cls, B, N = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1000)
    model = getattr(pcoe, cls)()                       # construction only (CPU): reference-identical init
    tr = OracleTrainer(kind, model.state_dict())
    batches = [(pcoe.synthetic.clouds(1, B, N, i), tuple(t.long() if not t.is_floating_point() else t
                                                        for t in make_targets(kind, B, pcoe, i))) for i in range(4)]
    torch.manual_seed(42)
    for i in range(args.warmup):
        tr.step(*batches[i % 4])
    t0 = time.perf_counter()
    for i in range(args.steps):
        tr.step(*batches[i % 4])
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    line = {
        "impl": "reference", "metric": "train clouds/sec (1024 pts, fwd+bwd)", "value": v, "unit": "clouds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.config], "clouds_per_step": B, "points": N},
        "cpu_baseline": {"value": v, "unit": "clouds/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full steps of {B} clouds (oracle port of the reference step, torch CPU)"},
        "e2e": {"value": v, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import pcoe

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this framework has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kind, cls, B, N = CONFIGS[args.config]
    pcoe.set_default_precision(args.precision)
    if args.trunk_tf32:                                  # the 1024-512-256 trunk + heads stay torch.nn (cuBLAS)
        torch.backends.cuda.matmul.allow_tf32 = True
    peaks = load_peaks()

    sampler = "randperm_host" if args.sampler == "host" else "randperm_device"
    torch.manual_seed(1000)
    model = getattr(pcoe, cls)(sampler=sampler).to(dev).train()
    engine = pcoe.dp.DataParallel(model, overlap=not args.no_overlap, exchange=args.exchange)
    if args.skip_allreduce:                              # diagnostic only: how much of the N>1 step is the exchange
        engine.allreduce_grads = lambda: None
        engine._late_off = None
        for m in model.modules():
            if hasattr(m, "_after_backward"):
                m._after_backward = None
    clip = 1.0 if kind == "mvm" else None                # train_multi_peaks_vonMises_KL.py:235
    if args.torch_optimizer:
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=not args.no_graph)
    else:                                                # clip + Adam + zero_grad as two libpcoe launches
        opt = pcoe.optim.FusedAdam(engine, lr=1e-3, max_grad_norm=clip, zero_grad_in_step=True)
    params = [p for p in model.parameters()]
    NB = 8                                               # distinct synthetic batches, rotated
    host = [(pcoe.synthetic.clouds(1, B, N, rank * NB + i).pin_memory(),
             tuple(t.pin_memory() for t in make_targets(kind, B, pcoe, rank * NB + i))) for i in range(NB)]
    resident = [(x.to(dev), tuple(t.to(dev) for t in tg)) for x, tg in host]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(xyz, tg):
        if args.torch_optimizer:
            engine.zero_grad()
        res = model(xyz)
        loss = loss_of(kind, res, tg, pcoe)
        loss.backward()
        engine.allreduce_grads()
        if clip is not None and args.torch_optimizer:
            torch.nn.utils.clip_grad_norm_(params, clip, foreach=True)
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region ---------------------------------------------------------
    # the whole step is captured in a CUDA graph (pcoe.GraphedTrainStep) and replayed per batch
    for i in range(args.warmup):
        step(*resident[i % NB])
    graphed, cap_launches = None, None
    if not args.no_graph:
        l0 = pcoe._lib.launch_count()
        graphed = pcoe.GraphedTrainStep(model, lambda res, *tg: loss_of(kind, res, tg, pcoe), opt, resident[0][0],
                                        resident[0][1], clip_norm=clip, engine=engine, warmup=1)
        cap_launches = (pcoe._lib.launch_count() - l0) // 2          # 1 warm-up + 1 captured step
        for i in range(2):
            graphed(*([resident[i][0]] + list(resident[i][1])))
    run = (lambda x, tg: graphed(x, *tg)) if graphed is not None else step
    barrier()
    clock_sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = pcoe._lib.launch_count()
    t_wall0 = time.time()
    for i in range(args.steps):
        flush.zero_()                                     # L2 flush, outside the event pair
        ev[i][0].record()
        run(*resident[i % NB])
        ev[i][1].record()
    barrier()
    t_wall1 = time.time()
    launches = pcoe._lib.launch_count() - launches0
    if graphed is not None:
        launches = cap_launches * args.steps              # replayed from the graph, not re-enqueued
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    clocks = clock_sampler.stop(t_wall0, t_wall1) if clock_sampler else None
    value = world * B * args.steps / (dev_ms / 1e3)

    if args.timed_only:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": dev_ms / args.steps, "gpu_launches": int(launches),
                              "precision": args.precision, "timed_only": True}), flush=True)
        if world > 1:                                    # same multi-rank exit as the full run (see the end)
            _leave_multirank(torch, dist, (graphed,))
        return

    # ---- end-to-end through the public API with host buffers ------------------------------------
    # (a) the graphed step fed from pinned host memory: H2D of xyz + targets, replay, loss D2H
    h2d = host[0][0].numel() * 4 + sum(t.numel() * t.element_size() for t in host[0][1])
    if sampler == "randperm_host":                        # + the replayed subsets: (B,128) and (B,32) int32
        h2d += 4 * B * (128 + 32)
    e2e_steps = max(3, min(args.steps, 100))

    def timed_e2e(fn):
        for i in range(2):
            float(fn(i))
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            float(fn(i))                                  # D2H read of the step's result
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * B * e2e_steps / float(tt.item())

    e2e_value = None
    if graphed is not None:
        if args.no_prefetch:
            e2e_value = timed_e2e(lambda i: graphed(host[i % NB][0], *host[i % NB][1]))
        else:
            # double-buffered feed: the H2D copy of batch i+1 is enqueued on a copy stream right after step i is
            # launched and BEFORE its loss is read back, so it overlaps the step; every timed step still contains one
            # pinned-host -> device copy of a full batch and the D2H read of its loss
            def fed(i):
                loss = graphed.step_prefetched()
                graphed.prefetch(host[(i + 1) % NB][0], *host[(i + 1) % NB][1])
                return loss
            graphed.prefetch(host[0][0], *host[0][1])
            e2e_value = timed_e2e(fed)
    # (b) the plain eager drop-in call (no CUDA graph), same sampler: what a caller gets without GraphedTrainStep
    if graphed is not None:
        graphed.release()
    e2e_eager = timed_e2e(lambda i: step(host[i % NB][0].to(dev, non_blocking=True),
                                         tuple(t.to(dev, non_blocking=True) for t in host[i % NB][1])).detach())
    if e2e_value is None:
        e2e_value = e2e_eager

    # ---- per-kernel CUDA-event profile of the same step (rank 0) -> roofline ---------------------
    roofline, kernels = None, {}
    if rank == 0:
        barrier_local = torch.cuda.synchronize
        barrier_local()
    psteps = 3
    pcoe._lib.profile(True)
    for i in range(psteps):
        flush.zero_()
        step(*resident[i % NB])
    torch.cuda.synchronize()
    pcoe._lib.profile(False)
    rep = pcoe._lib.profile_report()
    if rank == 0:
        work = kernel_work(B, N, args.precision)
        tot_ms = sum(ms for _, ms in rep.values())
        for name, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
            ent = {"launches_per_step": n / psteps, "ms_per_step": ms / psteps, "share_of_libpcoe_time": ms / tot_ms}
            if name in work:
                flops, nbytes = work[name]
                t_s = ms / psteps * 1e-3
                t_tensor, t_hbm = flops / (peaks["tensor"] * 1e12), nbytes / (peaks["hbm"] * 1e9)
                bound = "tensor" if t_tensor > t_hbm else "hbm"
                ach = flops / t_s / 1e12 if bound == "tensor" else nbytes / t_s / 1e9
                ent.update(bound=bound, achieved=ach, unit="TFLOP/s" if bound == "tensor" else "GB/s",
                           frac=ach / peaks[bound], tflops=flops / t_s / 1e12, gbs=nbytes / t_s / 1e9,
                           tensor_frac=flops / t_s / 1e12 / peaks["tensor"])
            kernels[name] = ent
        # dominant kernel = the longest single launch among the kernels with a roofline
        cands = [n for n in kernels if "achieved" in kernels[n]]
        top = max(cands, key=lambda n: kernels[n]["ms_per_step"] / kernels[n]["launches_per_step"]) if cands else None
        if top:
            k = kernels[top]
            traffic = load_traffic().get(args.precision, {}).get(top) if args.config == "c2" else None
            roofline = {"kernel": top, "bound": k["bound"], "achieved": k["achieved"], "peak": peaks[k["bound"]],
                        "unit": k["unit"], "frac": k["frac"], "traffic": traffic, "peak_source": peaks["source"],
                        "avg_launch_ms": k["ms_per_step"] / k["launches_per_step"],
                        "share_of_libpcoe_time": k["share_of_libpcoe_time"],
                        "algorithmic_bytes": work[top][1] / k["launches_per_step"],
                        "algorithmic_flops": work[top][0] / k["launches_per_step"],
                        "tensor_frac": k.get("tensor_frac"),
                        "note": "bound = the slower of FLOPs/tensor peak and compulsory bytes/HBM peak for this kernel"}

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.step import OracleTrainer
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        tr = OracleTrainer(kind, {k: v.detach().cpu() for k, v in model.state_dict().items()})
        cb = [(x.clone(), tuple(t.long() if not t.is_floating_point() else t.clone() for t in tg)) for x, tg in host[:2]]
        tr.step(*cb[0])
        n_cpu = 4
        t0 = time.perf_counter()
        for i in range(n_cpu):
            tr.step(*cb[i % 2])
        dt = time.perf_counter() - t0
        cpu = {"value": B * n_cpu / dt, "unit": "clouds/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} full steps of {B} clouds after 1 warm-up (oracle port of the reference step, torch CPU fp32)"}

    # ---- the reference algorithm in torch eager ON THE SAME GPU (rank 0, N=1) -----------------------------------------
    # The reference's scripts run on CUDA when it is available (train_multi_peaks_vonMises_KL.py:27).  The reference tree
    # does not travel to the GPU box, so this leg runs the oracle port of its training step (oracle/step.py: the same
    # torch ops - gather, cdist-style kNN + topk, conv-as-matmul, batch_norm, max, host randperm, host Hungarian) on
    # this device: the like-for-like "stock PyTorch on a B200" number next to `value`.
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.step import OracleTrainer
        tr = OracleTrainer(kind, {k: v.detach().cpu() for k, v in model.state_dict().items()}, device=dev)
        gb = [(x, tuple(t.long() if not t.is_floating_point() else t for t in tg)) for x, tg in resident[:4]]
        for i in range(2):
            tr.step(*gb[i])
        torch.cuda.synchronize()
        n_g = 10
        t0 = time.perf_counter()
        for i in range(n_g):
            tr.step(*gb[i % 4])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gpu_eager = {"value": B * n_g / dt, "unit": "clouds/s", "ms_per_step": 1e3 * dt / n_g, "kind": "port",
                     "sample": f"{n_g} full steps of {B} clouds after 2 warm-ups: oracle port of the reference step in torch eager "
                               "on this GPU (fp32, host randperm + host Hungarian as the reference), inputs resident"}

    # ---- the same step in the other precision modes (rank 0, N=1) ---------------------------------
    # `value` is the mode named in config.precision.  The other modes are timed here on the same workload so that all
    # three are on record: bf16x3 (parity-gated tensor-core mode), bf16 (throughput mode, stated tolerance), fp32
    # (CUDA-core kernels).  Tensor-core modes: captured step, 30 replays; fp32: eager, 10 steps.
    modes = None
    if rank == 0 and world == 1 and not args.no_modes_leg:
        modes = {args.precision: {"value": value, "ms_per_step": dev_ms / args.steps, "mode": "this run's `value`"}}
        for prec in ("bf16x3", "bf16", "fp32"):
            if prec == args.precision:
                continue
            m2 = getattr(pcoe, cls)(sampler="randperm_device", precision=prec).to(dev).train()
            m2.load_state_dict(model.state_dict())
            eng2 = pcoe.dp.DataParallel(m2)
            opt2 = pcoe.optim.FusedAdam(eng2, lr=1e-3, max_grad_norm=clip, zero_grad_in_step=True)

            def step2(x, tg):
                l = loss_of(kind, m2(x), tg, pcoe)
                l.backward()
                opt2.step()
                return l

            for i in range(3):
                step2(*resident[i % NB])
            n2, run2, how = 10, step2, "eager (no CUDA graph)"
            if prec != "fp32":
                g2 = pcoe.GraphedTrainStep(m2, lambda res, *tg: loss_of(kind, res, tg, pcoe), opt2, resident[0][0],
                                           resident[0][1], clip_norm=clip, engine=eng2, warmup=1)
                n2, run2, how = 30, (lambda x, tg: g2(x, *tg)), "captured step, device-side sampler"
            for i in range(2):
                run2(*resident[i % NB])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(n2):
                run2(*resident[i % NB])
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / n2
            modes[prec] = {"value": B / (ms2 / 1e3), "ms_per_step": ms2, "steps": n2, "mode": how + ", no L2 flush"}
            del m2, eng2, opt2

    sg = None
    if rank == 0 and world == 1 and not args.no_sampling_leg:
        sg = {f"{B} clouds x {N} points": sampling_grouping_leg(pcoe, torch, dev, B, N, peaks, flush)}
        if (B, N) != (32, 8192):                          # the FPS / ball-query stress shape of BASELINE configs[3]
            sg["32 clouds x 8192 points"] = sampling_grouping_leg(pcoe, torch, dev, 32, 8192, peaks, flush)

    if rank == 0:
        line = {
            "metric": "train clouds/sec (1024 pts, fwd+bwd)", "value": value, "unit": "clouds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16": "bf16", "bf16x3": "bf16x3 (operands split into bf16 planes, fp32 accumulate and activations)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[args.config], "clouds_per_gpu": B, "points": N,
                       "global_clouds_per_step": B * world, "parallelism": f"dp{world}",
                       "step": "zero_grad+fwd+loss+bwd+allreduce+clip+Adam",
                       "optimizer": "torch.optim.Adam(fused)+clip_grad_norm_" if args.torch_optimizer
                       else "pcoe.optim.FusedAdam (clip+Adam+zero_grad, 2 launches)",
                       "sampler": ("randperm_host: the reference's torch.randperm stream replayed on the host generator "
                                   "(bit-identical), uploaded as graph inputs every step - used by `value` AND `e2e`")
                       if sampler == "randperm_host" else "randperm_device (Philox subset kernel inside the graph; NOT the reference's stream)",
                       "precision": args.precision,
                       "l2": "512 MiB buffer written between timed steps (flush outside the event pair)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clouds/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps,
                    "mode": ("CUDA-graph step fed from pinned host buffers"
                             + ("" if args.no_prefetch else ", next batch's H2D copy overlapped with the running step")) if graphed else "eager",
                    "eager_host_sampler_value": e2e_eager},
            "cuda_graph": graphed is not None,
            "gpu_launches": int(launches),
            "gpu_launches_per_step": launches / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "gpu_eager_port": gpu_eager, "modes": modes, "kernels": kernels,
            "sampling_grouping": sg,
            "wall_ms_per_step_incl_flush": 1e3 * (t_wall1 - t_wall0) / args.steps,
            "grad_allreduce_bytes": engine.grads.nbytes() if world > 1 else 0,
            "grad_exchange": ("peer (libpcoe two-shot all-reduce over NVLink symmetric memory)" if engine.peer is not None else "nccl") if world > 1 else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        _leave_multirank(torch, dist, (graphed,))


# ------------------------------------------------------------------------------------------------
# c5: vanilla PointNet inference sweep (BASELINE configs[4]; not the headline metric - run with --config c5)
# ------------------------------------------------------------------------------------------------
def _pointnet_model(pcoe, torch):
    torch.manual_seed(1000)
    model = pcoe.PointNet(feature_transform=True)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():                                  # a trained checkpoint's BatchNorm state, not the identity
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.2, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.uniform_(-0.3, 0.3, generator=g)
    return model.eval()


def run_pointnet_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    import pcoe
    from oracle import pointnet_torch, sa_torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = sa_torch.clone_state(_pointnet_model(pcoe, torch).state_dict())
    Bc = 64
    x = pcoe.synthetic.clouds(5, Bc, 1024)
    with torch.no_grad():
        for _ in range(max(1, args.warmup)):
            pointnet_torch.pointnet_forward(sd, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pointnet_torch.pointnet_forward(sd, x)
        dt = time.perf_counter() - t0
    v = Bc * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "inference clouds/sec (PointNet, 1024 pts, eval forward)", "value": v, "unit": "clouds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES["c5"], "clouds_per_step": Bc, "points": 1024},
        "cpu_baseline": {"value": v, "unit": "clouds/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} eval forwards of {Bc} clouds (oracle port of models/pointnet.py, torch CPU)"},
        "e2e": {"value": v, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_pointnet(args):
    import torch
    import pcoe
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this framework has no CPU path (use --impl reference)")
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    model = _pointnet_model(pcoe, torch).to(dev)
    N = 1024
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    sweep = {}
    steps = max(5, min(args.steps, 50))
    clock_sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
        x = pcoe.synthetic.clouds(5, B, N).to(dev)
        with torch.no_grad():
            for _ in range(max(3, args.warmup)):
                y = model(x)
            torch.cuda.synchronize()
            # small batches are launch-bound: replay the forward from a CUDA graph (inputs copied into the static buffer)
            graph, static_x, static_y = None, x.clone(), None
            if B <= 64 and not args.no_graph:
                try:
                    sstream = torch.cuda.Stream()
                    sstream.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(sstream):
                        model(static_x)
                    torch.cuda.current_stream().wait_stream(sstream)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        static_y = model(static_x)
                except Exception as e:                      # capture is an optimisation, not the contract
                    graph, static_y = None, None
                    sys.stderr.write(f"c5: CUDA-graph capture failed at B={B}: {e}\n")
            run = (lambda: graph.replay()) if graph is not None else (lambda: model(static_x))
            for _ in range(2):
                run()
            l0 = pcoe._lib.launch_count()
            model(static_x)
            nl = pcoe._lib.launch_count() - l0
            ms = []
            for _ in range(steps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(); b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            ms.sort()
            med = ms[len(ms) // 2]
            # end to end: pinned host xyz -> device, forward, (B,3) result back
            hx = pcoe.synthetic.clouds(6, B, N).pin_memory()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                static_x.copy_(hx, non_blocking=True)
                if graph is not None:
                    graph.replay()
                    out = static_y.cpu()
                else:
                    out = model(static_x).cpu()
            torch.cuda.synchronize()
            e2e_ms = 1e3 * (time.perf_counter() - t0) / steps
        sweep[str(B)] = {"latency_ms": med, "clouds_per_s": B / med * 1e3, "e2e_ms": e2e_ms, "e2e_clouds_per_s": B / e2e_ms * 1e3,
                         "cuda_graph": graph is not None, "libpcoe_launches": int(nl)}
        del graph
    t_wall1 = time.time()
    clocks = clock_sampler.stop(t_wall0, t_wall1) if clock_sampler else None
    # per-kernel profile at B = 1024 -> roofline of the dominant kernel (the 128 -> 1024 layer + max pool)
    B = 1024
    x = pcoe.synthetic.clouds(5, B, N).to(dev)
    pcoe._lib.profile(True)
    with torch.no_grad():
        for _ in range(3):
            flush.zero_()
            model(x)
    torch.cuda.synchronize()
    pcoe._lib.profile(False)
    rep = pcoe._lib.profile_report()
    kernels = {k: {"launches_per_step": n / 3, "ms_per_step": ms_ / 3} for k, (n, ms_) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
    M = B * N
    roofline = None
    if "pointmlp_l3_pool" in kernels:
        k = kernels["pointmlp_l3_pool"]
        t_s = k["ms_per_step"] / k["launches_per_step"] * 1e-3
        flops = 2.0 * M * 128 * 1024
        ach = flops / t_s / 1e12
        roofline = {"kernel": "pointmlp_l3_pool (pm_layer_kernel<PoolPm, true>: TMA-fed 128 -> 1024 layer + max pool)", "bound": "tensor", "achieved": ach,
                    "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": ach / peaks["tensor"], "traffic": None,
                    "peak_source": peaks["source"], "avg_launch_ms": t_s * 1e3, "algorithmic_flops": flops,
                    "algorithmic_bytes": 4.0 * M * 128 + 10.0 * (M / 32) * 1024,
                    "tensor_pipe_occupancy": 3.0 * ach / peaks["tensor"],
                    "note": "algorithmic FLOPs = 2*M*128*1024 (one product per MAC); the kernel issues 3 tcgen05.mma per "
                            "product (two bf16 planes per operand: 16 significant bits), so the tensor pipe is busy 3x this "
                            "fraction of the (sustained) peak - `tensor_pipe_occupancy`"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pointnet_torch, sa_torch
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = sa_torch.clone_state(model.state_dict())
        xc = pcoe.synthetic.clouds(5, 64, N)
        with torch.no_grad():
            pointnet_torch.pointnet_forward(sd, xc)
            t0 = time.perf_counter()
            for _ in range(3):
                pointnet_torch.pointnet_forward(sd, xc)
            dt = time.perf_counter() - t0
        cpu = {"value": 64 * 3 / dt, "unit": "clouds/s", "cores": cores, "kind": "port",
               "sample": "3 eval forwards of 64 clouds after 1 warm-up (oracle port of models/pointnet.py, torch CPU fp32)"}
    if rank == 0:
        top = sweep["1024"]
        print(json.dumps({
            "metric": "inference clouds/sec (PointNet, 1024 pts, eval forward)", "value": top["clouds_per_s"], "unit": "clouds/s",
            "n_gpus": 1, "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": top["latency_ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 (operands split into bf16 planes, fp32 accumulate and activations)", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES["c5"], "clouds_per_gpu": 1024, "points": N,
                       "l2": "512 MiB buffer written between timed forwards (flush outside the event pair)",
                       "value_is": "B = 1024 row of `sweep`"},
            "clocks": clocks,
            "e2e": {"value": top["e2e_clouds_per_s"], "unit": "clouds/s", "h2d_bytes_per_step": 1024 * N * 12, "d2h_bytes_per_step": 1024 * 12},
            "gpu_launches": int(top["libpcoe_launches"]) * steps, "gpu_launches_per_step": int(top["libpcoe_launches"]),
            "sweep": sweep, "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c2")
    ap.add_argument("--precision", choices=["fp32", "bf16x3", "bf16"], default=os.environ.get("PCOE_PRECISION", "bf16x3"))
    ap.add_argument("--sampler", choices=["host", "device"], default="host",
                    help="host = the reference's randperm stream replayed on the CPU generator and fed to the graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes-leg", action="store_true", help="skip the throughput of the other precision modes")
    ap.add_argument("--no-sampling-leg", action="store_true", help="skip the stand-alone FPS / grouping / gather timings")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the CUDA-graph step")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e leg: copy each batch H2D in line with its step")
    ap.add_argument("--torch-optimizer", action="store_true",
                    help="torch.optim.Adam(fused) + clip_grad_norm_ instead of pcoe.optim.FusedAdam")
    ap.add_argument("--trunk-tf32", action="store_true", help="TF32 tensor-core cuBLAS kernels for the torch.nn trunk")
    ap.add_argument("--exchange", choices=["nccl", "peer", "auto"], default=os.environ.get("PCOE_EXCHANGE", "auto"),
                    help="gradient exchange at N > 1: torch.distributed (NCCL) all-reduce or libpcoe's NVLink peer-memory kernel")
    ap.add_argument("--no-overlap", action="store_true", help="one all-reduce after backward instead of two overlapped buckets")
    ap.add_argument("--skip-allreduce", action="store_true", help="diagnostic: N>1 without the gradient exchange (INVALID as a result)")
    ap.add_argument("--timed-only", action="store_true",
                    help="warm-up + timed region only (for ncu captures): no e2e / per-kernel / CPU legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.config == "c5":
        (run_pointnet_reference if args.impl == "reference" else run_pointnet)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
