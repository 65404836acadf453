"""pcoe.optim.FusedAdam (pcoe_adam_step) vs torch.optim.Adam + clip_grad_norm_, and direct gradient
accumulation of the set-abstraction backward into the flat gradient buffer vs the autograd path."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(37, 53)          # odd sizes: flat offsets that are not multiples of 4
        self.b = torch.nn.Linear(53, 7)
        self.c = torch.nn.Parameter(torch.randn(3))

    def forward(self, x):
        return self.b(torch.tanh(self.a(x))) * self.c.sum()


@pytest.mark.parametrize("max_norm,wd", [(None, 0.0), (1.0, 0.0), (0.05, 0.0), (1.0, 0.01)])
def test_fused_adam_matches_torch(pcoe, cuda, max_norm, wd):
    torch.manual_seed(3)
    ref = _Net().to(cuda)
    ours = copy.deepcopy(ref)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=wd)
    opt = pcoe.optim.FusedAdam(ours, lr=1e-3, weight_decay=wd, max_grad_norm=max_norm)
    for it in range(6):
        x = torch.randn(16, 37, device=cuda) * (1 + 3 * it)
        opt_ref.zero_grad()
        opt.zero_grad()
        ref(x).square().mean().backward()
        ours(x).square().mean().backward()
        want_norm = None
        if max_norm is not None:
            want_norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        opt_ref.step()
        opt.step()
        if want_norm is not None:
            assert torch.allclose(opt.grad_norm[0], want_norm, rtol=1e-5)
            for p, q in zip(ours.parameters(), ref.parameters()):       # grads hold the clipped gradient
                assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-7)
        for (n, p), q in zip(ours.named_parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=1e-4, atol=2e-6), (it, n, float((p - q).abs().max()))   # 2e-6 = 0.2 % of one lr step
    assert int(opt.step_dev) == 6
    # state_dict in torch.optim.Adam's layout; loads into a torch optimizer and back
    sd = opt.state_dict()
    sd_ref = opt_ref.state_dict()
    assert sd["state"].keys() == sd_ref["state"].keys()
    for i in sd["state"]:
        assert float(sd["state"][i]["step"]) == float(sd_ref["state"][i]["step"])
        assert torch.allclose(sd["state"][i]["exp_avg"], sd_ref["state"][i]["exp_avg"], rtol=1e-4, atol=1e-8)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], sd_ref["state"][i]["exp_avg_sq"], rtol=1e-4, atol=1e-12)
    opt2 = pcoe.optim.FusedAdam(copy.deepcopy(ref), lr=5e-4)
    opt2.load_state_dict(sd_ref)
    assert int(opt2.step_dev) == 6 and opt2.defaults["lr"] == 1e-3
    assert torch.allclose(opt2.exp_avg, opt.exp_avg, rtol=1e-4, atol=1e-8)


def test_fused_adam_zero_grad_in_step_and_graph(pcoe, cuda):
    """The step is capturable: the counter and the norm live on the device; gradients are cleared in the step."""
    torch.manual_seed(4)
    ref = _Net().to(cuda)
    ours = copy.deepcopy(ref)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3)
    opt = pcoe.optim.FusedAdam(ours, lr=1e-3, max_grad_norm=1.0, zero_grad_in_step=True)
    x = torch.randn(16, 37, device=cuda)
    xs = x.clone()

    def step():
        ours(xs).square().mean().backward()
        opt.step()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert int(opt.step_dev) == 4
    assert float(opt.grads.flat.abs().max()) == 0.0
    for _ in range(4):
        opt_ref.zero_grad()
        ref(x).square().mean().backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        opt_ref.step()
    for p, q in zip(ours.parameters(), ref.parameters()):
        assert torch.allclose(p, q, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_direct_grad_accumulation_equals_autograd_path(pcoe, cuda, precision):
    """FlatGradBuffer makes pcoe_sa_backward add into p.grad (pcoe_sa_grads.accumulate): same gradients
    as returning them to autograd, and a second backward accumulates (p.grad += g)."""
    torch.manual_seed(11)
    B, N, D = 4, 256, 32
    xyz = pcoe.synthetic.clouds(1, B, N, 0).to(cuda)
    feats = torch.randn(B, N, D, device=cuda)
    idx = torch.stack([torch.randperm(N)[:32] for _ in range(B)]).to(cuda)

    def run(direct):
        torch.manual_seed(12)
        layer = pcoe.PointNetSetAbstraction(32, 32, D, [64, 64, 128], precision=precision).to(cuda).train()
        if direct:
            buf = pcoe.dp.FlatGradBuffer(layer)
            assert layer.direct_grad_accumulation
        f = feats.clone().requires_grad_(True)
        for _ in range(2):                                   # two backward passes: accumulation semantics
            _, out = layer(xyz, f, fps_idx=idx)
            (out * torch.linspace(-1, 1, out.size(-1), device=cuda)).sum().backward()
        return [p.grad.clone() for p in layer.parameters()], f.grad.clone()

    ga, fa = run(False)
    gb, fb = run(True)
    tol = dict(rtol=1e-4, atol=1e-5) if precision == "fp32" else dict(rtol=2e-3, atol=2e-3)   # atomics order only
    assert torch.allclose(fa, fb, **tol)
    for a, b in zip(ga, gb):
        assert torch.allclose(a, b, **tol), float((a - b).abs().max())


def test_fused_adam_grad_scale_is_the_data_parallel_mean(pcoe, cuda):
    """grad_scale = 1/world (set when the optimizer is built on a pcoe.dp.DataParallel engine): the SUM all-reduced
    buffer is averaged inside the step kernel, the clip threshold applies to the averaged gradient."""
    torch.manual_seed(6)
    ref = _Net().to(cuda)
    ours = copy.deepcopy(ref)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3)
    opt = pcoe.optim.FusedAdam(ours, lr=1e-3, max_grad_norm=0.5)
    opt.grad_scale = 0.25                                   # as if 4 ranks had summed their gradients
    for it in range(3):
        x = torch.randn(16, 37, device=cuda)
        opt_ref.zero_grad(); opt.zero_grad()
        ref(x).square().mean().backward()
        (4.0 * ours(x).square().mean()).backward()          # "sum over 4 identical ranks"
        want_norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        opt_ref.step(); opt.step()
        assert torch.allclose(opt.grad_norm[0] * 0.25, want_norm, rtol=1e-4)
        for p, q in zip(ours.parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=1e-4, atol=2e-6)
            assert torch.allclose(p.grad, q.grad, rtol=1e-3, atol=1e-6)


def test_graphed_train_step_and_prefetched_feed(pcoe, cuda):
    """pcoe.GraphedTrainStep (the path bench.py times): the whole mvM training step captured once and replayed -
    losses stay finite and fall on a fixed batch, parameters move, a replay launches no new libpcoe kernels from the
    host, and the double-buffered feed (prefetch / step_prefetched) trains on exactly the staged batch."""
    torch.manual_seed(7)
    B, N = 8, 1024
    model = pcoe.PointNetPPMvM(sampler="randperm_device", precision="bf16").to(cuda).train()
    engine = pcoe.dp.DataParallel(model)
    opt = pcoe.optim.FusedAdam(engine, lr=1e-3, max_grad_norm=1.0, zero_grad_in_step=True)
    host = []
    for i in range(3):
        gt, K = pcoe.synthetic.mvm_targets(B, i)
        host.append((pcoe.synthetic.clouds(1, B, N, i).pin_memory(), gt.pin_memory(), K.to(torch.int32).pin_memory()))
    loss_fn = lambda res, gt, K: pcoe.match_loss(res[0], res[1], res[2], gt, None, K).mean()
    dev0 = tuple(t.to(cuda) for t in host[0])
    for _ in range(2):                                   # eager steps first (allocates counters, flat buffers)
        loss_fn(model(dev0[0]), dev0[1], dev0[2]).backward()
        opt.step()
    g = pcoe.GraphedTrainStep(model, loss_fn, opt, dev0[0], dev0[1:], clip_norm=1.0, engine=engine, warmup=1)
    w0 = model.fc1.weight.detach().clone()
    n0 = pcoe._lib.launch_count()
    losses = [float(g(*host[0])) for _ in range(30)]
    assert pcoe._lib.launch_count() == n0                # replays enqueue nothing from libpcoe's host side
    assert all(l == l and abs(l) < 1e4 for l in losses)
    assert sum(losses[-5:]) < sum(losses[:5])            # same batch 30 times: the loss goes down
    assert not torch.equal(model.fc1.weight.detach(), w0)
    # double-buffered feed: batch 1 staged while a step on batch 0 runs, then consumed
    g.prefetch(*host[1])
    l1 = float(g.step_prefetched())
    assert torch.equal(g.xyz.cpu(), host[1][0]) and torch.equal(g.targets[1].cpu(), host[1][2])
    g.prefetch(*host[2])
    l2 = float(g.step_prefetched())
    assert torch.equal(g.xyz.cpu(), host[2][0]) and l1 == l1 and l2 == l2


def test_graphed_step_trains_on_the_reference_subsets(pcoe, cuda):
    """GraphedTrainStep with the reference's own sampler (sampler="randperm_host", models/pointnet_pp_8dir.py:28): the
    subsets are replayed on torch's CPU generator before every graph replay and uploaded as graph inputs, so a seeded
    graphed run uses exactly the subsets of the seeded reference run: B x randperm(N) then B x randperm(128) per
    forward, step after step - also through the double-buffered feed."""
    B, N = 4, 512
    torch.manual_seed(3)
    model = pcoe.PointNetPPMvM(precision="bf16x3").to(cuda).train()          # default sampler = randperm_host
    engine = pcoe.dp.DataParallel(model)
    opt = pcoe.optim.FusedAdam(engine, lr=1e-3, max_grad_norm=1.0, zero_grad_in_step=True)
    gt, K = pcoe.synthetic.mvm_targets(B, 0)
    host = (pcoe.synthetic.clouds(1, B, N, 0).pin_memory(), gt.pin_memory(), K.to(torch.int32).pin_memory())
    dev0 = tuple(t.to(cuda) for t in host)
    loss_fn = lambda res, gt, K: pcoe.match_loss(res[0], res[1], res[2], gt, None, K).mean()
    loss_fn(model(dev0[0]), dev0[1], dev0[2]).backward()
    opt.step()
    g = pcoe.GraphedTrainStep(model, loss_fn, opt, dev0[0], dev0[1:], clip_norm=1.0, engine=engine, warmup=1)

    def reference_draws(n_forward):
        return [(torch.stack([torch.randperm(N)[:128] for _ in range(B)]), torch.stack([torch.randperm(128)[:32] for _ in range(B)]))
                for _ in range(n_forward)]

    torch.manual_seed(42)
    want = reference_draws(5)
    torch.manual_seed(42)
    for i in range(3):
        loss = g(*host)
        assert torch.equal(model.sa1.last_fps_idx.long().cpu(), want[i][0])
        assert torch.equal(model.sa2.last_fps_idx.long().cpu(), want[i][1])
        assert float(loss) == float(loss)
    g.prefetch(*host)                                    # draws forward 3 now, consumed by the next step
    g.step_prefetched()
    torch.cuda.synchronize()
    assert torch.equal(model.sa1.last_fps_idx.long().cpu(), want[3][0]) and torch.equal(model.sa2.last_fps_idx.long().cpu(), want[3][1])
    g.release()
    model(dev0[0])                                       # eager again: the layer draws for itself, same stream
    assert torch.equal(model.sa1.last_fps_idx.long().cpu(), want[4][0])


def test_graphed_step_with_true_fps_sampler_draws_new_start_points(pcoe, cuda):
    """sampler="fps" inside a captured step: the first centroid of every cloud comes from torch's CUDA generator (no
    host draw, no synchronising copy), so replays start from different points."""
    torch.manual_seed(5)
    B, N = 4, 512
    model = pcoe.PointNetPP8Dir(sampler="fps", precision="bf16x3").to(cuda).train()
    engine = pcoe.dp.DataParallel(model)
    opt = pcoe.optim.FusedAdam(engine, lr=1e-3, zero_grad_in_step=True)
    xyz = pcoe.synthetic.clouds(2, B, N, 0).to(cuda)
    p8 = pcoe.synthetic.dir8_targets(B, pcoe.DIRS_8).to(cuda)
    loss_fn = lambda res, p: pcoe.kl_loss_per_sample_from_logits(res, p).mean()
    loss_fn(model(xyz), p8).backward()
    opt.step()
    g = pcoe.GraphedTrainStep(model, loss_fn, opt, xyz, (p8,), engine=engine, warmup=1)
    firsts = []
    for _ in range(6):
        assert float(g(xyz, p8)) == float(g.loss)
        firsts.append(model.sa1.last_fps_idx[:, 0].clone())
        idx = model.sa1.last_fps_idx.long()
        assert all(len(set(r.tolist())) == 128 for r in idx.cpu())          # FPS never repeats a point
    assert len({tuple(f.tolist()) for f in firsts}) > 1
