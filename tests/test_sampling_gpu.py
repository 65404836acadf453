"""CUDA sampling / grouping kernels (through the C ABI) against golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from oracle import sampling

pytestmark = pytest.mark.gpu


def _unit_clouds(seed, B, N):
    g = torch.Generator("cpu").manual_seed(seed)
    x = torch.randn(B, N, 3, generator=g)
    x = x - x.mean(1, keepdim=True)
    return (x / x.norm(dim=-1).amax(1).view(B, 1, 1)).contiguous()


def test_fps_golden_bit_exact(pcoe, golden, cuda):
    g = golden("sampling")
    for tag in ("a", "b"):
        xyz = torch.from_numpy(g[f"{tag}_xyz"]).to(cuda)
        idx, nx = pcoe.ops.farthest_point_sample(xyz, 64, torch.from_numpy(g[f"{tag}_start"]).to(cuda), return_xyz=True)
        assert np.array_equal(idx.cpu().numpy(), g[f"{tag}_fps"])
        want = torch.gather(xyz, 1, idx.unsqueeze(-1).expand(-1, -1, 3))
        assert torch.equal(nx, want)


@pytest.mark.parametrize("N,S", [(128, 32), (1024, 128), (2048, 128), (8192, 128), (1000, 512), (10000, 64)])
def test_fps_matches_oracle_bit_exact(pcoe, cuda, N, S):
    B = 3
    xyz = _unit_clouds(N + S, B, N)
    start = torch.tensor([0, N // 2, N - 1])
    got = pcoe.ops.farthest_point_sample(xyz.to(cuda), S, start.to(cuda)).cpu().numpy()
    want = sampling.farthest_point_sample(xyz.numpy(), S, start.numpy())
    assert np.array_equal(got, want)
    assert all(len(set(r.tolist())) == S for r in got)       # size-independent property: all distinct


def test_fps_ties_and_duplicates(pcoe, cuda):
    xyz = torch.zeros(2, 40, 3)
    xyz[1, ::2] = 1.0                                          # two distinct locations, many duplicates
    start = torch.tensor([7, 1])
    got = pcoe.ops.farthest_point_sample(xyz.to(cuda), 6, start.to(cuda)).cpu().numpy()
    want = sampling.farthest_point_sample(xyz.numpy(), 6, start.numpy())
    assert np.array_equal(got, want)
    assert got[0].tolist() == [7, 0, 0, 0, 0, 0]


def test_ball_query_golden_slot_exact(pcoe, golden, cuda):
    g = golden("sampling")
    for tag in ("a", "b"):
        xyz = torch.from_numpy(g[f"{tag}_xyz"]).to(cuda)
        fps = torch.from_numpy(g[f"{tag}_fps"]).to(cuda)
        new_xyz = pcoe.index_points(xyz, fps)
        for r, ns in ((0.2, 16), (0.4, 32), (0.05, 8)):
            got = pcoe.ball_query(r, ns, xyz, new_xyz).cpu().numpy()
            assert np.array_equal(got, g[f"{tag}_ball_{r}_{ns}"]), (tag, r, ns)


@pytest.mark.parametrize("N,S,r,ns", [(1024, 128, 0.2, 32), (2048, 64, 0.4, 64), (8192, 32, 0.1, 32)])
def test_ball_query_matches_oracle(pcoe, cuda, N, S, r, ns):
    xyz = _unit_clouds(5 * N, 2, N)
    new_xyz = xyz[:, :S].contiguous()
    q = new_xyz.clone()
    q[0, 0] = 50.0                                             # a centroid with no point in range -> row of N
    got = pcoe.ball_query(r, ns, xyz.to(cuda), q.to(cuda)).cpu().numpy()
    want = sampling.ball_query(r, ns, xyz.numpy(), q.numpy())
    assert np.array_equal(got, want)
    assert (got[0, 0] == N).all()


def test_knn_golden_sets(pcoe, golden, cuda):
    g = golden("sampling")
    for tag in ("a", "b"):
        xyz = torch.from_numpy(g[f"{tag}_xyz"]).to(cuda)
        new_xyz = pcoe.index_points(xyz, torch.from_numpy(g[f"{tag}_fps"]).to(cuda))
        got = pcoe.query_ball_point(new_xyz, xyz, 32).cpu().numpy()
        assert got.dtype == np.int64
        assert np.array_equal(np.sort(got, -1), g[f"{tag}_knn_sorted"])


@pytest.mark.parametrize("N,S,K", [(1024, 128, 32), (2048, 128, 32), (8192, 128, 32), (128, 32, 32), (500, 17, 64), (64, 8, 7)])
def test_knn_matches_oracle_sets(pcoe, cuda, N, S, K):
    B = 4
    xyz = _unit_clouds(N + K, B, N)
    perm = torch.stack([torch.randperm(N, generator=torch.Generator().manual_seed(b))[:S] for b in range(B)])
    new_xyz = torch.gather(xyz, 1, perm.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    got = pcoe.query_ball_point(new_xyz.to(cuda), xyz.to(cuda), K).cpu().numpy()
    want, margin = sampling.knn(new_xyz.numpy(), xyz.numpy(), K)
    n, eq, tie, bad = sampling.knn_rows_match(got, want, margin)
    assert bad == 0, f"{bad} rows differ beyond fp32 near-ties"
    assert tie <= max(1, n // 1000)
    # properties: the centroid itself is a member; rows come out in ascending point index (this kernel's order)
    assert (got == perm.numpy()[..., None]).any(-1).all()
    assert (got[..., 1:] > got[..., :-1]).all()


def test_knn_ties_duplicates_and_clustered_clouds(pcoe, cuda):
    """Exact tie rule (lowest index among equal distances) and the candidate-overflow fallback of the K<=32
    kernel: clouds made of a few distinct points repeated many times, so that hundreds of distances are equal."""
    g = torch.Generator().manual_seed(9)
    for N, K, distinct in [(1024, 32, 5), (2048, 32, 3), (300, 16, 40), (1024, 32, 600)]:
        B, S = 3, 16
        base = torch.randn(B, distinct, 3, generator=g)
        pick = torch.randint(0, distinct, (B, N), generator=g)
        xyz = torch.gather(base, 1, pick.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        new_xyz = xyz[:, :S].contiguous()
        got = pcoe.query_ball_point(new_xyz.to(cuda), xyz.to(cuda), K).cpu().numpy()
        x, c = xyz.numpy(), new_xyz.numpy()
        d = (x[:, None, :, :] - c[:, :, None, :]).astype(np.float32)
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]          # the kernel's fp32 arithmetic
        order = np.lexsort((np.broadcast_to(np.arange(N), d2.shape), d2), axis=-1)[..., :K]   # by (distance, index)
        assert np.array_equal(got, np.sort(order, -1)), (N, K, distinct)


def test_gather_points_and_index_points_3d(pcoe, cuda):
    pts = torch.randn(3, 50, 7)
    idx2 = torch.randint(0, 50, (3, 9))
    idx3 = torch.randint(0, 50, (3, 4, 5))
    b = torch.arange(3)
    assert torch.equal(pcoe.index_points(pts.to(cuda), idx2.to(cuda)).cpu(), pts[b[:, None], idx2])
    assert torch.equal(pcoe.index_points(pts.to(cuda), idx3.to(cuda)).cpu(), pts[b[:, None, None], idx3])


def test_random_subset_properties(pcoe, cuda):
    B, N, S = 64, 1024, 128
    a = pcoe.ops.random_subset(B, N, S, 42, 1, cuda).cpu().numpy()
    b = pcoe.ops.random_subset(B, N, S, 42, 1, cuda).cpu().numpy()
    c = pcoe.ops.random_subset(B, N, S, 42, 2, cuda).cpu().numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)   # counter-based: reproducible per (seed, offset)
    assert a.min() >= 0 and a.max() < N
    assert all(len(set(r.tolist())) == S for r in a)           # without replacement
    many = np.concatenate([pcoe.ops.random_subset(B, N, S, 7, o, cuda).cpu().numpy().ravel() for o in range(40)])
    hist = np.bincount(many, minlength=N)
    expect = many.size / N
    assert abs(hist.mean() - expect) < 1e-9 and hist.std() < 3.0 * np.sqrt(expect)   # uniform marginals
    full = pcoe.ops.random_subset(2, 37, 37, 1, 0, cuda).cpu().numpy()
    assert sorted(full[0].tolist()) == list(range(37))         # S == N is a full permutation
    # fused gather: same draw, and new_xyz is exactly xyz[b, idx]
    xyz = torch.randn(B, N, 3, device=cuda)
    idx, nx = pcoe.ops.random_subset(B, N, S, 42, 1, cuda, xyz=xyz)
    assert np.array_equal(idx.cpu().numpy(), a)
    assert torch.equal(nx, torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)))


def test_shape_errors_raise(pcoe, cuda):
    xyz = torch.zeros(2, 16, 3, device=cuda)
    with pytest.raises(ValueError):
        pcoe.query_ball_point(xyz[:, :4], xyz, 32)             # k > N: topk would raise in the reference
    with pytest.raises(ValueError):
        pcoe.ops.farthest_point_sample(torch.zeros(2, 16, 4, device=cuda), 4)
