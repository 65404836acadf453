"""Device-side input pipeline (pcoe_resample_clouds_f32, PointCloudCache, DeviceLoader) against the integer oracle."""
import numpy as np
import pytest
import torch

from oracle import data as odata

pytestmark = pytest.mark.gpu


def _ragged(seed, sizes):
    g = np.random.RandomState(seed)
    return [g.randn(n, 3).astype(np.float32) for n in sizes]


def test_resample_indices_bit_exact_vs_oracle(pcoe, cuda):
    sizes = [1, 7, 100, 1024, 1025, 5000, 20000, 0, 3000]
    clouds = _ragged(0, sizes)
    cache = pcoe.data.PointCloudCache.from_arrays(clouds, np.zeros((len(sizes), 2)), kind="vm", device=cuda)
    ids = torch.tensor([8, 0, 1, 2, 3, 4, 5, 6, 7, 3, 3])
    for num, seed, draw in ((1024, 5, 0), (64, 9, 1000), (4096, 1, 7)):
        xyz, tgt, lab, idx = cache.batch(ids, num, seed=seed, return_idx=True, draw=draw)
        idx = idx.cpu().numpy()
        for b, c in enumerate(ids.tolist()):
            want = odata.resample_indices(sizes[c], num, seed, draw + b)
            assert np.array_equal(idx[b], want), (num, b, c)
            if sizes[c]:
                assert np.array_equal(xyz[b].cpu().numpy(), clouds[c][want])       # the gathered coordinates, bit for bit
            else:
                assert not xyz[b].any()
    # the same cloud in two slots of one batch gets two different subsets; a repeated call (running counter) differs too
    assert not np.array_equal(idx[4], idx[9])
    a = cache.batch(ids, 256, seed=3)[0]
    b_ = cache.batch(ids, 256, seed=3)[0]
    assert not torch.equal(a, b_)


def test_resample_device_counter_advances_the_stream(pcoe, cuda):
    clouds = _ragged(1, [4000, 4000])
    cache = pcoe.data.PointCloudCache.from_arrays(clouds, np.zeros((2, 8)), kind="8dir", device=cuda)
    ids = torch.tensor([0, 1])
    ctr = torch.zeros(1, dtype=torch.int64, device=cuda)
    i0 = cache.batch(ids, 512, seed=2, counter=ctr, return_idx=True, draw=10)[-1].cpu().numpy()
    ctr += 3
    i3 = cache.batch(ids, 512, seed=2, counter=ctr, return_idx=True, draw=10)[-1].cpu().numpy()
    assert np.array_equal(i0[1], odata.resample_indices(4000, 512, 2, 11))
    assert np.array_equal(i3[0], odata.resample_indices(4000, 512, 2, 10 + 3 * 2))
    assert np.array_equal(i3[1], odata.resample_indices(4000, 512, 2, 10 + 3 * 2 + 1))


def test_cache_file_roundtrip_and_loader_epoch(pcoe, cuda, tmp_path):
    rs = np.random.RandomState(4)
    samples = []
    for i in range(10):
        n = int(rs.randint(50, 400))
        pts = rs.randn(n, 3).astype(np.float32)
        ply = tmp_path / f"m{i}.ply"
        ply.write_text("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nend_header\n" % n
                       + "\n".join(" ".join(repr(float(v)) for v in p) for p in pts) + "\n")
        gt = tmp_path / f"m{i}_multi_peak_vM_gt.txt"
        K = [1, 2, 4][i % 3]
        gt.write_text(f"K {K}\nmu kappa w\n" + "".join(f"{0.1 * j} 8.0 {1.0 / K}\n" for j in range(K)))
        samples.append((str(ply), str(gt), f"cat{i % 4}"))
    pcoe.data.build_cache(samples, str(tmp_path / "ds.bin"), kind="mvm")
    cache = pcoe.data.PointCloudCache.load(str(tmp_path / "ds.bin"), cuda)
    assert len(cache) == 10 and cache.targets.shape == (10, 4, 3) and cache.K.tolist() == [[1, 2, 4][i % 3] for i in range(10)]
    p0 = pcoe.data.read_ply(samples[3][0])
    o = cache.offsets.cpu().numpy()
    assert np.array_equal(cache.points[o[3]:o[4]].cpu().numpy(), p0)
    loader = pcoe.data.DeviceLoader(cache, batch_size=4, num_points=128, shuffle=True, seed=1)
    seen = []
    for xyz, vm, K, lab in loader:
        assert xyz.is_cuda and xyz.shape[1:] == (128, 3) and vm.shape[1:] == (4, 3)
        seen += lab.tolist()
        # every output point is a point of its cloud
        assert torch.isfinite(xyz).all()
    assert len(seen) == 10 and len(loader) == 3
    # the batch feeds the model's loss directly: (mu,kappa,w) rows and K as the reference's collate gives them
    model = pcoe.PointNetPPMvM().to(cuda).train()
    xyz, vm, K, lab = next(iter(loader))
    mu, kappa, w = model(torch.cat([xyz, xyz], 0))
    loss = pcoe.match_loss(mu, kappa, w, torch.cat([vm, vm], 0), None, torch.cat([K, K], 0)).mean()
    assert torch.isfinite(loss)
