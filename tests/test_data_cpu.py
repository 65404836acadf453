"""Host side of the device input pipeline: file parsers with the reference loaders' semantics, the binary cache format,
and the resampler oracle's statistics against np.random.choice (dataloader_multi_peak_vonMises.py:6-67,
dataloader_single_peak_vonMises.py:36-45, dataloader_8dir_sampled.py:48-56)."""
import importlib.util
import os
import struct
import json

import numpy as np
import pytest

from oracle import data as odata

PLY = """ply
format ascii 1.0
element vertex 4
property float x
property float y
property float z
property float nx
end_header
0.5 -1.25 2.0 0.1
1e-3 2 3 0.2
-4 5.5 6 0.3
7 8 -9.75 0.4
"""
MVM = """# multi-peak von Mises ground truth
K 2
mu kappa weight
1.5707964 8.0 0.5
-1.5707964 8.0 0.5
"""


@pytest.fixture()
def files(tmp_path):
    (tmp_path / "a.ply").write_text(PLY)
    (tmp_path / "a_multi_peak_vM_gt.txt").write_text(MVM)
    (tmp_path / "a_single_peak_vM_gt.txt").write_text("# mu kappa\n0.25 -3.0\n")
    (tmp_path / "a_8dir.txt").write_text("0.1 0.2 0.3 0.4\n0.0 0.0 0.0 0.0\n")
    return tmp_path


def test_parsers_known_answers(pcoe, files):
    d = pcoe.data
    pts = d.read_ply(files / "a.ply")
    assert pts.dtype == np.float32 and pts.shape == (4, 3)
    assert np.array_equal(pts, np.array([[0.5, -1.25, 2.0], [1e-3, 2, 3], [-4, 5.5, 6], [7, 8, -9.75]], np.float32))
    t, K = d.read_mvm_gt(files / "a_multi_peak_vM_gt.txt", 4)
    assert K == 2 and t.shape == (4, 3) and np.allclose(t[:2], [[1.5707964, 8, 0.5], [-1.5707964, 8, 0.5]]) and not t[2:].any()
    assert d.read_vm_gt(files / "a_single_peak_vM_gt.txt") == (0.25, 0.0)          # negative kappa clamps to 0
    assert d.read_vm_gt(files / "missing.txt") == (0.0, 0.0)
    assert np.allclose(d.read_8dir_gt(files / "a_8dir.txt"), [0.1, 0.2, 0.3, 0.4, 0, 0, 0, 0])
    assert np.allclose(d.read_8dir_gt(files / "missing.txt"), 0.125) and np.allclose(d.read_8dir_gt(files / "a_8dir.txt", True), 0.125)
    with pytest.raises(RuntimeError):
        (files / "bad.txt").write_text("K 1\n")
        d.read_mvm_gt(files / "bad.txt")


@pytest.mark.skipif(not os.path.exists("/root/reference/dataloader_multi_peak_vonMises.py"), reason="reference tree not present")
def test_parsers_equal_the_reference_loaders(pcoe, files):
    def load(name):
        spec = importlib.util.spec_from_file_location(name, f"/root/reference/{name}.py")
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
        return m
    mv, sp = load("dataloader_multi_peak_vonMises"), load("dataloader_single_peak_vonMises")
    assert np.array_equal(mv.read_ply(str(files / "a.ply")), pcoe.data.read_ply(files / "a.ply"))
    rt, rK = mv.PointCloudDatasetMvM._read_mvM(str(files / "a_multi_peak_vM_gt.txt"), 4)
    t, K = pcoe.data.read_mvm_gt(files / "a_multi_peak_vM_gt.txt", 4)
    assert rK == K and np.array_equal(rt.numpy(), t)
    assert sp.PointCloudDatasetVonMises._read_vm(str(files / "a_single_peak_vM_gt.txt")) == pcoe.data.read_vm_gt(files / "a_single_peak_vM_gt.txt")
    ds = mv.PointCloudDatasetMvM([(str(files / "a.ply"), str(files / "a_multi_peak_vM_gt.txt"), "chair")], 3)
    assert ds.label_map == {"chair": 0}


def test_cache_file_layout(pcoe, files):
    samples = [(str(files / "a.ply"), str(files / "a_multi_peak_vM_gt.txt"), "sofa"),
               (str(files / "a.ply"), str(files / "a_multi_peak_vM_gt.txt"), "chair")]
    info = pcoe.data.build_cache(samples, str(files / "c.bin"), kind="mvm")
    assert info["n"] == 2 and info["points"] == 8
    raw = (files / "c.bin").read_bytes()
    assert raw[:8] == b"PCOECACH"
    version, hlen = struct.unpack("<II", raw[8:16])
    meta = json.loads(raw[16:16 + hlen])
    assert version == 1 and meta["label_map"] == {"chair": 0, "sofa": 1} and meta["arrays"]["points"]["shape"] == [8, 3]
    base = (16 + hlen + 63) // 64 * 64
    off = np.frombuffer(raw, np.int64, 3, base + meta["arrays"]["offsets"]["offset"])
    assert off.tolist() == [0, 4, 8]
    lab = np.frombuffer(raw, np.int64, 2, base + meta["arrays"]["labels"]["offset"])
    assert lab.tolist() == [1, 0]                                     # sorted-category label map, as the reference


def test_resampler_oracle_is_a_uniform_subset_like_numpy_choice():
    n, num, trials = 40, 10, 4000
    counts = np.zeros(n)
    for t in range(trials):
        idx = odata.resample_indices(n, num, seed=7, draw=t)
        assert len(set(idx.tolist())) == num and idx.min() >= 0 and idx.max() < n and (np.diff(idx) > 0).all()
        counts[idx] += 1
    # every point is selected with probability num / n (np.random.choice(n, num, replace=False) has the same marginals)
    p = counts / trials
    assert abs(p.mean() - num / n) < 1e-12 and np.abs(p - num / n).max() < 4 * np.sqrt(0.25 * 0.75 / trials)
    rng = np.random.RandomState(0)
    ref = np.zeros(n)
    for t in range(trials):
        ref[rng.choice(n, num, replace=False)] += 1
    assert np.abs(ref / trials - num / n).max() < 4 * np.sqrt(0.25 * 0.75 / trials)
    # n < num: with replacement, every draw uniform on [0, n)
    idx = np.concatenate([odata.resample_indices(5, 64, 3, t) for t in range(500)])
    assert idx.min() == 0 and idx.max() == 4 and np.abs(np.bincount(idx) / idx.size - 0.2).max() < 0.02
    assert (odata.resample_indices(0, 4, 1, 0) == -1).all()
    assert np.array_equal(odata.resample_indices(6, 6, 1, 2), np.arange(6))      # n == num: every point once
