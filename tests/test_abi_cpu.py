"""The C-ABI library loads on a machine without a GPU and exports what include/pcoe.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pcoe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcoe_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pcoe):
    lib = pcoe._lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libpcoe.so does not export {n}"
    assert set(names) == set(pcoe._lib.SIGNATURES), "ctypes prototypes and header disagree"


def test_version_and_launch_counter(pcoe):
    lib = pcoe._lib.load()
    assert lib.pcoe_version() == 100
    assert lib.pcoe_launch_count() >= 0


def test_argument_errors_without_gpu(pcoe):
    """Shape/NULL validation happens before any CUDA call, so it is testable on CPU."""
    L, lib = pcoe._lib, pcoe._lib.load()
    assert lib.pcoe_fps_f32(None, 0, 10, 4, None, None, None, None) == L.ERR_BAD_SHAPE
    assert b"fps" in lib.pcoe_last_error()
    assert lib.pcoe_fps_f32(None, 1, 10, 4, None, None, None, None) == L.ERR_NULL
    assert lib.pcoe_knn_f32(None, None, 1, 8, 2, 16, None, None) == L.ERR_BAD_SHAPE      # K > N
    assert lib.pcoe_knn_f32(None, None, 1, 4096, 2, 256, None, None) == L.ERR_UNSUPPORTED
    assert lib.pcoe_mvm_match_fwd_bwd(None, None, None, None, 3, None, 4, 5, None, None, None, None, None, None) == L.ERR_UNSUPPORTED
    assert lib.pcoe_soft_ce_fwd_bwd(None, None, 4, 0, None, None, None) == L.ERR_BAD_SHAPE
    # entry points added for the fused subset+gather and the trunk tail
    assert lib.pcoe_random_subset_xyz(2, 8, 9, 1, 0, None, None, None, None, None) == L.ERR_BAD_SHAPE      # S > N
    assert lib.pcoe_random_subset_xyz(2, 8, 4, 1, 0, None, None, None, None, None) == L.ERR_NULL
    assert lib.pcoe_ln_relu_dropout_heads_fwd(None, 1, None, None, None, None, 0, 256, 1e-5, 0.1, 1, 0, None, None, None, None,
                                              None, 3, None, None, None, None, 4, 0.7, 80.0, 1, None, None, None, None) == L.ERR_BAD_SHAPE
    assert lib.pcoe_heads_ln_relu_dropout_bwd(None, 4, 0.7, 80.0, 1, None, None, None, None, 3, None, None, None, None, None,
                                              None, None, None, 8, 256, 0.1, 1, None, None, None, None, None) == L.ERR_NULL
    # round-2 entry points: square_distance, multi-scale ball query, pointwise MLP stack (vanilla PointNet)
    import ctypes as C
    assert lib.pcoe_square_distance_f32(None, None, 1, 0, 4, 3, None, None) == L.ERR_BAD_SHAPE
    assert lib.pcoe_square_distance_f32(None, None, 1, 4, 4, 3, None, None) == L.ERR_NULL
    r, k, o = (C.c_double * 5)(*[0.1] * 5), (C.c_int * 5)(*[8] * 5), (C.c_void_p * 5)()
    assert lib.pcoe_ball_query_multi_f32(None, None, 1, 64, 4, 5, r, k, o, None) == L.ERR_UNSUPPORTED    # > 4 scales
    assert lib.pcoe_ball_query_multi_f32(None, None, 1, 64, 4, 2, r, k, o, None) == L.ERR_NULL
    pd = L.PointMlpDesc(M=2048, rows_per_cloud=1024, D=0, use_xyz=1, nlayers=3, C=(C.c_int32 * 3)(64, 128, 1024), relu_last=1, eps=1e-5)
    assert lib.pcoe_pointmlp_workspace_bytes(C.byref(pd)) > 2048 * (64 + 128) * 4
    pd.rows_per_cloud = 1000                                                  # not a multiple of 32 / does not divide M
    assert lib.pcoe_pointmlp_workspace_bytes(C.byref(pd)) == 0
    assert lib.pcoe_pointmlp_forward(C.byref(pd), None, None, None, None, None, 0, None) == L.ERR_BAD_SHAPE
    pd.rows_per_cloud, pd.nlayers = 1024, 4
    assert lib.pcoe_pointmlp_forward(C.byref(pd), None, None, None, None, None, 0, None) == L.ERR_UNSUPPORTED
    assert lib.pcoe_pointwise_linear_f32(None, 8, 9, None, None, None, 64, 1, None, None) == L.ERR_UNSUPPORTED
    # gradient exchange over peer memory: rank / world, NULL tables, 16-byte granularity
    two = (C.c_void_p * 2)(0x1000, 0x2000)
    assert lib.pcoe_peer_allreduce_f32(two, two, None, 2, 2, 0, 8, 0x1000, 0, None) == L.ERR_BAD_SHAPE     # rank >= world
    assert lib.pcoe_peer_allreduce_f32(two, two, None, 0, 9, 0, 8, 0x1000, 0, None) == L.ERR_BAD_SHAPE     # > 8 ranks
    assert lib.pcoe_peer_allreduce_f32(None, two, None, 0, 2, 0, 8, 0x1000, 0, None) == L.ERR_NULL
    assert lib.pcoe_peer_allreduce_f32(two, two, None, 0, 2, 2, 8, 0x1000, 0, None) == L.ERR_BAD_SHAPE     # offset % 4
    assert lib.pcoe_peer_allreduce_f32(two, two, None, 0, 2, 0, 0, 0x1000, 0, None) == L.OK                # n == 0: nothing to do
    with pytest.raises(ValueError):
        L.check(L.ERR_BAD_SHAPE)
    with pytest.raises(NotImplementedError):
        L.check(L.ERR_UNSUPPORTED)
    with pytest.raises(RuntimeError):
        L.check(L.ERR_CUDA)


def test_sa_descriptor_validation_and_sizes(pcoe):
    L, lib = pcoe._lib, pcoe._lib.load()
    d = L.SADesc(B=4, N=1024, S=128, K=32, D=0, C1=64, C2=64, C3=128, group_all=0, train=1, precision=0,
                 eps=1e-5, momentum=0.1)
    sv, ws = lib.pcoe_sa_saved_bytes(ctypes.byref(d)), lib.pcoe_sa_workspace_bytes(ctypes.byref(d))
    M = 4 * 128 * 32
    assert sv >= M * (64 + 64 + 128) * 4 and sv % 256 == 0
    assert ws >= M * (64 + 64) * 4                      # dz1, dz2 of the backward pass
    d.precision = 1
    assert lib.pcoe_sa_saved_bytes(ctypes.byref(d)) < sv  # bf16 activations
    d.train = 0
    assert lib.pcoe_sa_saved_bytes(ctypes.byref(d)) == 0
    d.K = 24                                            # not a power of two
    assert lib.pcoe_sa_workspace_bytes(ctypes.byref(d)) == 0
    assert lib.pcoe_sa_forward(ctypes.byref(d), None, None, None, None, None, None, None, 0, None, 0, None) == L.ERR_UNSUPPORTED
    d.K, d.group_all = 32, 1                            # group_all needs S == 1, K == N
    assert lib.pcoe_sa_forward(ctypes.byref(d), None, None, None, None, None, None, None, 0, None, 0, None) == L.ERR_BAD_SHAPE
    d2 = L.SADesc(B=1, N=1, S=1, K=1, D=0, C1=8, C2=8, C3=8, group_all=1, train=1, precision=0, eps=1e-5, momentum=0.1)
    assert lib.pcoe_sa_forward(ctypes.byref(d2), None, None, None, None, None, None, None, 0, None, 0, None) == L.ERR_BAD_SHAPE
    assert b"more than 1 value per channel" in lib.pcoe_last_error()


def test_product_has_no_cpu_path(pcoe):
    import torch
    sa = pcoe.PointNetSetAbstraction(16, 8, 0, [8, 8, 16])
    with pytest.raises(RuntimeError, match="no CPU"):
        sa(torch.zeros(2, 32, 3), None)
    with pytest.raises(RuntimeError, match="CUDA"):
        pcoe.kl_von_mises(torch.zeros(3), torch.ones(3), torch.zeros(3), torch.ones(3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "3d-pointcloud-orientation-estimation_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm: the oracle port of the reference step on the host
    cores) runs without a GPU and prints ONE JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "clouds/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
