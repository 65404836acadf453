"""ctypes binding of libpcoe.so (the C ABI declared in include/pcoe.h).

There is no CPU fallback: if the shared library is missing, importing the product path raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``csrc/build.sh``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCOE_LIB") or os.path.join(_HERE, "libpcoe.so")   # PCOE_LIB: debug builds only

OK, ERR_BAD_SHAPE, ERR_UNSUPPORTED, ERR_NULL, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4, -5
PRECISION_FP32, PRECISION_BF16, PRECISION_BF16X3 = 0, 1, 2
VM_SINGLE, VM_MULTI = 0, 1


class SADesc(C.Structure):
    """pcoe_sa_desc"""
    _fields_ = [(n, C.c_int32) for n in
                ("B", "N", "S", "K", "D", "C1", "C2", "C3", "group_all", "train", "precision")] + \
               [("eps", C.c_float), ("momentum", C.c_float)]


class SAParams(C.Structure):
    """pcoe_sa_params"""
    _fields_ = [(n, C.c_void_p * 3) for n in
                ("W", "bias", "gamma", "beta", "running_mean", "running_var", "num_batches_tracked")]


class SAGrads(C.Structure):
    """pcoe_sa_grads"""
    _fields_ = [(n, C.c_void_p * 3) for n in ("dW", "dbias", "dgamma", "dbeta")] + [("accumulate", C.c_int32)]


class PointMlpDesc(C.Structure):
    """pcoe_pointmlp_desc"""
    _fields_ = [(n, C.c_int32) for n in ("M", "rows_per_cloud", "D", "use_xyz", "nlayers")] + \
               [("C", C.c_int32 * 3), ("relu_last", C.c_int32), ("eps", C.c_float)]


_P, _I, _F, _D, _U64, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_uint64, C.c_size_t

# name -> (restype, argtypes); mirrors include/pcoe.h one to one
SIGNATURES = {
    "pcoe_version": (_I, []),
    "pcoe_last_error": (C.c_char_p, []),
    "pcoe_launch_count": (_U64, []),
    "pcoe_profile_enable": (_I, [_I]),
    "pcoe_profile_report": (_I, [C.c_char_p, _SZ]),
    "pcoe_fps_f32": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "pcoe_gather_points_f32": (_I, [_P, _I, _I, _I, _P, _I, _P, _P]),
    "pcoe_host_randperm_subsets": (_I, [_P, _SZ, _I, _I, _I, _P]),
    "pcoe_random_subset": (_I, [_I, _I, _I, _U64, _U64, _P, _P, _P]),
    "pcoe_random_subset_xyz": (_I, [_I, _I, _I, _U64, _U64, _P, _P, _P, _P, _P]),
    "pcoe_square_distance_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "pcoe_knn_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "pcoe_ball_query_f32": (_I, [_P, _P, _I, _I, _I, _I, _D, _P, _P]),
    "pcoe_ball_query_multi_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "pcoe_sa_saved_bytes": (_SZ, [C.POINTER(SADesc)]),
    "pcoe_sa_workspace_bytes": (_SZ, [C.POINTER(SADesc)]),
    "pcoe_sa_forward": (_I, [C.POINTER(SADesc), _P, _P, _P, _P, C.POINTER(SAParams), _P, _P, _SZ,
                             _P, _SZ, _P]),
    "pcoe_sa_backward": (_I, [C.POINTER(SADesc), _P, _P, _P, _P, C.POINTER(SAParams), _P, _P, _P,
                              _SZ, _P, C.POINTER(SAGrads), _P, _SZ, _P]),
    "pcoe_pointmlp_workspace_bytes": (_SZ, [C.POINTER(PointMlpDesc)]),
    "pcoe_pointmlp_forward": (_I, [C.POINTER(PointMlpDesc), _P, _P, C.POINTER(SAParams), _P, _P, _SZ, _P]),
    "pcoe_pointwise_linear_f32": (_I, [_P, _I, _I, _P, _P, _P, _I, _I, _P, _P]),
    "pcoe_resample_clouds_f32": (_I, [_P, _P, _I, _P, _I, _I, _U64, _U64, _P, _P, _P, _P]),
    "pcoe_vm_kl_fwd_bwd": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "pcoe_mvm_match_fwd_bwd": (_I, [_P, _P, _P, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pcoe_soft_ce_fwd_bwd": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "pcoe_mvm_head_fwd": (_I, [_P, _P, _P, _I, _I, _F, _F, _I, _P, _P, _P, _P]),
    "pcoe_mvm_head_bwd": (_I, [_P, _P, _P, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P]),
    "pcoe_linear_fwd": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P]),
    "pcoe_linear_bwd_dx": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "pcoe_linear_bwd_dw": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "pcoe_ln_relu_dropout_fwd": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _F, _F, _I, _U64, _P, _P, _P, _P, _P, _P]),
    "pcoe_ln_relu_dropout_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P, _P, _P, _P, _P]),
    "pcoe_ln_relu_dropout_heads_fwd": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _F, _F, _I, _U64, _P, _P, _P, _P, _P,
                                            _I, _P, _P, _P, _P, _I, _F, _F, _I, _P, _P, _P, _P]),
    "pcoe_heads_ln_relu_dropout_bwd": (_I, [_P, _I, _F, _F, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P,
                                            _I, _I, _F, _I, _P, _P, _P, _P, _P]),
    "pcoe_peer_allreduce_f32": (_I, [_P, _P, _P, _I, _I, _SZ, _SZ, _P, _I, _P]),
    "pcoe_adam_workspace_bytes": (_SZ, []),
    "pcoe_adam_step": (_I, [_P, _P, _P, _P, _SZ, _F, _F, _F, _F, _F, _F, _F, _I, _P, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load libpcoe.so once and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (run "
            "`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Convert a pcoe_status into the exception the reference would have raised."""
    if rc == OK:
        return
    msg = load().pcoe_last_error().decode("utf-8", "replace")
    if rc in (ERR_BAD_SHAPE, ERR_NULL):
        raise ValueError(f"pcoe: {msg}")
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(f"pcoe: {msg}")
    raise RuntimeError(f"pcoe (status {rc}): {msg}")


def launch_count() -> int:
    return int(load().pcoe_launch_count())


profiling = False      # per-kernel CUDA-event timing on: side-stream overlap is switched off (kernels timed alone)


def profile(enable: bool) -> None:
    global profiling
    check(load().pcoe_profile_enable(int(enable)))
    profiling = bool(enable)


def profile_report() -> dict:
    """{kernel name: (launches, total_ms)} since the last report; synchronises on the recorded events."""
    buf = C.create_string_buffer(1 << 16)
    check(load().pcoe_profile_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(",", 2)
        out[name] = (int(n), float(ms))
    return out
