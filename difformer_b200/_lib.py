"""ctypes binding of libdifformer_b200.so (the C ABI declared in include/difformer_b200.h).

The product path has NO CPU or PyTorch fallback: if the shared library is missing this module
raises at import time, and every op raises when handed a non-CUDA tensor.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdifformer_b200.so")

DIF_IMPL_AUTO, DIF_IMPL_GENERIC, DIF_IMPL_TCGEN05 = 0, 1, 2
DIF_DTYPE_F32, DIF_DTYPE_BF16, DIF_DTYPE_F16 = 0, 1, 2

c_i32, c_i64, c_f64, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


class Epilogue(ctypes.Structure):
    """dif_epilogue_t"""
    _fields_ = [("mode", c_i32), ("attn_scale", ctypes.c_float), ("n_add", c_i32),
                ("add", c_vp * 3), ("add_scale", ctypes.c_float * 3),
                ("ln_weight", c_vp), ("ln_bias", c_vp), ("ln_eps", ctypes.c_float), ("relu", c_i32),
                ("gcn_rowptr", c_vp), ("gcn_idx", c_vp), ("gcn_val", c_vp), ("gcn_x", c_vp), ("gcn_scale", ctypes.c_float)]


# name -> (restype, argtypes); mirrors include/difformer_b200.h one to one
SIGNATURES = {
    "dif_version": (c_i32, []),
    "dif_last_error": (ctypes.c_char_p, []),
    "dif_device_supported": (c_i32, []),
    "dif_simple_partials_len": (c_i64, [c_i32] * 4),
    "dif_simple_workspace_bytes": (c_i64, [c_i64] + [c_i32] * 4),
    "dif_simple_prepared_bytes": (c_i64, [c_i32] * 4),
    "dif_simple_reduce": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dif_simple_apply": (c_i32, [c_vp, c_vp, c_vp, c_f64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, ctypes.POINTER(Epilogue), c_i32, c_vp]),
    "dif_simple_forward_workspace_bytes": (c_i64, [c_i64] + [c_i32] * 4),
    "dif_simple_forward": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, c_f64, c_vp, c_vp, c_vp, c_i64,
                                   ctypes.POINTER(c_vp), c_i32, c_i32, ctypes.c_uint64, c_vp]),
    "dif_simple_project_workspace_bytes": (c_i64, [c_i32]),
    "dif_simple_project": (c_i32, [c_vp] * 7 + [c_f64, c_i32] + [c_vp] * 4 + [c_i64, c_vp]),
    "dif_simple_project_values": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "dif_simple_apply_projected": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_i64, c_i32, c_vp, ctypes.POINTER(Epilogue), c_vp]),
    "dif_simple_bwd_partials_len": (c_i64, [c_i32] * 3),
    "dif_simple_bwd_rowscal_len": (c_i64, [c_i64, c_i32, c_i32, c_i32, c_i32]),
    "dif_simple_bwd_reduce": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_f64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dif_simple_bwd_apply": (c_i32, [c_vp] * 8 + [c_f64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp]),
    "dif_sumsq2": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "dif_segmented_simple_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "dif_segmented_simple_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32,
                                         c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "dif_segmented_simple_bwd_phase": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32,
                                               c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dif_segmented_workspace_bytes": (c_i64, [c_i32]),
    "dif_segmented_plan_bytes": (c_i64, [c_i64, c_i32]),
    "dif_segmented_plan_build": (c_i32, [c_vp, c_i32, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "dif_segmented_simple_bwd_tc": (c_i32, [c_vp] * 6 + [c_i64, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dif_segmented_simple_fwd_tc": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "dif_sigmoid_set_impl": (c_i32, [c_i32]),
    "dif_sigmoid_fwd_workspace_bytes": (c_i64, [c_i64, c_i64, c_i32, c_i32, c_i32, c_i32]),
    "dif_sigmoid_fwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "dif_sigmoid_bwd": (c_i32, [c_vp] * 6 + [c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "dif_sigmoid_bwd_workspace_bytes": (c_i64, [c_i64, c_i64, c_i32, c_i32, c_i32, c_i32]),
    "dif_csr_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "dif_csr_build": (c_i32, [c_vp, c_vp, c_i64, c_i64] + [c_vp] * 7 + [c_vp, c_i64, c_vp]),
    "dif_gcn_spmm": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "dif_head_mean": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "dif_simple_reduce_allreduce": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64,
                                            ctypes.POINTER(c_vp), c_i32, c_i32, ctypes.c_uint64, c_vp]),
    "dif_comm_buffer_bytes": (c_i64, [c_i64]),
    "dif_comm_alloc": (c_i32, [ctypes.POINTER(c_vp), c_i64]),
    "dif_comm_free": (c_i32, [c_vp]),
    "dif_comm_export": (c_i32, [c_vp, c_vp]),
    "dif_comm_open": (c_i32, [c_vp, ctypes.POINTER(c_vp)]),
    "dif_comm_close": (c_i32, [c_vp]),
    "dif_comm_status": (c_i32, [c_vp, ctypes.POINTER(c_i32)]),
    "dif_comm_reset": (c_i32, [c_vp]),
    "dif_comm_allreduce": (c_i32, [ctypes.POINTER(c_vp), c_i32, c_i32, c_i64, ctypes.c_uint64, c_vp, c_vp, c_vp]),
}


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"difformer_b200: {LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C difformer_b200/csrc`. There is no CPU/PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


class DifformerError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.dif_last_error().decode("utf-8", "replace")
        kind = {-1: "bad argument", -2: "unsupported shape", -3: "CUDA error"}.get(rc, f"code {rc}")
        raise DifformerError(f"{what or 'difformer_b200'}: {kind}: {msg}")
