"""ctypes binding of the C ABI in include/b2y.h (libb2y.so).  There is NO fallback: if the library is
missing or fails to load, importing/using the CUDA path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb2y.so")
ABI_VERSION = 7

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM = 0, -1, -2, -3, -4
BF16, F32 = 0, 1
(OP_CONV, OP_DWCONV, _OP_RESERVED3, OP_SPPF_POOL, OP_RESAMPLE, OP_ATTN, OP_ROWCOL_MEAN, OP_ELA_GATE, OP_ELA_APPLY, OP_GAP4,
 OP_MSCA_GATE, OP_WSUM4, OP_DECODE, OP_CONV_PAIR, _OP_RESERVED15, OP_PMSFA_DW, OP_MSCA_BRANCH, OP_D2S) = range(1, 19)
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
RES_NONE, RES_ADD, RES_MUL, RES_PREADD_UP2 = 0, 1, 2, 3
OP_NAMES = {1: "conv", 2: "dwconv", 4: "sppf_pool", 5: "resample", 6: "attn", 7: "rowcol_mean", 8: "ela_gate",
            9: "ela_apply", 10: "gap4", 11: "msca_gate", 12: "wsum4", 13: "decode", 14: "conv_pair", 16: "pmsfa_dw", 17: "msca_branch", 18: "d2s"}


class Buf(C.Structure):
    _fields_ = [("offset", C.c_int64), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("dtype", C.c_int32)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in_", C.c_int32 * 4), ("in_coff", C.c_int32 * 4), ("out", C.c_int32),
                ("out_coff", C.c_int32), ("res", C.c_int32), ("res_coff", C.c_int32), ("res_mode", C.c_int32),
                ("cin", C.c_int32), ("cout", C.c_int32), ("cout_pad", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("sh", C.c_int32), ("sw", C.c_int32), ("ph", C.c_int32), ("pw", C.c_int32), ("act", C.c_int32),
                ("impl", C.c_int32), ("w_off", C.c_int64), ("b_off", C.c_int64), ("p_off", C.c_int64 * 4),
                ("ip", C.c_int32 * 8), ("fp", C.c_float * 4)]


EXPORTS = ["b2y_abi_version", "b2y_last_error", "b2y_plan_create", "b2y_plan_destroy", "b2y_plan_run",
           "b2y_plan_graph_build", "b2y_plan_graph_launch", "b2y_plan_set_mid_op", "b2y_plan_wait_mid", "b2y_plan_autotune", "b2y_plan_num_launches", "b2y_plan_op_impl", "b2y_conv_pair_fits", "b2y_preprocess",
           "b2y_decode", "b2y_nms_workspace_bytes", "b2y_nms", "b2y_scale_boxes", "b2y_preprocess_chw", "b2y_box_iou",
           "b2y_match_predictions", "b2y_gather_extras", "b2y_process_mask"]

_lib = None


class B2YError(RuntimeError):
    pass


def lib():
    """Load libb2y.so (once). Raises if it has not been built -- the product path has no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2YError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t
    L.b2y_abi_version.restype = C.c_int
    L.b2y_last_error.restype = C.c_char_p
    L.b2y_plan_create.argtypes = [C.POINTER(Op), i32, C.POINTER(Buf), i32, vp, sz, vp, sz, i32, i32, C.POINTER(vp)]
    L.b2y_plan_destroy.argtypes = [vp]
    L.b2y_plan_run.argtypes = [vp, i32, i32, vp]
    L.b2y_plan_graph_build.argtypes = [vp, vp]
    L.b2y_plan_graph_launch.argtypes = [vp, vp]
    L.b2y_plan_autotune.argtypes = [vp, vp]
    L.b2y_plan_set_mid_op.argtypes = [vp, i32]
    L.b2y_plan_wait_mid.argtypes = [vp, vp]
    L.b2y_plan_num_launches.argtypes = [vp]
    L.b2y_plan_op_impl.argtypes = [vp, i32]
    L.b2y_conv_pair_fits.argtypes = [i32, i32, i32, i32, i32]
    L.b2y_preprocess.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    L.b2y_decode.argtypes = [C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(f32), i32, i32, i32, vp, vp]
    L.b2y_nms_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.b2y_nms_workspace_bytes.restype = sz
    L.b2y_nms.argtypes = [vp, i32, i32, i32, f32, f64, i32, i32, f32, i32, i32, C.POINTER(i32), i32, vp, sz, vp, vp, vp, vp]
    L.b2y_scale_boxes.argtypes = [vp, vp, vp, i32, i32, vp]
    L.b2y_preprocess_chw.argtypes = [vp, i32, i32, i32, vp, vp]
    L.b2y_box_iou.argtypes = [vp, i32, vp, i32, f32, vp, vp]
    L.b2y_match_predictions.argtypes = [vp, vp, vp, vp, C.POINTER(f32), i32, i32, i32, i32, vp, vp]
    L.b2y_gather_extras.argtypes = [C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(f32), i32, i32, i32, i32, i32, vp, vp,
                                    vp, i32, i32, vp, vp, vp]
    L.b2y_process_mask.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, i32, i32, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("b2y_abi_version",):
            fn.restype = C.c_int
    if L.b2y_abi_version() != ABI_VERSION:
        raise B2YError(f"libb2y.so ABI {L.b2y_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int):
    """Map ABI error codes to the exception types the reference raises for the same situations."""
    if rc == OK:
        return
    msg = (lib().b2y_last_error() or b"").decode(errors="replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise B2YError(msg)
