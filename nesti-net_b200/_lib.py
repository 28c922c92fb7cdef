"""ctypes binding of libmups_b200.so (the C ABI declared in include/mups.h).

There is no CPU fallback: if the library has not been built (``__graft_entry__.build()`` or
``python nesti-net_b200/build.py``) importing the compute entry points raises, and every call
fails loudly when no CUDA device is usable.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmups_b200.so")

MUPS_OK = 0
MUPS_ERR_INVALID = -1
MUPS_ERR_CUDA = -2
MUPS_ERR_NOMEM = -3
MUPS_ERR_UNSUPPORTED = -4

FLAG_MASKED = 1
LAYOUT_MUPS = 0
LAYOUT_CHANNEL = 2
FLAG_NO_FASTPATH = 4
FLAG_WIDE_STORES = 8
MAX_SCALES = 8
MAX_POINTS_PER_PATCH = 2048

# every symbol include/mups.h declares (tests check the library exports all of them)
EXPORTS = (
    "mups_abi_version", "mups_last_error", "mups_launch_count", "mups_set_option",
    "mups_index_create", "mups_index_bbox", "mups_index_size", "mups_index_destroy",
    "mups_ball_query",
    "mups_gmm_create", "mups_gmm_size", "mups_gmm_is_separable", "mups_gmm_destroy",
    "mups_3dmfv", "mups_features", "mups_ball_query_select", "mups_3dmfv_selected",
    "mups_moe_pack_input", "mups_conv3d_bn_relu", "mups_pool3d", "mups_conv1_split_bn_relu", "mups_avgpool3d_bn_relu",
    "mups_split_bf16x3", "mups_pool3d_bf16x3", "mups_avgpool3d_f32_bn_relu_x3", "mups_conv3d_bn_relu_x3",
)

_lib = None


def load():
    """Load (once) and return the ctypes handle of libmups_b200.so."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmups_b200.so is not built (%s missing): run __graft_entry__.build(); "
            "this package has no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, u32, u64, dbl = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint32,
                                   ctypes.c_uint64, ctypes.c_double)
    fp = ctypes.POINTER(ctypes.c_float)
    dp = ctypes.POINTER(ctypes.c_double)
    L.mups_abi_version.restype = i32
    L.mups_last_error.restype = ctypes.c_char_p
    L.mups_launch_count.restype = i64
    L.mups_set_option.argtypes = [ctypes.c_char_p, i64]
    L.mups_set_option.restype = i32
    L.mups_index_create.argtypes = [ctypes.POINTER(vp), vp, i64, dbl, vp]
    L.mups_index_create.restype = i32
    L.mups_index_bbox.argtypes = [vp, fp, fp]
    L.mups_index_bbox.restype = i32
    L.mups_index_size.argtypes = [vp]
    L.mups_index_size.restype = i64
    L.mups_index_destroy.argtypes = [vp]
    L.mups_index_destroy.restype = None
    L.mups_ball_query.argtypes = [vp, vp, i64, dp, i32, i32, u64, vp, vp, vp, vp, vp]
    L.mups_ball_query.restype = i32
    L.mups_gmm_create.argtypes = [ctypes.POINTER(vp), fp, fp, fp, i32]
    L.mups_gmm_create.restype = i32
    L.mups_gmm_size.argtypes = [vp]
    L.mups_gmm_size.restype = i32
    L.mups_gmm_is_separable.argtypes = [vp]
    L.mups_gmm_is_separable.restype = i32
    L.mups_gmm_destroy.argtypes = [vp]
    L.mups_gmm_destroy.restype = None
    L.mups_3dmfv.argtypes = [vp, vp, vp, i64, i32, i32, u32, vp, vp]
    L.mups_3dmfv.restype = i32
    L.mups_features.argtypes = [vp, vp, vp, i64, dp, i32, i32, u64, u32, vp, vp, vp, vp, vp]
    L.mups_features.restype = i32
    L.mups_ball_query_select.argtypes = [vp, vp, i64, dp, i32, i32, u64, vp, vp, vp, vp]
    L.mups_ball_query_select.restype = i32
    L.mups_3dmfv_selected.argtypes = [vp, vp, vp, i64, dp, i32, i32, vp, vp, u32, vp, vp]
    L.mups_3dmfv_selected.restype = i32
    L.mups_moe_pack_input.argtypes = [vp, i64, i32, vp, vp]
    L.mups_moe_pack_input.restype = i32
    L.mups_conv3d_bn_relu.argtypes = [vp, i64, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp, vp]
    L.mups_conv3d_bn_relu.restype = i32
    L.mups_pool3d.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, vp, vp]
    L.mups_pool3d.restype = i32
    L.mups_conv1_split_bn_relu.argtypes = [vp, i64, i32, i32, i32, i32, vp, i32, i32, vp, vp, i32, vp, i32, i32, i32, vp, i32, i32, vp]
    L.mups_conv1_split_bn_relu.restype = i32
    L.mups_avgpool3d_bn_relu.argtypes = [vp, i64, i32, i32, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp]
    L.mups_avgpool3d_bn_relu.restype = i32
    L.mups_split_bf16x3.argtypes = [vp, i64, i32, i32, i32, vp, i32, i32, i32, vp]
    L.mups_split_bf16x3.restype = i32
    L.mups_pool3d_bf16x3.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, vp, i32, i32, vp]
    L.mups_pool3d_bf16x3.restype = i32
    L.mups_avgpool3d_f32_bn_relu_x3.argtypes = [vp, i64, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp]
    L.mups_avgpool3d_f32_bn_relu_x3.restype = i32
    L.mups_conv3d_bn_relu_x3.argtypes = [vp, i64, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp, i32, vp, i32, i32, i32, vp]
    L.mups_conv3d_bn_relu_x3.restype = i32
    if L.mups_abi_version() != 1:
        raise RuntimeError("libmups_b200.so ABI version %d, expected 1" % L.mups_abi_version())
    _lib = L
    v = os.environ.get("MUPS_STATS_VARIANT")        # benchmarking / A-B testing only (see mups_set_option)
    if v:
        L.mups_set_option(b"stats_variant", int(v))
    return L


def check(rc, what=""):
    """Map a mups_status to the exception the reference would raise (ValueError for bad
    options, e.g. pcpnet_dataset.py:210,340; RuntimeError otherwise)."""
    if rc == MUPS_OK:
        return
    msg = load().mups_last_error().decode("utf-8", "replace")
    if rc == MUPS_ERR_INVALID:
        raise ValueError(msg or what)
    if rc == MUPS_ERR_NOMEM:
        raise MemoryError(msg or what)
    raise RuntimeError(msg or what)


def launch_count():
    return int(load().mups_launch_count())


def set_option(name, value):
    check(load().mups_set_option(name.encode(), int(value)), "mups_set_option")
