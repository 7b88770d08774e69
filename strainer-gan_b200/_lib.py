"""ctypes binding of libstrainer_b200.so (the C ABI declared in include/strainer_b200.h).

There is no fallback of any kind: if the shared library is missing, or the device is not an
sm_100 GPU, every call raises."""
import ctypes
import os
from ctypes import c_double, c_float, c_int, c_int64, c_size_t, c_uint32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libstrainer_b200.so")

SG_LT, SG_LE, SG_GE, SG_GT = 0, 1, 2, 3
SG_NOT = 4  # OR-ed into a comparison: logical negation (NaN-correct complement)
SG_LERP_NUMPY, SG_LERP_TORCH = 0, 1
SG_CONV_BF16, SG_CONV_BF16X3, SG_CONV_FP16 = 0, 1, 2
SG_LAYOUT_NCHW, SG_LAYOUT_NHWC = 0, 1
SG_SELECT_WS_WORDS = 2048
SG_SELECT_WS_NANCOUNT = 256
SG_SELECT_WS_MINABOVE = 257
SG_SELECT_NUM_PASSES = 4
SG_MOMENT_CHUNK = 4096

P = c_void_p
_SIGS = {
    "sg_version": (c_int, []),
    "sg_last_error_string": (ctypes.c_char_p, []),
    "sg_init": (c_int, [c_int]),
    "sg_sm_count": (c_int, []),
    "sg_launch_count": (ctypes.c_longlong, []),
    "sg_synth_images": (c_int, [P, c_int64, c_int64, c_uint32, P]),
    "sg_u8_normalize": (c_int, [P, c_int64, c_int, c_int64, c_int, P, P, P, P]),
    "sg_host_threads": (c_int, []),
    "sg_host_f32_to_f16": (c_int, [P, c_int64, P, c_int, c_int]),
    "sg_f16_expand": (c_int, [P, c_int64, P, P]),
    "sg_d64_packed_bytes": (c_size_t, [c_int]),
    "sg_d64_pack": (c_int, [P] * 17 + [c_float, c_int, P, P]),
    "sg_d64_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "sg_d64_score": (c_int, [P, c_int64, P, P, c_int, P, P, P, P]),
    "sg_d64_score_train": (c_int, [P, c_int64, P, P, c_int, P, P, P, P, P, P, c_float, c_float, P, P, P, P]),
    "sg_d64_score_status": (c_int, [P, c_int64, P, P, c_int, P, P, P, P, P]),
    "sg_d64_score_train_status": (c_int, [P, c_int64, P, P, c_int, P, P, P, P, P, P, c_float, c_float, P, P, P, P, P]),
    "sg_d64_run_layer": (c_int, [P, c_int64, P, P, c_int, c_int, P, P, P, P]),
    "sg_d64_check": (c_int, [P, P]),
    "sg_d64_read_activation": (c_int, [P, c_int64, c_int, c_int, P, P]),
    "sg_d64_train_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "sg_d64_train_workspace_init": (c_int, [P, c_int64, c_int, P]),
    "sg_d64_train_packed_bytes": (c_size_t, [c_int]),
    "sg_d64_train_pack": (c_int, [P, c_int, P, P]),
    "sg_d64_train_forward": (c_int, [P, c_int64, c_int64, c_int, P, P, P, c_float, c_float, P, P, P, P]),
    "sg_d64_train_backward": (c_int, [P, c_int64, c_int64, c_int, P, P, P, P, P]),
    "sg_d64_train_check": (c_int, [P, P]),
    "sg_d64_train_read": (c_int, [P, c_int64, c_int64, c_int, c_int, P, P]),
    "sg_ae_workspace_bytes": (c_size_t, [c_int64]),
    "sg_ae_score": (c_int, [P, c_int64, P, P, P, P, P]),
    "sg_ae_bf16_workspace_bytes": (c_size_t, [c_int64]),
    "sg_ae_score_bf16": (c_int, [P, c_int64, P, P, P, P, P]),
    "sg_ae_bf16_check": (c_int, [P, P]),
    "sg_ae_tc_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "sg_ae_score_tc": (c_int, [P, c_int64, P, P, c_int, P, P, P]),
    "sg_ae_pack_tc": (c_int, [P, P, c_int, P]),
    "sg_ae_forward_tc": (c_int, [P, c_int64, P, P, c_int, P, P, P]),
    "sg_mlp_workspace_bytes": (c_size_t, [c_int64]),
    "sg_mlp_score": (c_int, [P, c_int64, P, P, P, P, P, P]),
    "sg_mlp_tc_packed_bytes": (c_size_t, []),
    "sg_mlp_tc_workspace_bytes": (c_size_t, [c_int64]),
    "sg_mlp_tc_pack": (c_int, [P, P, P]),
    "sg_mlp_score_tc": (c_int, [P, c_int64, P, P, P, P, P, P, P, P]),
    "sg_d28_packed_bytes": (c_size_t, []),
    "sg_d28_workspace_bytes": (c_size_t, [c_int64]),
    "sg_d28_pack": (c_int, [P, P, P, P, P, P, c_float, P, P]),
    "sg_d28_score": (c_int, [P, c_int64, P, P, P, P, P, P, P, P]),
    "sg_select_begin": (c_int, [P, c_int64, P]),
    "sg_select_hist": (c_int, [P, c_int64, P, c_int, P]),
    "sg_select_step": (c_int, [P, c_int, P]),
    "sg_peer_buffer_bytes": (c_size_t, [c_int]),
    "sg_peer_alloc": (c_int, [c_int, P, P]),
    "sg_peer_open": (c_int, [P, P]),
    "sg_peer_close": (c_int, [P]),
    "sg_peer_free": (c_int, [P]),
    "sg_select_step_peer": (c_int, [P, c_int, P, c_int, c_int, c_uint32, P]),
    "sg_select_finish": (c_int, [P, P, P]),
    "sg_radix_select": (c_int, [P, c_int64, c_int64, P, P, P]),
    "sg_select_workspace_bytes": (c_size_t, [c_int64]),
    "sg_select_kth": (c_int, [P, c_int64, c_int64, P, c_size_t, P, P]),
    "sg_select_check": (c_int, [P, P]),
    "sg_lerp_threshold": (c_int, [P, c_float, c_int, P, P]),
    "sg_segment_order_stats": (c_int, [P, c_int64, c_int, c_int, c_int, P, P]),
    "sg_compact_workspace_bytes": (c_size_t, [c_int64]),
    "sg_compact_indices": (c_int, [P, c_int64, P, c_int, c_int64, P, P, P, P, P]),
    "sg_compact_rows": (c_int, [P, c_int64, c_int64, P, P, P, P, P, P]),
    "sg_strain_rows": (c_int, [P, c_int64, c_int, c_int, c_float, c_int, c_int, P, c_int64, P, P, P, P, P, P, P]),
    "sg_strain_rows_concat": (c_int, [P, c_int64, c_int, c_int, c_float, c_int, c_int, P, c_int64, P, P, P, P, P, P, P]),
    "sg_concat_rows": (c_int, [P, c_int64, P, c_int64, c_int64, P, P]),
    "sg_gather_rows": (c_int, [P, c_int64, P, c_int64, P, P, P]),
    "sg_sort_workspace_bytes": (c_size_t, [c_int64]),
    "sg_sort_f32": (c_int, [P, c_int64, P, P, P, P]),
    "sg_dbscan1d_workspace_bytes": (c_size_t, [c_int64]),
    "sg_dbscan1d": (c_int, [P, c_int64, c_double, c_int, P, P, P, P]),
    "sg_dbscan_nd_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "sg_dbscan_nd": (c_int, [P, c_int64, c_int, P, P, c_double, c_int, P, P, P]),
    "sg_dbscan_nd_check": (c_int, [P, P]),
    "sg_gmm1d_workspace_bytes": (c_size_t, []),
    "sg_gmm1d_begin": (c_int, [P, c_int, P, P]),
    "sg_gmm1d_accumulate": (c_int, [P, c_int64, P, P]),
    "sg_gmm1d_update": (c_int, [c_int64, c_double, c_double, c_int, P, P]),
    "sg_chunk_moments": (c_int, [P, c_int64, P, P]),
    "sg_moments_finish": (c_int, [P, c_int64, c_int64, c_float, P, P, P]),
    "sg_standardize": (c_int, [P, c_int64, P, c_int, P, P]),
    "sg_col_moments_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "sg_col_moments": (c_int, [P, c_int64, c_int, c_int, c_float, P, P, P, P]),
    "sg_row_max_absz": (c_int, [P, c_int64, c_int, P, P, P, P]),
    "sg_minmax": (c_int, [P, c_int64, P, P]),
    "sg_hist_uniform": (c_int, [P, c_int64, P, c_int, P, P]),
}

_OPTIONAL = {"sg_ae_workspace_bytes", "sg_ae_score"}   # only in -DSG_AB_VARIANTS builds (csrc/ae.cu)
_lib = None
_inited_devices = set()


def load():
    """dlopen the library (works without a GPU; compute calls then fail with SG_ENOINIT/SG_EARCH)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built. Run `python __graft_entry__.py` (build()) first; "
                "strainer_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            if name in _OPTIONAL and not hasattr(lib, name):
                continue
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().sg_last_error_string().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"strainer_b200: {what} failed with status {rc}: {last_error()}")


def init(device_index):
    """Prepare the library for a CUDA device (once per process and device; per-device state lives in the library,
    the caller's current device is left untouched)."""
    lib = load()
    if device_index not in _inited_devices:
        check(lib.sg_init(int(device_index)), "sg_init")
        _inited_devices.add(device_index)
    return lib


def call(name, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)
