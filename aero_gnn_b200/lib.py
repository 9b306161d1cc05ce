"""ctypes binding of libaero_sm100.so (include/aero_gnn.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, the
caller gets a RuntimeError.  `load()` only dlopens; nothing here touches the GPU.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libaero_sm100.so"

AERO_F32, AERO_BF16 = 0, 1
AERO_PATH_SIMT, AERO_PATH_UMMA = 0, 1
AERO_BLOCK_AGG_NO_CLEAR = 1
ACT_CODES = {"relu": 0, "tanh": 1, "sigmoid": 2, "elu": 3, "leaky_relu": 4}

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)


class BlockDesc(C.Structure):
    """Mirror of struct aero_block_desc (include/aero_gnn.h)."""

    _fields_ = [
        ("dtype", C.c_int32), ("path", C.c_int32), ("L", C.c_int32), ("act", C.c_int32),
        ("use_ln", C.c_int32), ("main_f32", C.c_int32), ("has_resid_grad", C.c_int32), ("flags", C.c_int32),
        ("rows", C.c_int64), ("n_nodes", C.c_int64), ("ldp", C.c_int64), ("poff0", C.c_int64), ("poff1", C.c_int64),
        ("main", C.c_void_p), ("main_scale", C.c_void_p), ("resid", C.c_void_p), ("P", C.c_void_p),
        ("idx0", C.c_void_p), ("idx1", C.c_void_p), ("rowptr", C.c_void_p), ("prepared", C.c_void_p),
        ("out", C.c_void_p), ("agg", C.c_void_p),
        ("g_out", C.c_void_p), ("g_agg", C.c_void_p), ("g_main", C.c_void_p), ("g_h0", C.c_void_p),
        ("g_w", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("h0", C.c_void_p),
        ("main_lat", C.c_void_p), ("h_hidden", C.c_void_p * 2),
    ]


class WecDesc(C.Structure):
    """Mirror of struct aero_wec_desc (include/aero_gnn.h)."""

    _fields_ = [
        ("dtype", C.c_int32), ("mean", C.c_int32), ("compute_w", C.c_int32), ("pos_dim", C.c_int32),
        ("N", C.c_int64), ("E", C.c_int64), ("out_dim", C.c_int64), ("ldq", C.c_int64),
        ("Q", C.c_void_p), ("pos", C.c_void_p), ("w1_len", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("rowptr", C.c_void_p), ("src", C.c_void_p), ("dst", C.c_void_p), ("perm", C.c_void_p),
        ("sptr", C.c_void_p), ("sperm", C.c_void_p),
        ("w", C.c_void_p), ("out", C.c_void_p), ("g_out", C.c_void_p), ("g_w_ext", C.c_void_p),
        ("dQ", C.c_void_p), ("g_w", C.c_void_p), ("g_small", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class CopySeg(C.Structure):
    """Mirror of struct aero_copy_seg (include/aero_gnn.h)."""

    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64),
                ("src_ld", C.c_int64), ("dst_ld", C.c_int64), ("src_dtype", C.c_int32), ("dst_dtype", C.c_int32)]


class AdamSeg(C.Structure):
    """Mirror of struct aero_adam_seg (include/aero_gnn.h)."""

    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("master", C.c_void_p),
                ("n", C.c_int64), ("p_dtype", C.c_int32), ("g_dtype", C.c_int32)]


MAX_COPY_SEGS = 48

# name -> (restype, argtypes); every symbol declared in include/aero_gnn.h
SIGNATURES = {
    "aero_last_error": (C.c_char_p, []),
    "aero_version": (C.c_int, []),
    "aero_has_umma": (C.c_int, []),
    "aero_has_umma_bwd": (C.c_int, []),
    "aero_umma_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "aero_last_launch_count": (C.c_int, []),
    "aero_graph_plan_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "aero_graph_plan_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 7 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "aero_sort_pairs_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "aero_sort_pairs_u64": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aero_hash_u64": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "aero_gather_rows": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "aero_segment_reduce": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aero_segment_reduce_ld": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p]),
    "aero_segment_bcast": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "aero_block_prepared_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "aero_block_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "aero_block_workspace_bytes": (C.c_size_t, [C.POINTER(BlockDesc), C.c_int]),
    "aero_block_fwd": (C.c_int, [C.POINTER(BlockDesc), C.c_void_p]),
    "aero_block_bwd": (C.c_int, [C.POINTER(BlockDesc), C.c_void_p]),
    "aero_stride_pool_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "aero_stride_pool_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "aero_coarsen_edges_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "aero_coarsen_edges": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "aero_group_lists_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "aero_group_lists": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "aero_bfs_levels_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "aero_bfs_levels": (C.c_int, [C.c_void_p] * 3 + [C.c_int64] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    "aero_bistride_select_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "aero_bistride_select": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "aero_filter_edges_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "aero_filter_edges": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64] + [C.c_void_p] * 3
                          + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "aero_multi_copy": (C.c_int, [C.POINTER(CopySeg), C.c_int, C.c_void_p]),
    "aero_wgrad_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "aero_wgrad": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                             C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aero_row_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "aero_thin_linear_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                       C.c_void_p]),
    "aero_thin_linear_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "aero_thin_linear_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                       C.c_size_t, C.c_void_p]),
    "aero_mse_workspace_bytes": (C.c_size_t, []),
    "aero_mse_loss_grad": (C.c_int, [C.c_void_p] * 4 + [C.c_int64] * 4 + [C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "aero_adam_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_void_p]),
    "aero_wec_workspace_bytes": (C.c_size_t, [C.POINTER(WecDesc), C.c_int]),
    "aero_wec_fwd": (C.c_int, [C.POINTER(WecDesc), C.c_void_p]),
    "aero_wec_bwd": (C.c_int, [C.POINTER(WecDesc), C.c_void_p]),
}

# hardware probes: a separate debug library (include/aero_gnn_debug.h), never needed by the product path
PROBE_LIB_PATH = PKG / "libaero_probe.so"
PROBE_SIGNATURES = {
    "aero_umma_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "aero_umma_rate_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None
_probe = None


def load_probe() -> C.CDLL:
    """dlopen libaero_probe.so (after the product library it links against)."""
    global _probe
    if _probe is None:
        load()
        if not PROBE_LIB_PATH.exists():
            raise RuntimeError(f"{PROBE_LIB_PATH} is missing: build it with `python -m aero_gnn_b200.build`")
        lib = C.CDLL(str(PROBE_LIB_PATH))
        for name, (res, args) in PROBE_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _probe = lib
    return _probe


def load() -> C.CDLL:
    """dlopen the in-tree library and type every symbol; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m aero_gnn_b200.build` "
            "(there is no CPU / PyTorch fallback for the message-passing path)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().aero_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
