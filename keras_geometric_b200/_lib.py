"""ctypes binding of libkgb200.so (the C ABI declared in include/kgb200.h).

There is no CPU fallback: if the shared library cannot be loaded (or built with nvcc) every
entry point raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

from . import _build

OP_SUM, OP_MEAN, OP_MAX, OP_MIN, OP_MAX_RAW, OP_SQDEV = 0, 1, 2, 3, 4, 5
OPS = {"sum": OP_SUM, "mean": OP_MEAN, "max": OP_MAX, "min": OP_MIN, "max_raw": OP_MAX_RAW}
MAX_OPS = (OP_MAX, OP_MIN, OP_MAX_RAW)
ACT_NONE, ACT_RELU = 0, 1
STATUS_OOB_INDEX = 1


class KgbError(RuntimeError):
    pass


class GatherReduceArgs(Structure):
    """Mirror of ``struct kgb_gather_reduce_args`` (include/kgb200.h)."""

    _fields_ = [
        ("x", c_void_p), ("ldx", c_int64), ("n_src_rows", c_int64), ("F", c_int32), ("op", c_int32),
        ("rowptr", c_void_p), ("col", c_void_p), ("n_rows", c_int64), ("row_ids", c_void_p),
        ("edge_w", c_void_p), ("src_scale", c_void_p), ("out_scale", c_void_p),
        ("addend", c_void_p), ("ld_addend", c_int64), ("addend_scale", c_float),
        ("bias", c_void_p), ("act", c_int32), ("out", c_void_p), ("ldo", c_int64), ("arg", c_void_p),
        ("hub_row", c_void_p), ("hub_chunk_base", c_void_p), ("hub_nchunks", c_void_p),
        ("chunk_hub", c_void_p), ("n_hubs", c_int32), ("n_chunks", c_int32),
        ("hub_threshold", c_int32), ("hub_chunk", c_int32), ("partial", c_void_p),
        ("work", c_void_p), ("unit_order", c_void_p),
        ("x2", c_void_p), ("ldx2", c_int64), ("n_split_src", c_int64),
        ("out2", c_void_p), ("ldo2", c_int64), ("n_split_out", c_int64), ("out2_push", c_void_p),
        ("col_hot", c_void_p),
        ("drop_p", c_float), ("drop_seed", ctypes.c_uint64), ("edge_id", c_void_p),
    ]


MAX_PEERS = 16


class HaloPushArgs(Structure):
    """Mirror of ``struct kgb_halo_push_args`` (include/kgb200.h)."""

    _fields_ = [("src", c_void_p), ("lds", c_int64), ("idx", c_void_p), ("F", c_int32), ("n_peers", c_int32),
                ("slot_begin", c_int64 * (MAX_PEERS + 1)), ("dst", c_void_p * MAX_PEERS),
                ("dst_row0", c_int64 * MAX_PEERS), ("ldd", c_int64), ("slot_rot", c_int64)]


class GatDropout(Structure):
    """Mirror of ``struct kgb_gat_dropout`` (include/kgb200.h)."""

    _fields_ = [("p", c_float), ("seed", ctypes.c_uint64), ("edge_id", c_void_p)]


class HubTable(Structure):
    """Mirror of ``struct kgb_hub_table`` (include/kgb200.h)."""

    _fields_ = [("hub_row", c_void_p), ("hub_chunk_base", c_void_p), ("hub_nchunks", c_void_p),
                ("chunk_hub", c_void_p), ("n_hubs", c_int32), ("n_chunks", c_int32), ("threshold", c_int32),
                ("chunk", c_int32), ("partial", c_void_p), ("work", c_void_p), ("unit_order", c_void_p)]


# name -> (restype, argtypes); every name declared in include/kgb200.h must appear here
SIGNATURES = {
    "kgb_version": (c_int, []),
    "kgb_last_error": (c_char_p, []),
    "kgb_launch_count": (c_int64, []),
    "kgb_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int64)]),
    "kgb_csr_build_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "kgb_csr_build": (c_int, [c_int, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kgb_csr_hubs": (c_int, [c_int, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int64, c_int64, c_void_p, c_void_p]),
    "kgb_gcn_norm": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "kgb_relu_bwd": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p]),
    "kgb_colsum_parts": (c_int32, [c_int, c_int64]),
    "kgb_relu_bwd_colsum": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int64,
                                    c_void_p, c_int32, c_void_p]),
    "kgb_softmax_xent_fwd": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "kgb_softmax_xent_bwd": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_float, c_void_p,
                                     c_int64, c_void_p]),
    "kgb_permute_f32": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "kgb_gather_reduce_partial_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "kgb_gather_reduce": (c_int, [c_int, POINTER(GatherReduceArgs), c_void_p]),
    "kgb_gather_max_bwd_acc_bytes": (c_size_t, [c_int64, c_int32]),
    "kgb_gather_max_bwd": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                   c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int64,
                                   c_int64, c_void_p, POINTER(HubTable), c_void_p]),
    "kgb_gather_unit_rows": (c_int32, []),
    "kgb_gatv2_unit_rows": (c_int32, []),
    "kgb_gather_max_bwd_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "kgb_dropout_mask": (c_int, [c_int, c_void_p, c_int64, c_int32, c_int32, c_float, ctypes.c_uint64, c_void_p,
                                 c_void_p]),
    "kgb_gather_rows": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_float, c_void_p, c_int64,
                                c_void_p]),
    "kgb_window_alloc": (c_int, [c_int, c_size_t, POINTER(c_void_p)]),
    "kgb_window_free": (c_int, [c_int, c_void_p]),
    "kgb_ipc_export": (c_int, [c_int, c_void_p, ctypes.c_char_p]),
    "kgb_ipc_open": (c_int, [c_int, ctypes.c_char_p, POINTER(c_void_p)]),
    "kgb_ipc_close": (c_int, [c_int, c_void_p]),
    "kgb_halo_push": (c_int, [c_int, POINTER(HaloPushArgs), c_void_p]),
    "kgb_gatv2_partial_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "kgb_gatv2_fwd": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_float,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(GatDropout),
                              POINTER(HubTable), c_void_p]),
    "kgb_gatv2_bwd_parts": (c_int, [c_int, c_int64, c_int32, c_int32]),
    "kgb_gatv2_bwd_dst": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                                  c_int32, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int32, c_void_p, POINTER(GatDropout),
                                  POINTER(HubTable), c_void_p]),
    "kgb_gatv2_rec_floats": (c_int32, [c_int32, c_int32]),
    "kgb_gatv2_bwd_src_rec": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_float, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(HubTable), c_void_p]),
    "kgb_gatv2_bwd_src": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                  c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, POINTER(GatDropout), POINTER(HubTable), c_void_p]),
    "kgb_reduce_parts": (c_int, [c_int, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "kgb_l2_normalize": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int32, c_float, c_void_p, c_int64, c_void_p,
                                 c_void_p]),
    "kgb_l2_normalize_bwd": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_float,
                                     c_void_p, c_int64, c_void_p]),
    "kgb_linear_tc_rows": (c_int32, [c_int32]),
    "kgb_split_tf32": (c_int, [c_int, c_void_p, c_int32, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "kgb_linear_tc": (c_int, [c_int, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                              c_int64, c_void_p, c_int32, c_void_p, c_int64, c_void_p]),
    "kgb_split_tf32_ld": (c_int, [c_int, c_void_p, c_int32, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_int64,
                                  c_void_p]),
    "kgb_linear_tc2_k": (c_int32, [c_int32, c_int32]),
    "kgb_linear_tc2": (c_int, [c_int, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                               c_int32, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_int64, c_void_p]),
    "kgb_linear_tc_dw_parts": (c_int32, [c_int, c_int64]),
    "kgb_linear_tc_dw": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p,
                                 c_int32, c_void_p]),
    "kgb_linear_tc_dw2_cols": (c_int32, [c_int32, c_int32]),
    "kgb_linear_tc_dw_x2": (c_int, [c_int, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_int64,
                                    c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "kgb_linear_tc_dw2": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_int32,
                                  c_int32, c_void_p, c_int32, c_void_p]),
}
ABI_VERSION = 204  # must equal kgb_version() of the loaded library (bumped with every ABI change)

_lock = threading.Lock()
_lib = None


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load libkgb200.so (building it with nvcc if it is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("KGB200_LIB") or _build.LIB  # KGB200_LIB: tuning builds of the same ABI
        if path == _build.LIB and build_if_missing and not _build.is_fresh():
            # the sources (or the header) changed since the library was linked: the ctypes mirrors below may no
            # longer match the binary, so a failed rebuild is an error even when an old .so is lying around
            try:
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    "keras_geometric_b200: libkgb200.so is missing or stale and could not be (re)compiled "
                    f"({e}). There is no CPU fallback for the message-passing hot path.") from e
        if not os.path.exists(path):
            raise RuntimeError("keras_geometric_b200: libkgb200.so not found; run "
                               "`python -m keras_geometric_b200._build`. There is no CPU fallback.")
        lib = ctypes.CDLL(path)
        lib.kgb_version.restype = c_int
        lib.kgb_version.argtypes = []
        got = int(lib.kgb_version())
        if got != ABI_VERSION:
            raise RuntimeError(f"keras_geometric_b200: {path} exports ABI version {got}, this package binds "
                               f"version {ABI_VERSION}; rebuild with `python -m keras_geometric_b200._build --force`")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().kgb_last_error()
        raise KgbError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
