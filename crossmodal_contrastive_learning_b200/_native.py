"""ctypes binding of libcrossclr_b200.so (the C ABI declared in include/crossclr_b200.h).

There is deliberately no fallback here: if the shared library cannot be loaded or built, importing the
criterion on a GPU box fails loudly (`NativeLibraryError`).  Nothing in this package computes the loss
on the CPU or through PyTorch ops.
"""
from __future__ import annotations

import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcrossclr_b200.so")

# element types / paths (mirror include/crossclr_b200.h)
F32, F16, BF16, F16X2 = 0, 1, 2, 3      # F16X2: fp16 hi + lo rows of PATH_TC_SPLIT (stacked matrix only)
PATH_AUTO, PATH_SIMT, PATH_TC, PATH_TC_SPLIT = 0, 1, 2, 3
ROW_TAIL = 64      # CROSSCLR_ROW_TAIL: extra elements per stacked row on the tensor-core paths (residual scale)

EXPORTS = (
    "crossclr_version", "crossclr_last_error", "crossclr_device_supported", "crossclr_choose_path",
    "crossclr_feature_dtype", "crossclr_workspace_bytes", "crossclr_pack", "crossclr_fwd",
    "crossclr_finalize", "crossclr_bwd", "crossclr_shift", "crossclr_launch_count", "crossclr_selftest",
    "crossclr_timing_enable", "crossclr_timing_read", "crossclr_pack2", "crossclr_forward",
    "crossclr_maxmargin_workspace_bytes", "crossclr_maxmargin_fwd", "crossclr_maxmargin_bwd",
    "crossclr_maxmargin_kernel_name", "crossclr_retrieval_ranks",
    "crossclr_peer_alloc", "crossclr_peer_free", "crossclr_peer_export", "crossclr_peer_import", "crossclr_peer_release",
    "crossclr_peer_exchange",
    "crossclr_bwd_kernel_name", "crossclr_feature_pitch", "crossclr_segment_rows", "crossclr_bwd_accumulate", "crossclr_bwd_finish", "crossclr_bwd_scale_grad",
)
KERNEL_FAMILIES = ("pack", "fwd", "finalize", "bwd", "grad_finish")


class NativeLibraryError(RuntimeError):
    pass


class Problem(ctypes.Structure):
    """crossclr_problem_t"""
    _fields_ = [
        ("nseg", ctypes.c_int32), ("bseg", ctypes.c_int32), ("dim", ctypes.c_int32),
        ("row_begin", ctypes.c_int32), ("row_count", ctypes.c_int32),
        ("temperature", ctypes.c_float), ("negative_weight", ctypes.c_float),
    ]


_lock = threading.Lock()
_lib = None


def _declare(lib):
    c = ctypes
    P = c.POINTER(Problem)
    vp = c.c_void_p
    lib.crossclr_version.restype = c.c_int
    lib.crossclr_version.argtypes = []
    lib.crossclr_last_error.restype = c.c_char_p
    lib.crossclr_last_error.argtypes = []
    lib.crossclr_device_supported.restype = c.c_int
    lib.crossclr_device_supported.argtypes = [c.c_int]
    lib.crossclr_choose_path.restype = c.c_int
    lib.crossclr_choose_path.argtypes = [P, c.c_int, c.c_int]
    lib.crossclr_feature_dtype.restype = c.c_int
    lib.crossclr_feature_dtype.argtypes = [c.c_int]
    lib.crossclr_feature_pitch.restype = c.c_int64
    lib.crossclr_feature_pitch.argtypes = [c.c_int, c.c_int32]
    lib.crossclr_segment_rows.restype = c.c_int64
    lib.crossclr_segment_rows.argtypes = [c.c_int, c.c_int32]
    lib.crossclr_workspace_bytes.restype = c.c_size_t
    lib.crossclr_workspace_bytes.argtypes = [P, c.c_int]
    lib.crossclr_pack.restype = c.c_int
    lib.crossclr_pack.argtypes = [vp, c.c_int, c.c_int64, c.c_int32, c.c_int32, vp, c.c_int, vp, vp]
    lib.crossclr_pack2.restype = c.c_int
    lib.crossclr_pack2.argtypes = [vp, vp, c.c_int, c.c_int64, c.c_int64, c.c_int32, c.c_int32, vp, c.c_int, vp, vp]
    lib.crossclr_forward.restype = c.c_int
    lib.crossclr_forward.argtypes = [P, c.c_int, vp, vp, c.c_int, c.c_int64, c.c_int64, vp, vp, vp, vp, vp, vp, vp]
    lib.crossclr_fwd.restype = c.c_int
    lib.crossclr_fwd.argtypes = [P, c.c_int, vp, vp, vp, c.c_size_t, vp]
    lib.crossclr_finalize.restype = c.c_int
    lib.crossclr_finalize.argtypes = [P, c.c_int, vp, vp, vp, vp, vp]
    lib.crossclr_bwd.restype = c.c_int
    lib.crossclr_bwd.argtypes = [P, c.c_int, vp, vp, vp, vp, vp, c.c_float, vp, c.c_int64, vp, c.c_int64,
                                 c.c_int, vp, c.c_size_t, vp]
    lib.crossclr_bwd_accumulate.restype = c.c_int
    lib.crossclr_bwd_accumulate.argtypes = [P, c.c_int, vp, vp, vp, vp, c.c_size_t, vp]
    lib.crossclr_bwd_finish.restype = c.c_int
    lib.crossclr_bwd_finish.argtypes = [P, c.c_int, vp, vp, vp, vp, vp, c.c_float, vp, c.c_int64, vp, c.c_int64, c.c_int, vp, vp]
    lib.crossclr_bwd_scale_grad.restype = c.c_int
    lib.crossclr_bwd_scale_grad.argtypes = [P, c.c_int, vp, vp, vp, vp, c.c_float, c.c_float, vp, vp, vp]
    lib.crossclr_shift.restype = c.c_float
    lib.crossclr_shift.argtypes = [P]
    lib.crossclr_bwd_kernel_name.restype = c.c_char_p
    lib.crossclr_bwd_kernel_name.argtypes = [P, c.c_int]
    lib.crossclr_launch_count.restype = c.c_int64
    lib.crossclr_launch_count.argtypes = []
    lib.crossclr_timing_enable.restype = c.c_int
    lib.crossclr_timing_enable.argtypes = [c.c_int]
    lib.crossclr_timing_read.restype = c.c_int
    lib.crossclr_timing_read.argtypes = [c.c_int, c.POINTER(c.c_double), c.POINTER(c.c_int64)]
    lib.crossclr_selftest.restype = c.c_int
    lib.crossclr_selftest.argtypes = [c.c_int, vp, vp, vp, c.c_int32, c.c_int32]
    lib.crossclr_maxmargin_workspace_bytes.restype = c.c_size_t
    lib.crossclr_maxmargin_workspace_bytes.argtypes = [c.c_int32, c.c_int32, c.c_int]
    lib.crossclr_maxmargin_kernel_name.restype = c.c_char_p
    lib.crossclr_maxmargin_kernel_name.argtypes = [vp, vp, c.c_int, c.c_int64, c.c_int64, c.c_int32, c.c_int32]
    lib.crossclr_peer_alloc.restype = c.c_int
    lib.crossclr_peer_alloc.argtypes = [c.c_size_t, c.POINTER(c.c_void_p)]
    lib.crossclr_peer_free.restype = c.c_int
    lib.crossclr_peer_free.argtypes = [vp]
    lib.crossclr_peer_export.restype = c.c_int
    lib.crossclr_peer_export.argtypes = [vp, vp]
    lib.crossclr_peer_import.restype = c.c_int
    lib.crossclr_peer_import.argtypes = [vp, c.POINTER(c.c_void_p)]
    lib.crossclr_peer_release.restype = c.c_int
    lib.crossclr_peer_release.argtypes = [vp]
    lib.crossclr_peer_exchange.restype = c.c_int
    lib.crossclr_peer_exchange.argtypes = [vp, vp, c.c_int32, c.c_int32, c.c_size_t, c.c_size_t, c.c_int, vp, vp]
    lib.crossclr_retrieval_ranks.restype = c.c_int
    lib.crossclr_retrieval_ranks.argtypes = [vp, vp, c.c_int, c.c_int64, c.c_int64, c.c_int32, c.c_int32, vp, c.c_size_t,
                                             vp, vp, vp]
    lib.crossclr_maxmargin_fwd.restype = c.c_int
    lib.crossclr_maxmargin_fwd.argtypes = [vp, vp, c.c_int, c.c_int64, c.c_int64, c.c_int32, c.c_int32, c.c_float, vp,
                                           c.c_size_t, vp, vp]
    lib.crossclr_maxmargin_bwd.restype = c.c_int
    lib.crossclr_maxmargin_bwd.argtypes = [vp, vp, c.c_int, c.c_int64, c.c_int64, c.c_int32, c.c_int32, c.c_float, vp,
                                           c.c_size_t, vp, vp, c.c_int64, vp, c.c_int64, c.c_int, vp]


def load(build_if_missing: bool = True):
    """Load (building first if the .so is missing/stale and nvcc is present) and return the CDLL."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing:
            try:
                from . import build as _build
                if _build.is_stale():
                    _build.build()
            except Exception as e:  # no nvcc on this box: fine if a prebuilt library travelled with the tree
                if not os.path.exists(LIB_PATH):
                    raise NativeLibraryError(
                        f"libcrossclr_b200.so is missing and could not be built: {e}") from e
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(f"{LIB_PATH} not found; run `python -m crossmodal_contrastive_learning_b200.build`")
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        missing = [s for s in EXPORTS if not hasattr(lib, s)]
        if missing:
            raise NativeLibraryError(f"{LIB_PATH} lacks symbols {missing}")
        _declare(lib)
        _lib = lib
        return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().crossclr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().crossclr_launch_count())


def timing_enable(on: bool):
    check(load().crossclr_timing_enable(1 if on else 0), "crossclr_timing_enable")


def timing_read():
    """{family: (total_ms, launches)} accumulated since the last read (synchronises the recorded events)."""
    lib = load()
    out = {}
    for k, name in enumerate(KERNEL_FAMILIES):
        ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
        check(lib.crossclr_timing_read(k, ctypes.byref(ms), ctypes.byref(n)), "crossclr_timing_read")
        out[name] = (ms.value, n.value)
    return out
