"""ctypes binding of libfast_rnnt_b200.so (C ABI: include/fast_rnnt_b200.h).

This is the TensorFlow-free twin of the ``tf.load_op_library`` call of the
reference (tf_fast_rnnt/python/tf_fast_rnnt/__init__.py:38-40).  There is no CPU
fallback: if the shared library is missing or does not load, importing the
package fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_size_t, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "FAST_RNNT_B200_LIB",
    os.path.join(os.path.dirname(_PKG_DIR), "lib", "libfast_rnnt_b200.so"))

FRN_OK = 0
REGULAR, MODIFIED, CONSTRAINED = 0, 1, 2
F32, BF16, F16 = 0, 1, 2
EUNSUPPORTED = -4
NONE, MEAN, SUM = 0, 1, 2
RNNT_TYPES = {"regular": REGULAR, "modified": MODIFIED, "constrained": CONSTRAINED}
REDUCTIONS = {"none": NONE, "mean": MEAN, "sum": SUM}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"fast_rnnt_b200: {LIB_PATH} not found. Build it with "
        "`make -C tf-fast-rnnt_b200/csrc` (or __graft_entry__.build()); "
        "there is no CPU fallback.")


class _Library:
    """The loaded C-ABI library.  Normally the product build; `use_debug_hooks(True)` swaps in the
    -DFRN_DEBUG_HOOKS build (libfast_rnnt_b200_dbg.so: same sources, plus the FRN_* environment overrides
    that force one of two implementations of a stage) for the tests that cross-check kernels against
    each other.  Attribute access goes to whichever build is current."""

    def __init__(self, path):
        self._product = ctypes.CDLL(path)
        self._debug = None
        self._cur = self._product

    def __getattr__(self, name):
        return getattr(self._cur, name)


lib = _Library(LIB_PATH)
DEBUG_LIB_PATH = LIB_PATH[:-3] + "_dbg.so"


def _bind(cdll):
    for name, (res, args) in _SIGS.items():
        fn = getattr(cdll, name)      # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args


def use_debug_hooks(on: bool) -> None:
    if on and lib._debug is None:
        if not os.path.exists(DEBUG_LIB_PATH):
            raise ImportError(f"fast_rnnt_b200: {DEBUG_LIB_PATH} not found (make -C tf-fast-rnnt_b200/csrc)")
        lib._debug = ctypes.CDLL(DEBUG_LIB_PATH)
        _bind(lib._debug)
    lib._cur = lib._debug if on else lib._product


_P = c_void_p
_SIGS = {
    "frn_version": (c_int, []),
    "frn_status_string": (ctypes.c_char_p, [c_int]),
    "frn_last_cuda_error": (c_int, []),
    "frn_kernel_launches": (ctypes.c_ulonglong, []),
    "frn_mi_workspace_bytes": (c_size_t, [c_int] * 4),
    "frn_mi_fwd_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "frn_band_mi_workspace_bytes": (c_size_t, [c_int] * 4),
    "frn_band_mi_fwd_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P,
                                    c_size_t, _P]),
    "frn_cummin": (c_int, [_P, _P, c_int, c_int, _P]),
    "frn_simple_logprobs_workspace_bytes": (c_size_t, [c_int] * 4),
    "frn_simple_logprobs": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_float, c_float, _P, _P, _P, c_size_t, _P]),
    "frn_simple_loss_workspace_bytes": (c_size_t, [c_int] * 4),
    "frn_simple_loss": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_float, c_float, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "frn_simple_loss_bcast": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_float, c_float, c_float, c_int, _P, _P, _P, c_int, _P, c_int, _P, _P, _P, _P,
                                      c_size_t, _P]),
    "frn_simple_loss_bwd_workspace_bytes": (c_size_t, [c_int] * 4),
    "frn_simple_loss_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int,
                                    _P, _P, _P, c_size_t, _P]),
    "frn_smoothed_loss_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_float, c_float, _P, _P, _P, c_size_t, _P]),
    "frn_cast_to_f32": (c_int, [_P, c_int, c_size_t, _P, _P]),
    "frn_reduce_pair": (c_int, [_P, _P, c_int, c_int, c_float, _P, _P, _P]),
    "frn_prune_ranges_width": (c_int, [c_int, c_int]),
    "frn_prune_ranges_workspace_bytes": (c_size_t, [c_int, c_int]),
    "frn_prune_ranges": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "frn_do_pruning": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "frn_do_pruning_lp": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "frn_do_pruning_add_joiner": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "frn_broadcast_am_pruned": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "frn_do_pruning_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "frn_add_joiner": (c_int, [_P, _P, _P, c_size_t, _P]),
    "frn_pruned_add_joiner": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "frn_pruned_logprobs_workspace_bytes": (c_size_t, [c_int] * 4),
    "frn_pruned_logprobs": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    _P, _P, _P, c_size_t, _P]),
    "frn_smoothed_unigram_sums": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "frn_simple_logprobs_sharded": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                            c_float, _P, _P, _P, _P, c_size_t, _P]),
    "frn_simple_loss_sharded": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                        c_float, _P, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "frn_simple_logprobs_lp": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                       c_float, _P, _P, _P, _P, c_size_t, _P]),
    "frn_simple_loss_lp": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                   c_float, _P, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "frn_smoothed_loss_bwd_sharded": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int,
                                              c_float, c_float, _P, _P, c_int, _P, _P, _P, c_size_t, _P]),
    "frn_allreduce_sum": (c_int, [_P, c_size_t, _P, _P]),
    "frn_pruned_logprobs_bwd": (c_int, [_P, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        _P, _P, c_size_t, _P]),
    "frn_pruned_loss_workspace_bytes": (c_size_t, [c_int] * 4),
    "frn_pruned_loss_min_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_float]),
    "frn_pruned_loss": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_float, _P, _P, _P, _P, c_size_t, _P]),
    "frn_joint_loss_workspace_bytes": (c_size_t, [c_int] * 3),
    "frn_joint_loss": (c_int, [_P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                               _P, _P, _P, _P, c_size_t, _P]),
    "frn_reduce": (c_int, [_P, c_int, c_int, c_float, _P, _P]),
}
EXPORTS = tuple(_SIGS)
_bind(lib._product)


class FastRnntError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status != FRN_OK:
        msg = lib.frn_status_string(status).decode()
        extra = f" (cudaError {lib.frn_last_cuda_error()})" if status == -3 else ""
        raise FastRnntError(f"{what}: {msg}{extra}")
