"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy + the plain-C recursion in ``mi_recursion.c``) of the
pruned RNN-T loss hot path of Samsung/tf-fast-rnnt.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product (``tf-fast-rnnt_b200/``) never does.

Every function cites the reference lines it restates (paths are under
/root/reference/tf_fast_rnnt/python/tf_fast_rnnt/ unless stated otherwise).
``dtype=np.float32`` follows the reference's arithmetic type; ``np.float64``
is the "truth" mode the float tolerances are measured against.

Parity status.  The reference's tests contain no expected values, so they do
not pin anything ("parity unpinned" by the reference's own tests).  What pins
this oracle instead:
  * tests/golden/*.npz — outputs of the reference's *own* ``rnnt_loss.py``
    executed here, unmodified, on top of a NumPy stand-in for the TensorFlow ops
    it calls (``oracle/tf_emu``; generator ``tests/golden/make_golden.py``);
  * the reference's CUDA kernels, compiled in place into ``oracle/_ref`` and run
    on the B200 (lattice recursion fwd/bwd and cummin);
  * ``torchaudio.functional.rnnt_loss`` on CPU (regular type, no penalty).
Where the reference is shape-broken (``modified``/``constrained`` in the
simple/smoothed log-probs) the documented semantics are
used and said so at the spot.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
TINY = float(np.nextafter(np.float32(0.0), np.float32(1.0)))  # rnnt_loss.py:181


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "liborc.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", _HERE, "lib"])
        _LIB = ctypes.CDLL(so)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _suffix(dtype):
    return "f64" if np.dtype(dtype) == np.float64 else "f32"


# --------------------------------------------------------------------------
# A4  lattice recursion  (__init__.py:42-149, mutual_information_cuda.cu)
# --------------------------------------------------------------------------
def mutual_information_recursion(px, py, boundary, calc_gradients=False,
                                 dtype=np.float32, return_p=False):
    """ans[b] = p[b, s_end, t_end] of the log-add recursion; with
    ``calc_gradients`` also the occupation counts (px_grad, py_grad) obtained
    with ans_grad = 1 (tf_fast_rnnt_op.cc:100-110).  ``px_grad`` has the shape
    of ``px`` (the reference allocates [B,S,T+1] even for the modified
    recursion, tf_fast_rnnt_op.cc:84 — defect D2)."""
    px = np.ascontiguousarray(px, dtype=dtype)
    py = np.ascontiguousarray(py, dtype=dtype)
    boundary = np.ascontiguousarray(boundary, dtype=np.int32)
    B, S, T1 = px.shape
    T = py.shape[2]
    assert py.shape == (B, S + 1, T) and T1 in (T, T + 1)
    assert boundary.shape == (B, 4)
    suf = _suffix(dtype)
    lib = _lib()
    p = np.full((B, S + 1, T + 1), -np.inf, dtype=dtype)
    ans = np.zeros((B,), dtype=dtype)
    getattr(lib, f"orc_mi_forward_{suf}")(
        _ptr(px), _ptr(py), _ptr(boundary), B, S, T, T1, _ptr(p), _ptr(ans))
    if not calc_gradients:
        return (ans, p) if return_p else ans
    p_grad = np.zeros_like(p)
    px_grad = np.zeros_like(px)
    py_grad = np.zeros_like(py)
    ans_grad = np.ones((B,), dtype=dtype)
    getattr(lib, f"orc_mi_backward_{suf}")(
        _ptr(px), _ptr(py), _ptr(boundary), _ptr(p), _ptr(ans_grad),
        B, S, T, T1, _ptr(p_grad), _ptr(px_grad), _ptr(py_grad))
    if return_p:
        return ans, (px_grad, py_grad), p, p_grad
    return ans, (px_grad, py_grad)


def cummin(x):
    """Inclusive running minimum along the last axis of an int32 matrix
    (op Cummin, tf_fast_rnnt_op.cc:36-38; kernel cu:895-1012)."""
    x = np.ascontiguousarray(x, dtype=np.int32)
    out = np.empty_like(x)
    _lib().orc_cummin_i32(_ptr(x), _ptr(out), x.shape[0], x.shape[1])
    return out


# --------------------------------------------------------------------------
# A1 / A2  simple and smoothed log-probs (rnnt_loss.py:63-223, 1132-1367)
# --------------------------------------------------------------------------
def _neg_inf_at_t_end(px, boundary):
    """fix_for_boundary, rnnt_loss.py:28-61: px[b, :, boundary[b,3]] = -inf."""
    if boundary is None:
        return px
    B = px.shape[0]
    px = px.copy()
    px[np.arange(B), :, np.asarray(boundary)[:, 3]] = -np.inf
    return px


def _normalizers(lm, am, dtype):
    """rnnt_loss.py:175-186.  Returns (normalizers[B,S+1,T], lm_probs, am_probs,
    lm_max, am_max)."""
    am_max = am.max(axis=2, keepdims=True)
    lm_max = lm.max(axis=2, keepdims=True)
    am_probs = np.exp(am - am_max)
    lm_probs = np.exp(lm - lm_max)
    prod = np.matmul(lm_probs, np.swapaxes(am_probs, 1, 2))
    norm = np.log(prod + dtype(TINY))
    norm = norm + lm_max + np.swapaxes(am_max, 1, 2)
    return norm.astype(dtype), lm_probs, am_probs, lm_max, am_max


def get_rnnt_logprobs(lm, am, symbols, termination_symbol, rnnt_type="regular",
                      boundary=None, dtype=np.float32):
    """rnnt_loss.py:63-223.  For ``modified``/``constrained`` the reference
    subtracts a [B,S,T+1] normaliser from a [B,S,T] px (line 211, defect D1);
    the documented semantics — subtract over the T real columns — is used."""
    assert rnnt_type in ("regular", "modified", "constrained")
    dtype = np.dtype(dtype).type
    lm = np.asarray(lm, dtype=dtype)
    am = np.asarray(am, dtype=dtype)
    symbols = np.asarray(symbols)
    B, T, C = am.shape
    S = lm.shape[1] - 1
    norm, *_ = _normalizers(lm, am, dtype)
    bi = np.arange(B)[:, None]
    # px_am[b,s,t] = am[b,t,symbols[b,s]]            (:187-192)
    px_am = np.swapaxes(am, 1, 2)[bi, symbols, :]          # [B,S,T]
    px_lm = lm[bi, np.arange(S)[None, :], symbols][:, :, None]  # [B,S,1]
    px_core = px_am + px_lm - norm[:, :S, :]
    if rnnt_type == "regular":
        px = np.concatenate(
            [px_core, np.full((B, S, 1), -np.inf, dtype=dtype)], axis=2)
    else:
        px = px_core
    py = (am[:, :, termination_symbol][:, None, :]
          + lm[:, :, termination_symbol][:, :, None] - norm)    # (:214-216)
    if rnnt_type == "regular":
        px = _neg_inf_at_t_end(px, boundary)
    elif rnnt_type == "constrained":
        px = px + py[:, 1:, :]
    return px.astype(dtype), py.astype(dtype)


def smoothed_unigram_sums(lm, dtype=np.float64):
    """One shard's share of the batch-global unigram of rnnt_loss.py:1279-1280: the column sums of
    softmax(lm rows) [C] and the row count.  Summed over the shards of a batch (an all-reduce on the GPUs) and
    handed to get_rnnt_logprobs_smoothed / rnnt_loss_smoothed as `unigram_sums`, every shard sees the unigram of
    the whole batch."""
    dtype = np.dtype(dtype).type
    lm = np.asarray(lm, dtype=dtype)
    probs = np.exp(lm - lm.max(axis=2, keepdims=True))
    r = probs / probs.sum(axis=2, keepdims=True)
    return r.sum(axis=(0, 1)), float(lm.shape[0] * lm.shape[1])


def get_rnnt_logprobs_smoothed(lm, am, symbols, termination_symbol,
                               lm_only_scale=0.1, am_only_scale=0.1,
                               boundary=None, rnnt_type="regular",
                               dtype=np.float32, unigram_sums=None):
    """rnnt_loss.py:1132-1367 (same D1 decision as get_rnnt_logprobs)."""
    assert rnnt_type in ("regular", "modified", "constrained")
    dtype = np.dtype(dtype).type
    lm = np.asarray(lm, dtype=dtype)
    am = np.asarray(am, dtype=dtype)
    symbols = np.asarray(symbols)
    B, T, C = am.shape
    S = lm.shape[1] - 1
    norm, lm_probs, am_probs, lm_max, am_max = _normalizers(lm, am, dtype)
    lmonly = lm_probs.sum(axis=2, keepdims=True)                       # :1276
    unigram = (lm_probs / lmonly).mean(axis=(0, 1), keepdims=True,
                                       dtype=dtype) + dtype(TINY)      # :1279
    if unigram_sums is not None:                                       # sharded batch: the global mean
        sums, count = unigram_sums
        unigram = (np.asarray(sums, dtype=dtype) / dtype(count)).reshape(1, 1, C) + dtype(TINY)
    amonly = np.log(am_probs.reshape(-1, C) @ unigram.reshape(C)
                    ).reshape(B, T, 1) + am_max                        # :1281
    amonly = np.swapaxes(amonly, 1, 2).astype(dtype)                   # [B,1,T]
    log_unigram = np.log(unigram).astype(dtype)
    lmonly = (np.log(lmonly) + lm_max).astype(dtype)                   # [B,S+1,1]

    bi = np.arange(B)[:, None]
    px_am = np.swapaxes(am, 1, 2)[bi, symbols, :]                      # [B,S,T]
    px_lm = lm[bi, np.arange(S)[None, :], symbols][:, :, None]
    px_lm_unigram = log_unigram.reshape(-1)[symbols][:, :, None]       # :1319
    px = px_am + px_lm - norm[:, :S, :]
    px_amonly = px_am + px_lm_unigram - amonly
    px_lmonly = px_lm - lmonly[:, :S, :]

    py_am = am[:, :, termination_symbol][:, None, :]
    py_lm = lm[:, :, termination_symbol][:, :, None]
    py = py_am + py_lm - norm
    py_amonly = py_am + log_unigram[0, 0, termination_symbol] - amonly
    py_lmonly = py_lm - lmonly

    combined = dtype(1.0 - lm_only_scale - am_only_scale)
    lms = dtype(1.0e-20 if lm_only_scale == 0.0 else lm_only_scale)    # :1346
    ams = dtype(1.0e-20 if am_only_scale == 0.0 else am_only_scale)
    px_i = px * combined + px_lmonly * lms + px_amonly * ams
    py_i = py * combined + py_lmonly * lms + py_amonly * ams
    if rnnt_type == "regular":
        px_i = np.concatenate(
            [px_i, np.full((B, S, 1), -np.inf, dtype=dtype)], axis=2)
        px_i = _neg_inf_at_t_end(px_i, boundary)
    elif rnnt_type == "constrained":
        px_i = px_i + py_i[:, 1:, :]
    return px_i.astype(dtype), py_i.astype(dtype)


# --------------------------------------------------------------------------
# A3  delay penalty + reduction (rnnt_loss.py:305-338)
# --------------------------------------------------------------------------
def apply_delay_penalty(px, boundary, delay_penalty):
    """px[b,s,t] += f32(((boundary[b,3]-1)/2 - t) * delay_penalty), computed in
    float64 and cast (int32 true division gives float64 in TF), :316-321."""
    if not delay_penalty > 0.0:
        return px
    B, S, T0 = px.shape
    offset = (np.asarray(boundary)[:, 3].astype(np.int64) - 1) / 2    # float64
    penalty = offset.reshape(B, 1, 1) - np.arange(T0, dtype=np.float64).reshape(1, 1, T0)
    penalty = penalty * delay_penalty
    return px + penalty.astype(px.dtype)


def _reduce(scores, reduction):
    if reduction == "none":
        return -scores
    if reduction == "mean":      # rnnt_loss.py:331 says torch.mean (defect D3)
        return -scores.mean(dtype=scores.dtype)
    if reduction == "sum":
        return -scores.sum(dtype=scores.dtype)
    raise ValueError(
        f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")


def _loss_from_logprobs(px, py, boundary, delay_penalty, reduction,
                        calc_gradients, dtype):
    px = apply_delay_penalty(px, boundary, delay_penalty)
    out = mutual_information_recursion(px, py, boundary, calc_gradients, dtype)
    scores = out[0] if calc_gradients else out
    loss = _reduce(scores, reduction)
    return (loss, out[1]) if calc_gradients else loss


def rnnt_loss_simple(lm, am, symbols, termination_symbol, boundary,
                     rnnt_type="regular", delay_penalty=0.0, reduction="mean",
                     calc_gradients=False, dtype=np.float32):
    """rnnt_loss.py:225-338."""
    px, py = get_rnnt_logprobs(lm, am, symbols, termination_symbol, rnnt_type,
                               boundary, dtype)
    return _loss_from_logprobs(px, py, boundary, delay_penalty, reduction,
                               calc_gradients, dtype)


def rnnt_loss_smoothed(lm, am, symbols, termination_symbol, lm_only_scale=0.1,
                       am_only_scale=0.1, boundary=None, rnnt_type="regular",
                       delay_penalty=0.0, reduction="mean",
                       calc_gradients=False, dtype=np.float32, unigram_sums=None):
    """rnnt_loss.py:1369-1494."""
    px, py = get_rnnt_logprobs_smoothed(lm, am, symbols, termination_symbol,
                                        lm_only_scale, am_only_scale, boundary,
                                        rnnt_type, dtype, unigram_sums)
    return _loss_from_logprobs(px, py, boundary, delay_penalty, reduction,
                               calc_gradients, dtype)


# --------------------------------------------------------------------------
# A5  prune ranges (rnnt_loss.py:553-761)
# --------------------------------------------------------------------------
def _rev_cummin(x):
    """_monotonic_lower_bound, rnnt_loss.py:582-585."""
    return cummin(x[:, ::-1])[:, ::-1]


def _adjust_pruning_lower_bound(s_begin, r):
    """rnnt_loss.py:623-641, int32 throughout."""
    T = s_begin.shape[1]
    ramp = (np.int32(r - 1) * np.arange(T, dtype=np.int32))[None, :]
    x = _rev_cummin(s_begin)
    x = -(x - ramp)
    x = _rev_cummin(x)
    x = np.maximum(x, 0)
    x = -(x - ramp)
    return x.astype(np.int32)


def get_rnnt_prune_ranges(px_grad, py_grad, boundary, s_range):
    """rnnt_loss.py:647-761.  Float order is *defined* here:
    sequential float32 cumsum over s, then cs[k+R]-cs[k], then the px_grad
    subtraction, then first-index argmax."""
    px_grad = np.asarray(px_grad, dtype=np.float32)
    py_grad = np.asarray(py_grad, dtype=np.float32)
    boundary = np.asarray(boundary, dtype=np.int32)
    B, S, T1 = px_grad.shape
    T = py_grad.shape[2]
    S1 = S + 1
    R = S + 1 if s_range > S else s_range                              # :710
    cs = np.zeros((B, S1 + 1, T), dtype=np.float32)
    np.cumsum(py_grad, axis=1, dtype=np.float32, out=cs[:, 1:, :])     # :722
    blk = cs[:, R:, :] - cs[:, :S1 - R + 1, :]                         # :725
    pxp = np.zeros((B, S1, T), dtype=np.float32)
    pxp[:, 1:, :] = px_grad[:, :, :T]
    fin = blk - pxp[:, :S1 - R + 1, :]                                 # :728
    s_begin = np.argmax(fin, axis=1).astype(np.int32)                  # :729
    mask = np.arange(T)[None, :] < (boundary[:, 3][:, None] - 1)       # :741
    pad = np.maximum(boundary[:, 2][:, None] - R + 1, 0)               # :744
    s_begin = np.where(mask, s_begin, pad).astype(np.int32)
    s_begin = _adjust_pruning_lower_bound(s_begin, 2 if T1 == T else R)  # :756
    ranges = s_begin[:, :, None] + np.arange(R, dtype=np.int32)[None, None, :]
    return ranges.astype(np.int32)


# --------------------------------------------------------------------------
# A6  pruning gather (rnnt_loss.py:763-812) and its gradient
# --------------------------------------------------------------------------
def do_rnnt_pruning(am, lm, ranges):
    am = np.asarray(am)
    lm = np.asarray(lm)
    B, T, R = ranges.shape
    C = am.shape[2]
    am_p = np.broadcast_to(am[:, :, None, :], (B, T, R, C)).copy()
    lm_p = lm[np.arange(B)[:, None, None], ranges, :]
    return am_p, lm_p


def do_rnnt_pruning_bwd(am_p_grad, lm_p_grad, ranges, S1):
    """Gradient TensorFlow derives for broadcast_to + gather (:802-811):
    am_grad = sum over the s_range axis, lm_grad = scatter-add over ranges."""
    B, T, R, C = am_p_grad.shape
    am_grad = am_p_grad.sum(axis=2, dtype=np.float64)
    lm_grad = np.zeros((B, S1, C), dtype=np.float64)
    for b in range(B):
        np.add.at(lm_grad[b], ranges[b].reshape(-1),
                  lm_p_grad[b].reshape(T * R, C).astype(np.float64))
    return am_grad, lm_grad


# --------------------------------------------------------------------------
# A7  pruned log-probs (rnnt_loss.py:853-1020) / joint (340-452)
# --------------------------------------------------------------------------
def _logsumexp(x, axis):
    m = x.max(axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0)
    return (np.log(np.exp(x - m).sum(axis=axis, keepdims=True, dtype=x.dtype)) + m
            ).squeeze(axis)


def pruned_compact_logprobs(logits, symbols, ranges, termination_symbol,
                            dtype=np.float32):
    """Band-local log-probs: pxc/pyc[b,t,i] for the arc leaving (ranges[b,t,i], t)
    (rnnt_loss.py:942-965, 995-996).  logits may be any float type; they are
    upcast to ``dtype`` first (bf16 inputs: 'upcast then reference math')."""
    dtype = np.dtype(dtype).type
    logits = np.asarray(logits).astype(dtype)
    B, T, R, C = logits.shape
    S = symbols.shape[1]
    norm = _logsumexp(logits, 3)
    sym_t = np.concatenate(
        [np.asarray(symbols), np.full((B, 1), termination_symbol, dtype=symbols.dtype)], axis=1)
    psym = sym_t[np.arange(B)[:, None, None], ranges]                  # [B,T,R]
    pxc = np.take_along_axis(logits, psym[..., None].astype(np.int64), axis=3)[..., 0] - norm
    pyc = logits[:, :, :, termination_symbol] - norm
    return pxc.astype(dtype), pyc.astype(dtype)


def get_rnnt_logprobs_pruned(logits, symbols, ranges, termination_symbol,
                             boundary, rnnt_type="regular", dtype=np.float32):
    """Dense [B,S,T1]/[B,S+1,T] lattice from the band (pad + roll + transpose,
    rnnt_loss.py:968-1018), written as index arithmetic:
    px[b,s,t] = pxc[b,t,s-ranges[b,t,0]] if that index is in [0,R) and s<S."""
    dtype = np.dtype(dtype).type
    pxc, pyc = pruned_compact_logprobs(logits, symbols, ranges,
                                       termination_symbol, dtype)
    B, T, R = pxc.shape
    S = symbols.shape[1]
    r0 = np.asarray(ranges)[:, :, 0]
    s_idx = np.arange(S + 1)[None, None, :]
    i = (s_idx - r0[:, :, None]) % (S + 1)         # roll index, :849
    inb = i < R
    ic = np.minimum(i, R - 1)
    px_ts = np.where(inb, np.take_along_axis(pxc, ic, axis=2), -np.inf)[:, :, :S]
    py_ts = np.where(inb, np.take_along_axis(pyc, ic, axis=2), -np.inf)
    px = np.swapaxes(px_ts, 1, 2).astype(dtype)
    py = np.swapaxes(py_ts, 1, 2).astype(dtype)
    if rnnt_type == "regular":
        px = np.concatenate([px, np.full((B, S, 1), -np.inf, dtype=dtype)], axis=2)
        px = _neg_inf_at_t_end(px, boundary)
    elif rnnt_type == "constrained":
        px = px + py[:, 1:, :]
    return px, py


def get_rnnt_logprobs_joint(logits, symbols, termination_symbol, boundary=None,
                            rnnt_type="regular", dtype=np.float32):
    """rnnt_loss.py:340-452 (D1 decision for the non-regular types)."""
    dtype = np.dtype(dtype).type
    logits = np.asarray(logits).astype(dtype)
    B, T, S1, C = logits.shape
    S = S1 - 1
    norm = np.swapaxes(_logsumexp(logits, 3), 1, 2)                     # [B,S+1,T]
    idx = np.broadcast_to(np.asarray(symbols)[:, None, :, None], (B, T, S, 1))
    px = np.take_along_axis(logits[:, :, :S, :], idx.astype(np.int64), axis=3)[..., 0]
    px = np.swapaxes(px, 1, 2) - norm[:, :S, :]
    if rnnt_type == "regular":
        px = np.concatenate([px, np.full((B, S, 1), -np.inf, dtype=dtype)], axis=2)
    py = np.swapaxes(logits[:, :, :, termination_symbol], 1, 2) - norm
    if rnnt_type == "regular":
        px = _neg_inf_at_t_end(px, boundary)
    elif rnnt_type == "constrained":
        px = px + py[:, 1:, :]
    return px.astype(dtype), py.astype(dtype)


def rnnt_loss_pruned(logits, symbols, ranges, termination_symbol, boundary,
                     rnnt_type="regular", delay_penalty=0.0, reduction="mean",
                     calc_gradients=False, dtype=np.float32):
    """rnnt_loss.py:1022-1130.  The reference drops the occupation counts from
    its return value (:1116-1117); they are returned here when asked for."""
    px, py = get_rnnt_logprobs_pruned(logits, symbols, ranges,
                                      termination_symbol, boundary, rnnt_type, dtype)
    return _loss_from_logprobs(px, py, boundary, delay_penalty, reduction,
                               calc_gradients, dtype)


def rnnt_loss(logits, symbols, termination_symbol, boundary, rnnt_type="regular",
              delay_penalty=0.0, reduction="mean", calc_gradients=False,
              dtype=np.float32):
    """rnnt_loss.py:454-551."""
    px, py = get_rnnt_logprobs_joint(logits, symbols, termination_symbol,
                                     boundary, rnnt_type, dtype)
    return _loss_from_logprobs(px, py, boundary, delay_penalty, reduction,
                               calc_gradients, dtype)


# --------------------------------------------------------------------------
# Backward chains TensorFlow's autodiff would produce (A7 bwd, A9)
# --------------------------------------------------------------------------
def pruned_logits_grad(logits, symbols, ranges, termination_symbol, boundary,
                       rnnt_type="regular", delay_penalty=0.0, loss_grad=None,
                       dtype=np.float64, return_scores=False):
    """d(sum_b loss_grad[b] * loss[b]) / d logits for rnnt_loss_pruned with
    reduction='none' (loss = -score): chain of _RNNTLossGrad (__init__.py:154-162)
    through the gathers of rnnt_loss.py:942-1018.
    dlogits[b,t,i,c] = -g_b ( [c=sym'] gx + [c=blank] gy - softmax_c (gx+gy) ),
    gx/gy = occupation of the px/py arc leaving (ranges[b,t,i], t); for
    ``constrained`` the px arc's count also flows into py[s+1,t] (:1018)."""
    dtype = np.dtype(dtype).type
    logits64 = np.asarray(logits).astype(dtype)
    B, T, R, C = logits64.shape
    S = symbols.shape[1]
    px, py = get_rnnt_logprobs_pruned(logits64, symbols, ranges,
                                      termination_symbol, boundary, rnnt_type, dtype)
    px = apply_delay_penalty(px, boundary, delay_penalty)
    scores, (gpx, gpy) = mutual_information_recursion(px, py, boundary, True, dtype)
    if rnnt_type == "constrained":
        gpy = gpy.copy()
        gpy[:, 1:, :] += gpx
    rr = np.asarray(ranges)
    bi = np.arange(B)[:, None, None]
    ti = np.arange(T)[None, :, None]
    gx = np.where(rr < S, gpx[bi, np.minimum(rr, S - 1), ti], 0.0)     # [B,T,R]
    gy = gpy[bi, rr, ti]
    sym_t = np.concatenate(
        [np.asarray(symbols), np.full((B, 1), termination_symbol, dtype=symbols.dtype)], axis=1)
    psym = sym_t[bi, rr]
    m = logits64.max(axis=3, keepdims=True)
    e = np.exp(logits64 - m)
    sm = e / e.sum(axis=3, keepdims=True)
    d = -sm * (gx + gy)[..., None]
    np.put_along_axis(d, psym[..., None].astype(np.int64),
                      np.take_along_axis(d, psym[..., None].astype(np.int64), axis=3) + gx[..., None], axis=3)
    d[:, :, :, termination_symbol] += gy
    g = np.ones((B,), dtype=dtype) if loss_grad is None else np.asarray(loss_grad, dtype=dtype)
    grad = (-g[:, None, None, None]) * d
    return (grad, scores) if return_scores else grad


def simple_am_lm_grad(lm, am, symbols, termination_symbol, boundary,
                      rnnt_type="regular", delay_penalty=0.0, loss_grad=None,
                      dtype=np.float64):
    """d(sum_b loss_grad[b]*loss[b]) / d(am, lm) for rnnt_loss_simple
    (reduction 'none'): TF autodiff through rnnt_loss.py:175-221 (A9)."""
    dtype = np.dtype(dtype).type
    lm = np.asarray(lm, dtype=dtype)
    am = np.asarray(am, dtype=dtype)
    B, T, C = am.shape
    S = lm.shape[1] - 1
    px, py = get_rnnt_logprobs(lm, am, symbols, termination_symbol, rnnt_type,
                               boundary, dtype)
    px = apply_delay_penalty(px, boundary, delay_penalty)
    _, (gpx, gpy) = mutual_information_recursion(px, py, boundary, True, dtype)
    if rnnt_type == "constrained":
        gpy = gpy.copy()
        gpy[:, 1:, :] += gpx
    gpx = gpx[:, :, :T]
    G = gpy.copy()                        # weight on -norm[b,s,t]
    G[:, :S, :] += gpx
    norm, lm_probs, am_probs, lm_max, am_max = _normalizers(lm, am, dtype)
    # norm = log(sum_c lm_probs*am_probs) + maxes ; d norm/d am[b,t,c] = softmax
    Z = np.matmul(lm_probs, np.swapaxes(am_probs, 1, 2)) + dtype(TINY)
    W = G / Z                                                        # [B,S+1,T]
    am_grad = -np.matmul(np.swapaxes(W, 1, 2), lm_probs) * am_probs   # [B,T,C]
    lm_grad = -np.matmul(W, am_probs) * lm_probs                      # [B,S+1,C]
    bi = np.arange(B)[:, None]
    sym = np.asarray(symbols)
    for b in range(B):
        np.add.at(am_grad[b].T, sym[b], gpx[b])                       # am[b,t,sym[s]] += gpx[s,t]
        np.add.at(lm_grad[b], (np.arange(S), sym[b]), gpx[b].sum(axis=1))
    am_grad[:, :, termination_symbol] += gpy.sum(axis=1)
    lm_grad[:, :, termination_symbol] += gpy.sum(axis=2)
    g = np.ones((B,), dtype=dtype) if loss_grad is None else np.asarray(loss_grad, dtype=dtype)
    return -g[:, None, None] * am_grad, -g[:, None, None] * lm_grad


def smoothed_am_lm_grad(lm, am, symbols, termination_symbol, boundary,
                        lm_only_scale=0.1, am_only_scale=0.1,
                        rnnt_type="regular", delay_penalty=0.0, loss_grad=None,
                        dtype=np.float64):
    """d(sum_b loss_grad[b]*loss[b]) / d(am, lm) for rnnt_loss_smoothed
    (reduction 'none'): what TF autodiff derives through rnnt_loss.py:1266-1365
    (A9 for A2), INCLUDING the path through the batch-global unigram
    (rnnt_loss.py:1279-1280) into every lm row."""
    dtype = np.dtype(dtype).type
    lm = np.asarray(lm, dtype=dtype)
    am = np.asarray(am, dtype=dtype)
    sym = np.asarray(symbols)
    B, T, C = am.shape
    S = lm.shape[1] - 1
    blank = termination_symbol
    px, py = get_rnnt_logprobs_smoothed(lm, am, symbols, termination_symbol,
                                        lm_only_scale, am_only_scale, boundary,
                                        rnnt_type, dtype)
    px = apply_delay_penalty(px, boundary, delay_penalty)
    _, (gpx, gpy) = mutual_information_recursion(px, py, boundary, True, dtype)
    if rnnt_type == "constrained":
        gpy = gpy.copy()
        gpy[:, 1:, :] += gpx
    Gx = gpx[:, :, :T]                     # [B,S,T]   weight on px_i
    Gy = gpy                               # [B,S+1,T] weight on py_i
    G = Gy.copy()
    G[:, :S, :] += Gx
    comb = dtype(1.0 - lm_only_scale - am_only_scale)
    lms = dtype(1.0e-20 if lm_only_scale == 0.0 else lm_only_scale)
    ams = dtype(1.0e-20 if am_only_scale == 0.0 else am_only_scale)
    norm, lm_probs, am_probs, lm_max, am_max = _normalizers(lm, am, dtype)
    Z = np.matmul(lm_probs, np.swapaxes(am_probs, 1, 2)) + dtype(TINY)
    W = G / Z
    Zl = lm_probs.sum(axis=2, keepdims=True)
    r = lm_probs / Zl                                              # softmax(lm rows)
    N = B * (S + 1)
    u = r.sum(axis=(0, 1)) / N + dtype(TINY)                       # [C]
    D = am_probs @ u                                               # [B,T]
    q = am_probs * u[None, None, :] / D[:, :, None]                # [B,T,C]
    Gt = G.sum(axis=1)                                             # [B,T]
    Gs = G.sum(axis=2)                                             # [B,S+1]
    Sx = Gx.sum(axis=2)                                            # [B,S]
    Sy = Gy.sum(axis=2)                                            # [B,S+1]
    g = np.ones((B,), dtype=dtype) if loss_grad is None else np.asarray(loss_grad, dtype=dtype)
    g = -g                                                         # loss = -score

    am_grad = -comb * np.matmul(np.swapaxes(W, 1, 2), lm_probs) * am_probs - ams * q * Gt[:, :, None]
    lm_grad = -comb * np.matmul(W, am_probs) * lm_probs - lms * r * Gs[:, :, None]
    for b in range(B):
        np.add.at(am_grad[b].T, sym[b], (comb + ams) * Gx[b])
        np.add.at(lm_grad[b], (np.arange(S), sym[b]), (comb + lms) * Sx[b])
    am_grad[:, :, blank] += (comb + ams) * Gy.sum(axis=1)
    lm_grad[:, :, blank] += (comb + lms) * Sy
    am_grad *= g[:, None, None]
    lm_grad *= g[:, None, None]
    # path through the unigram: d/du[c], then u = mean softmax(lm rows) + tiny
    du = np.zeros((C,), dtype=dtype)
    for b in range(B):
        np.add.at(du, sym[b], g[b] * ams * Sx[b] / u[sym[b]])
        du[blank] += g[b] * ams * Sy[b].sum() / u[blank]
        du -= g[b] * ams * ((Gt[b] / D[b])[:, None] * am_probs[b]).sum(axis=0)
    dot = (r * du[None, None, :]).sum(axis=2, keepdims=True)
    lm_grad += r * (du[None, None, :] - dot) / N
    return am_grad, lm_grad
