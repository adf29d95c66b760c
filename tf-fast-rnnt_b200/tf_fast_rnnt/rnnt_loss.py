"""Host-side mirror of the reference's ``tf_fast_rnnt/rnnt_loss.py`` API.

Same function names, argument names, defaults and return conventions as
/root/reference/tf_fast_rnnt/python/tf_fast_rnnt/rnnt_loss.py; every function is
a thin call into the C ABI (include/fast_rnnt_b200.h) of hand-written sm_100a
kernels.  No math is done here and nothing falls back to the CPU.

Tensors: CUDA ``torch.Tensor`` in -> CUDA ``torch.Tensor`` out (zero copies;
torch is used for device memory and streams only).  ``numpy.ndarray`` (host
buffers) in -> copied to the current device, result copied back, ``numpy`` out;
that is the path a host-resident caller such as a TensorFlow-free harness uses.

Differences from the reference, all decisions of SURVEY.md §9:
  * ``modified`` / ``constrained`` work in the simple/smoothed losses (the
    reference raises a shape error there, rnnt_loss.py:211);
  * ``px_grad`` has the shape of ``px`` (tf_fast_rnnt_op.cc:84 always made it
    [B,S,T+1]);
  * ``reduction="mean"`` of ``rnnt_loss_simple`` works (rnnt_loss.py:331 is a
    NameError);
  * ``boundary=None`` means ``[0, 0, S, T]`` for every utterance.
"""
from __future__ import annotations

import functools
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

Tensor = Union[torch.Tensor, np.ndarray]


# ---------------------------------------------------------------- plumbing
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.FastRnntError(
            "fast_rnnt_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _on_device(fn):
    """Run `fn` with the CUDA device of its tensor arguments current: the C ABI launches on the process's
    current device, and function attributes / streams are per device.  All CUDA inputs must share one device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        devs = {x.device for x in list(args) + list(kwargs.values()) if isinstance(x, torch.Tensor) and x.is_cuda}
        if len(devs) > 1:
            raise ValueError(f"all CUDA inputs must live on one device, got {sorted(map(str, devs))}")
        if not devs:
            return fn(*args, **kwargs)
        with torch.cuda.device(next(iter(devs))):
            return fn(*args, **kwargs)
    return wrapper


class _Io:
    """Remembers whether the caller handed host (numpy) or device buffers."""

    def __init__(self, *xs):
        self.host = any(isinstance(x, np.ndarray) for x in xs if x is not None) or not any(
            isinstance(x, torch.Tensor) and x.is_cuda for x in xs if x is not None)
        devs = [x.device for x in xs if isinstance(x, torch.Tensor) and x.is_cuda]
        self.dev = devs[0] if devs else _device()

    def dev_tensor(self, x, dtype) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            t = x
        else:
            t = torch.from_numpy(np.ascontiguousarray(x))
        if (t.is_cuda and dtype == torch.float32 and t.dtype in (torch.bfloat16, torch.float16)
                and not t.requires_grad):
            # bf16 / fp16 am, lm on the device: widened by the library's own streaming kernel
            src = t.contiguous()
            t = torch.empty(src.shape, dtype=torch.float32, device=src.device)
            check(lib.frn_cast_to_f32(src.data_ptr(), 1 if src.dtype == torch.bfloat16 else 2, src.numel(),
                                      t.data_ptr(), _stream(src.device)), "frn_cast_to_f32")
        if t.dtype != dtype:
            t = t.to(dtype)
        if not t.is_cuda:
            t = t.to(self.dev, non_blocking=True)
        return t.contiguous()

    def out(self, t: torch.Tensor):
        return t.cpu().numpy() if self.host else t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_LP_CODES = {torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


def _low_precision_pair(lm, am, multiple=4):
    """(lm, am, dtype code) when both are bf16 (or both fp16) CUDA tensors the library can consume as they are
    (frn_simple_loss_lp / frn_simple_logprobs_lp - no widening pass, no float32 copies), else None."""
    if not (isinstance(lm, torch.Tensor) and isinstance(am, torch.Tensor) and lm.is_cuda and am.is_cuda):
        return None
    if lm.dtype != am.dtype or lm.dtype not in _LP_CODES or lm.requires_grad or am.requires_grad:
        return None
    if lm.dim() != 3 or am.dim() != 3 or am.shape[2] % multiple != 0:
        return None
    return lm.contiguous(), am.contiguous(), _LP_CODES[lm.dtype]


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _workspace(nbytes: int, dev) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def _boundary(io: _Io, boundary, B: int, S: int, T: int) -> torch.Tensor:
    if boundary is None:
        b = torch.tensor([0, 0, S, T], dtype=torch.int32).repeat(B, 1)
        return b.to(io.dev)
    b = io.dev_tensor(boundary, torch.int32)
    if tuple(b.shape) != (B, 4):
        raise ValueError(f"boundary must have shape ({B}, 4), got {tuple(b.shape)}")
    return b


def _rnnt_type(rnnt_type: str) -> int:
    if rnnt_type not in _lib.RNNT_TYPES:
        raise ValueError(f"rnnt_type must be one of {list(_lib.RNNT_TYPES)}, given {rnnt_type}")
    return _lib.RNNT_TYPES[rnnt_type]


def _logits_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"logits must be float32 or bfloat16, got {t.dtype}")


def _reduce(scores: torch.Tensor, reduction: str, group=None) -> torch.Tensor:
    """loss = -scores with the reference's reductions (rnnt_loss.py:327-338).
    With a ``torch.distributed`` group the sum/mean is completed across ranks by
    one scalar all-reduce (the only collective on this path)."""
    if reduction not in _lib.REDUCTIONS:
        raise ValueError(
            f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")
    B = scores.shape[0]
    red = _lib.REDUCTIONS[reduction]
    out = torch.empty(B if red == _lib.NONE else 1, dtype=torch.float32, device=scores.device)
    denom = 0.0
    if group is not None and red == _lib.MEAN:
        import torch.distributed as dist
        n = torch.tensor([B], dtype=torch.int64, device=scores.device)
        dist.all_reduce(n, group=group)
        denom = float(n.item())
    check(lib.frn_reduce(_ptr(scores), B, red, denom, _ptr(out), _stream(scores.device)), "frn_reduce")
    if red == _lib.NONE:
        return out
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(out, group=group)
    return out.reshape(())


def _unigram_sums(lm_d: torch.Tensor, group) -> Optional[torch.Tensor]:
    """rnnt_loss_smoothed on a batch sharded by utterance: the unigram of rnnt_loss.py:1279-1280 is a mean over
    the GLOBAL batch.  This rank's column sums of softmax(lm rows) and its row count (C+1 floats), all-reduced
    over `group`; the *_sharded entry points turn them into the same unigram on every rank."""
    if group is None:
        return None
    import torch.distributed as dist
    B, S1, C = lm_d.shape
    sums = torch.empty(C + 1, dtype=torch.float32, device=lm_d.device)
    ws = _workspace(lib.frn_simple_logprobs_workspace_bytes(B, S1 - 1, 1, C), lm_d.device)
    check(lib.frn_smoothed_unigram_sums(_ptr(lm_d), B, S1 - 1, C, _ptr(sums), _ptr(ws), ws.numel(),
                                        _stream(lm_d.device)), "frn_smoothed_unigram_sums")
    dist.all_reduce(sums, group=group)
    return sums


def _reduce_autograd(scores: torch.Tensor, reduction: str, group=None) -> torch.Tensor:
    """The same reductions on the autograd path.  With a process group the batch is sharded by utterance:
    the returned VALUE is the global sum / mean (one scalar all-reduce of the detached local part), the
    GRADIENT is that of the local part at the global scale (1 for 'sum', 1 / N_global for 'mean'), so every
    rank's am / lm / logits gradients are exactly those of the unsharded loss."""
    if reduction == "none":
        return -scores
    if group is None:
        return -scores.sum() if reduction == "sum" else -scores.mean()
    import torch.distributed as dist
    local = -scores.sum()
    if reduction == "mean":
        n = torch.tensor([scores.shape[0]], dtype=torch.int64, device=scores.device)
        dist.all_reduce(n, group=group)
        local = local / float(n.item())
    total = local.detach().clone()
    dist.all_reduce(total, group=group)
    return local + (total - local.detach())


# ---------------------------------------------------------------- A4 / cummin
class _MutualInformationFn(torch.autograd.Function):
    """FastRNNTLoss with the gradient the reference registers for it (__init__.py:154-162):
    d ans[b] / d px = px_grad[b], d ans[b] / d py = py_grad[b] (the occupation counts)."""

    @staticmethod
    def forward(ctx, px, py, boundary):
        B, S, T1 = px.shape
        T = py.shape[2]
        ans = torch.empty(B, dtype=torch.float32, device=px.device)
        gx, gy = torch.empty_like(px), torch.empty_like(py)
        ws = _workspace(lib.frn_mi_workspace_bytes(B, S, T, T1), px.device)
        check(lib.frn_mi_fwd_bwd(_ptr(px), _ptr(py), _ptr(boundary), B, S, T, T1, 1, _ptr(ans), _ptr(gx), _ptr(gy),
                                 _ptr(ws), ws.numel(), _stream(px.device)), "frn_mi_fwd_bwd")
        ctx.save_for_backward(gx, gy)
        ctx.mark_non_differentiable(gx, gy)
        return ans, gx, gy

    @staticmethod
    def backward(ctx, g, _gx, _gy):
        gx, gy = ctx.saved_tensors
        g = g.reshape(-1, 1, 1)
        return g * gx, g * gy, None


@_on_device
def mutual_information_recursion(px: Tensor, py: Tensor, boundary: Optional[Tensor] = None,
                                 calc_gradients: bool = False):
    """Reference: tf_fast_rnnt/__init__.py:42-149 (op FastRNNTLoss)."""
    io = _Io(px, py)
    px_d = io.dev_tensor(px, torch.float32)
    py_d = io.dev_tensor(py, torch.float32)
    if px_d.dim() != 3 or py_d.dim() != 3:
        raise ValueError("px and py must be 3-dimensional")
    B, S, T1 = px_d.shape
    T = py_d.shape[2]
    if T1 not in (T, T + 1) or tuple(py_d.shape) != (B, S + 1, T):
        raise ValueError(f"bad shapes px {tuple(px_d.shape)} py {tuple(py_d.shape)}")
    bd = _boundary(io, boundary, B, S, T)
    if _wants_grad(px, py):
        ans, gx, gy = _MutualInformationFn.apply(px_d, py_d, bd)
        return (ans, (gx, gy)) if calc_gradients else ans
    ans = torch.empty(B, dtype=torch.float32, device=io.dev)
    gx = torch.empty_like(px_d) if calc_gradients else None
    gy = torch.empty_like(py_d) if calc_gradients else None
    nbytes = lib.frn_mi_workspace_bytes(B, S, T, T1)
    ws = _workspace(nbytes, io.dev)
    check(lib.frn_mi_fwd_bwd(_ptr(px_d), _ptr(py_d), _ptr(bd), B, S, T, T1, int(calc_gradients),
                             _ptr(ans), _ptr(gx), _ptr(gy), _ptr(ws), ws.numel(), _stream(io.dev)),
          "frn_mi_fwd_bwd")
    if calc_gradients:
        return io.out(ans), (io.out(gx), io.out(gy))
    return io.out(ans)


@_on_device
def cummin(x: Tensor):
    """Reference: tf_fast_rnnt/__init__.py:151-152 (op Cummin)."""
    io = _Io(x)
    x_d = io.dev_tensor(x, torch.int32)
    if x_d.dim() != 2:
        raise ValueError("cummin expects a 2-D int32 tensor")
    out = torch.empty_like(x_d)
    check(lib.frn_cummin(_ptr(x_d), _ptr(out), x_d.shape[0], x_d.shape[1], _stream(io.dev)), "frn_cummin")
    return io.out(out)


# ---------------------------------------------------------------- A1 / A2
def _simple_logprobs(lm, am, symbols, termination_symbol, rnnt_type, boundary, smoothed,
                     lm_only_scale, am_only_scale, usums=None):
    io = _Io(lm, am)
    lp = _low_precision_pair(lm, am) if usums is None else None
    lm_d = lp[0] if lp else io.dev_tensor(lm, torch.float32)
    am_d = lp[1] if lp else io.dev_tensor(am, torch.float32)
    sym_d = io.dev_tensor(symbols, torch.int32)
    B, T, C = am_d.shape
    S = lm_d.shape[1] - 1
    if tuple(lm_d.shape) != (B, S + 1, C) or tuple(sym_d.shape) != (B, S):
        raise ValueError("lm must be [B,S+1,C], am [B,T,C], symbols [B,S]")
    rt = _rnnt_type(rnnt_type)
    bd = _boundary(io, boundary, B, S, T)
    T1 = T + 1 if rt == _lib.REGULAR else T
    px = torch.empty((B, S, T1), dtype=torch.float32, device=io.dev)
    py = torch.empty((B, S + 1, T), dtype=torch.float32, device=io.dev)
    ws = _workspace(lib.frn_simple_logprobs_workspace_bytes(B, S, T, C), io.dev)
    if lp:
        rc = lib.frn_simple_logprobs_lp(_ptr(lm_d), _ptr(am_d), lp[2], _ptr(sym_d), _ptr(bd), B, S, T, C,
                                        int(termination_symbol), rt, int(smoothed), float(lm_only_scale),
                                        float(am_only_scale), None, _ptr(px), _ptr(py), _ptr(ws), ws.numel(),
                                        _stream(io.dev))
        if rc != _lib.EUNSUPPORTED:
            check(rc, "frn_simple_logprobs_lp")
            return io.out(px), io.out(py)
        lm_d, am_d = io.dev_tensor(lm, torch.float32), io.dev_tensor(am, torch.float32)    # widen, then the float32 path
    check(lib.frn_simple_logprobs_sharded(_ptr(lm_d), _ptr(am_d), _ptr(sym_d), _ptr(bd), B, S, T, C,
                                          int(termination_symbol), rt, int(smoothed), float(lm_only_scale),
                                          float(am_only_scale), _ptr(usums), _ptr(px), _ptr(py), _ptr(ws), ws.numel(),
                                          _stream(io.dev)), "frn_simple_logprobs")
    return io.out(px), io.out(py)


class _SimpleLogprobsFn(torch.autograd.Function):
    """(px, py)(lm, am) with the backward TensorFlow autodiff runs through rnnt_loss.py:175-221 (1266-1365 for
    the smoothed variant): the same two contractions as the loss gradient, fed with the cotangents of px / py
    instead of occupation counts."""

    @staticmethod
    def forward(ctx, lm, am, symbols, termination_symbol, boundary, rnnt_type, smoothed, lm_only_scale,
                am_only_scale, group=None):
        usums = _unigram_sums(lm, group) if smoothed else None
        px, py = _simple_logprobs(lm, am, symbols, termination_symbol, rnnt_type, boundary, smoothed,
                                  lm_only_scale, am_only_scale, usums)
        ctx.save_for_backward(lm, am, symbols, boundary)
        ctx.args = (termination_symbol, rnnt_type, smoothed, lm_only_scale, am_only_scale)
        ctx.group, ctx.usums = group, usums
        return px, py

    @staticmethod
    def backward(ctx, dpx, dpy):
        lm, am, symbols, boundary = ctx.saved_tensors
        term, rnnt_type, smoothed, lms, ams = ctx.args
        am_g, lm_g = simple_loss_backward(lm, am, symbols, term, boundary, dpx.contiguous(), dpy.contiguous(), None,
                                          rnnt_type, smoothed, lms, ams, ctx.group, ctx.usums)
        return lm_g, am_g, None, None, None, None, None, None, None, None


def _simple_logprobs_public(lm, am, symbols, termination_symbol, rnnt_type, boundary, smoothed, lms, ams, group=None):
    if _wants_grad(lm, am):
        io = _Io(lm, am)
        B, T, _ = am.shape
        S = lm.shape[1] - 1
        return _SimpleLogprobsFn.apply(lm.contiguous().float(), am.contiguous().float(),
                                       io.dev_tensor(symbols, torch.int32), int(termination_symbol),
                                       _boundary(io, boundary, B, S, T), rnnt_type, bool(smoothed), float(lms),
                                       float(ams), group)
    usums = None
    if smoothed and group is not None:
        usums = _unigram_sums(_Io(lm, am).dev_tensor(lm, torch.float32), group)
    return _simple_logprobs(lm, am, symbols, termination_symbol, rnnt_type, boundary, smoothed, lms, ams, usums)


@_on_device
def get_rnnt_logprobs(lm: Tensor, am: Tensor, symbols: Tensor, termination_symbol: int,
                      rnnt_type: str = "regular", boundary: Optional[Tensor] = None):
    """Reference: rnnt_loss.py:63-223.  CUDA tensors that require grad get gradients w.r.t. lm and am."""
    return _simple_logprobs_public(lm, am, symbols, termination_symbol, rnnt_type, boundary, False, 0.0, 0.0)


@_on_device
def get_rnnt_logprobs_smoothed(lm: Tensor, am: Tensor, symbols: Tensor, termination_symbol: int,
                               lm_only_scale: float = 0.1, am_only_scale: float = 0.1,
                               boundary: Optional[Tensor] = None, rnnt_type: str = "regular", group=None):
    """Reference: rnnt_loss.py:1132-1367.  CUDA tensors that require grad get gradients w.r.t. lm and am.
    ``group``: the batch is sharded by utterance over that torch.distributed group (exact batch-global unigram)."""
    return _simple_logprobs_public(lm, am, symbols, termination_symbol, rnnt_type, boundary, True,
                                   lm_only_scale, am_only_scale, group)


def _simple_loss(lm, am, symbols, termination_symbol, boundary, rnnt_type, delay_penalty, reduction,
                 calc_gradients, smoothed, lm_only_scale, am_only_scale, group=None):
    if reduction not in _lib.REDUCTIONS:
        raise ValueError(
            f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")
    io = _Io(lm, am)
    # bf16 / fp16 am, lm on the device are consumed as they are (the sharded smoothed loss widens them: its
    # unigram sums are formed by a float32 entry point)
    lp = _low_precision_pair(lm, am) if not (smoothed and group is not None) else None
    lm_d = lp[0] if lp else io.dev_tensor(lm, torch.float32)
    am_d = lp[1] if lp else io.dev_tensor(am, torch.float32)
    sym_d = io.dev_tensor(symbols, torch.int32)
    B, T, C = am_d.shape
    S = lm_d.shape[1] - 1
    if tuple(lm_d.shape) != (B, S + 1, C) or tuple(sym_d.shape) != (B, S):
        raise ValueError("lm must be [B,S+1,C], am [B,T,C], symbols [B,S]")
    rt = _rnnt_type(rnnt_type)
    bd = _boundary(io, boundary, B, S, T)
    T1 = T + 1 if rt == _lib.REGULAR else T
    scores = torch.empty(B, dtype=torch.float32, device=io.dev)
    gx = torch.empty((B, S, T1), dtype=torch.float32, device=io.dev) if calc_gradients else None
    gy = torch.empty((B, S + 1, T), dtype=torch.float32, device=io.dev) if calc_gradients else None
    ws = _workspace(lib.frn_simple_loss_workspace_bytes(B, S, T, C), io.dev)
    dp = float(delay_penalty) if delay_penalty > 0.0 else 0.0
    if lp:
        rc = lib.frn_simple_loss_lp(_ptr(lm_d), _ptr(am_d), lp[2], _ptr(sym_d), _ptr(bd), B, S, T, C,
                                    int(termination_symbol), rt, int(smoothed), float(lm_only_scale),
                                    float(am_only_scale), None, dp, int(calc_gradients), _ptr(scores), _ptr(gx), _ptr(gy),
                                    _ptr(ws), ws.numel(), _stream(io.dev))
        if rc != _lib.EUNSUPPORTED:
            check(rc, "frn_simple_loss_lp")
            loss = io.out(_reduce(scores, reduction, group))
            return (loss, (io.out(gx), io.out(gy))) if calc_gradients else loss
        lm_d, am_d = io.dev_tensor(lm, torch.float32), io.dev_tensor(am, torch.float32)    # widen, then the float32 path
    usums = _unigram_sums(lm_d, group) if smoothed else None
    check(lib.frn_simple_loss_sharded(_ptr(lm_d), _ptr(am_d), _ptr(sym_d), _ptr(bd), B, S, T, C,
                                      int(termination_symbol), rt, int(smoothed), float(lm_only_scale),
                                      float(am_only_scale), _ptr(usums), dp, int(calc_gradients), _ptr(scores),
                                      _ptr(gx), _ptr(gy), _ptr(ws), ws.numel(), _stream(io.dev)), "frn_simple_loss")
    loss = io.out(_reduce(scores, reduction, group))
    return (loss, (io.out(gx), io.out(gy))) if calc_gradients else loss


@_on_device
def simple_loss_backward(lm, am, symbols, termination_symbol, boundary, px_grad, py_grad, scores_grad=None,
                         rnnt_type: str = "regular", smoothed: bool = False, lm_only_scale: float = 0.0,
                         am_only_scale: float = 0.0, group=None, unigram_sums=None):
    """d(sum_b scores_grad[b] * scores[b]) / d(am, lm) for rnnt_loss_simple, from the
    occupation counts (px_grad, py_grad) that ``rnnt_loss_simple(..., calc_gradients=True)``
    returned — the chain TensorFlow autodiff runs through rnnt_loss.py:175-221.
    ``smoothed``: the same for rnnt_loss_smoothed (rnnt_loss.py:1266-1365, with the
    path through the batch-global unigram).  ``group`` (smoothed only): the batch is sharded by utterance over
    that torch.distributed group - the unigram sums and d loss / d unigram are all-reduced (C+1 and C floats),
    and every rank gets the gradients of the unsharded loss for its own utterances."""
    io = _Io(lm, am)
    lm_d = io.dev_tensor(lm, torch.float32)
    am_d = io.dev_tensor(am, torch.float32)
    sym_d = io.dev_tensor(symbols, torch.int32)
    gx = io.dev_tensor(px_grad, torch.float32)
    gy = io.dev_tensor(py_grad, torch.float32)
    B, T, C = am_d.shape
    S = lm_d.shape[1] - 1
    rt = _rnnt_type(rnnt_type)
    bd = _boundary(io, boundary, B, S, T)
    sg = None if scores_grad is None else io.dev_tensor(scores_grad, torch.float32)
    am_g = torch.empty_like(am_d)
    lm_g = torch.empty_like(lm_d)
    ws = _workspace(lib.frn_simple_loss_bwd_workspace_bytes(B, S, T, C), io.dev)
    if smoothed and (group is not None or unigram_sums is not None):
        import torch.distributed as dist
        usums = unigram_sums if unigram_sums is not None else _unigram_sums(lm_d, group)
        du = torch.empty(C, dtype=torch.float32, device=io.dev)
        for phase in (1, 2):
            check(lib.frn_smoothed_loss_bwd_sharded(_ptr(lm_d), _ptr(am_d), _ptr(sym_d), _ptr(bd), _ptr(gx), _ptr(gy),
                                                    _ptr(sg), B, S, T, C, int(termination_symbol), rt,
                                                    float(lm_only_scale), float(am_only_scale), _ptr(usums), _ptr(du),
                                                    phase, _ptr(am_g), _ptr(lm_g), _ptr(ws), ws.numel(),
                                                    _stream(io.dev)), "frn_smoothed_loss_bwd_sharded")
            if phase == 1 and group is not None:
                dist.all_reduce(du, group=group)
    elif smoothed:
        check(lib.frn_smoothed_loss_bwd(_ptr(lm_d), _ptr(am_d), _ptr(sym_d), _ptr(bd), _ptr(gx), _ptr(gy), _ptr(sg),
                                        B, S, T, C, int(termination_symbol), rt, float(lm_only_scale),
                                        float(am_only_scale), _ptr(am_g), _ptr(lm_g), _ptr(ws), ws.numel(),
                                        _stream(io.dev)), "frn_smoothed_loss_bwd")
    else:
        check(lib.frn_simple_loss_bwd(_ptr(lm_d), _ptr(am_d), _ptr(sym_d), _ptr(bd), _ptr(gx), _ptr(gy), _ptr(sg),
                                      B, S, T, C, int(termination_symbol), rt, _ptr(am_g), _ptr(lm_g), _ptr(ws),
                                      ws.numel(), _stream(io.dev)), "frn_simple_loss_bwd")
    return io.out(am_g), io.out(lm_g)


def smoothed_loss_backward(lm, am, symbols, termination_symbol, boundary, px_grad, py_grad, scores_grad=None,
                           lm_only_scale: float = 0.1, am_only_scale: float = 0.1, rnnt_type: str = "regular"):
    """A9 for rnnt_loss_smoothed: see simple_loss_backward."""
    return simple_loss_backward(lm, am, symbols, termination_symbol, boundary, px_grad, py_grad, scores_grad,
                                rnnt_type, True, lm_only_scale, am_only_scale)


class _SimpleLossFn(torch.autograd.Function):
    """scores(lm, am) with gradients; stands where TF's GradientTape +
    _RNNTLossGrad (__init__.py:154-162) stand in the reference."""

    @staticmethod
    def forward(ctx, lm, am, symbols, termination_symbol, boundary, rnnt_type, delay_penalty, smoothed=False,
                lm_only_scale=0.0, am_only_scale=0.0, group=None):
        B, T, C = am.shape
        S = lm.shape[1] - 1
        rt = _rnnt_type(rnnt_type)
        T1 = T + 1 if rt == _lib.REGULAR else T
        scores = torch.empty(B, dtype=torch.float32, device=am.device)
        gx = torch.empty((B, S, T1), dtype=torch.float32, device=am.device)
        gy = torch.empty((B, S + 1, T), dtype=torch.float32, device=am.device)
        ws = _workspace(lib.frn_simple_loss_workspace_bytes(B, S, T, C), am.device)
        dp = float(delay_penalty) if delay_penalty > 0.0 else 0.0
        usums = _unigram_sums(lm, group) if smoothed else None
        check(lib.frn_simple_loss_sharded(_ptr(lm), _ptr(am), _ptr(symbols), _ptr(boundary), B, S, T, C,
                                          int(termination_symbol), rt, int(bool(smoothed)), float(lm_only_scale),
                                          float(am_only_scale), _ptr(usums), dp, 1, _ptr(scores), _ptr(gx), _ptr(gy),
                                          _ptr(ws), ws.numel(), _stream(am.device)), "frn_simple_loss")
        ctx.save_for_backward(lm, am, symbols, boundary, gx, gy)
        ctx.args = (termination_symbol, rnnt_type, bool(smoothed), float(lm_only_scale), float(am_only_scale))
        ctx.group, ctx.usums = group, usums
        ctx.mark_non_differentiable(gx, gy)
        return scores, gx, gy

    @staticmethod
    def backward(ctx, g, _gx, _gy):
        lm, am, symbols, boundary, gx, gy = ctx.saved_tensors
        term, rnnt_type, smoothed, lms, ams = ctx.args
        am_g, lm_g = simple_loss_backward(lm, am, symbols, term, boundary, gx, gy, g.contiguous(), rnnt_type,
                                          smoothed, lms, ams, ctx.group, ctx.usums)
        return lm_g, am_g, None, None, None, None, None, None, None, None, None


@_on_device
def rnnt_loss_simple(lm: Tensor, am: Tensor, symbols: Tensor, termination_symbol: int,
                     boundary: Optional[Tensor] = None, rnnt_type: str = "regular",
                     delay_penalty: float = 0.0, reduction: Optional[str] = "mean",
                     calc_gradients: bool = False, group=None):
    """Reference: rnnt_loss.py:225-338.  ``group``: optional torch.distributed
    process group over which 'sum'/'mean' are completed (batch sharded by
    utterance).  CUDA tensors that require grad get gradients w.r.t. lm and am."""
    if (isinstance(lm, torch.Tensor) and isinstance(am, torch.Tensor) and lm.is_cuda and am.is_cuda
            and (lm.requires_grad or am.requires_grad) and torch.is_grad_enabled()):
        if reduction not in _lib.REDUCTIONS:
            raise ValueError(
                f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")
        io = _Io(lm, am)
        B, T, _ = am.shape
        S = lm.shape[1] - 1
        scores, gx, gy = _SimpleLossFn.apply(lm.contiguous().float(), am.contiguous().float(),
                                             io.dev_tensor(symbols, torch.int32), int(termination_symbol),
                                             _boundary(io, boundary, B, S, T), rnnt_type, float(delay_penalty))
        loss = _reduce_autograd(scores, reduction, group)
        return (loss, (gx, gy)) if calc_gradients else loss
    return _simple_loss(lm, am, symbols, termination_symbol, boundary, rnnt_type, delay_penalty,
                        reduction, calc_gradients, False, 0.0, 0.0, group)


@_on_device
def rnnt_loss_smoothed(lm: Tensor, am: Tensor, symbols: Tensor, termination_symbol: int,
                       lm_only_scale: float = 0.1, am_only_scale: float = 0.1,
                       boundary: Optional[Tensor] = None, rnnt_type: str = "regular",
                       delay_penalty: float = 0.0, reduction: Optional[str] = "mean",
                       calc_gradients: bool = False, group=None):
    """Reference: rnnt_loss.py:1369-1494.  CUDA tensors that require grad get gradients
    w.r.t. lm and am (through the batch-global unigram as well)."""
    if (isinstance(lm, torch.Tensor) and isinstance(am, torch.Tensor) and lm.is_cuda and am.is_cuda
            and (lm.requires_grad or am.requires_grad) and torch.is_grad_enabled()):
        if reduction not in _lib.REDUCTIONS:
            raise ValueError(
                f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")
        io = _Io(lm, am)
        B, T, _ = am.shape
        S = lm.shape[1] - 1
        scores, gx, gy = _SimpleLossFn.apply(lm.contiguous().float(), am.contiguous().float(),
                                             io.dev_tensor(symbols, torch.int32), int(termination_symbol),
                                             _boundary(io, boundary, B, S, T), rnnt_type, float(delay_penalty),
                                             True, float(lm_only_scale), float(am_only_scale), group)
        loss = _reduce_autograd(scores, reduction, group)
        return (loss, (gx, gy)) if calc_gradients else loss
    return _simple_loss(lm, am, symbols, termination_symbol, boundary, rnnt_type, delay_penalty,
                        reduction, calc_gradients, True, lm_only_scale, am_only_scale, group)


# ---------------------------------------------------------------- A5 / A6
@_on_device
def get_rnnt_prune_ranges(px_grad: Tensor, py_grad: Tensor, boundary: Tensor, s_range: int):
    """Reference: rnnt_loss.py:647-761."""
    io = _Io(px_grad, py_grad)
    gx = io.dev_tensor(px_grad, torch.float32)
    gy = io.dev_tensor(py_grad, torch.float32)
    B, S, T1 = gx.shape
    T = gy.shape[2]
    if T1 not in (T, T + 1) or tuple(gy.shape) != (B, S + 1, T):
        raise ValueError(f"bad shapes px_grad {tuple(gx.shape)} py_grad {tuple(gy.shape)}")
    bd = _boundary(io, boundary, B, S, T)
    R = lib.frn_prune_ranges_width(S, int(s_range))
    ranges = torch.empty((B, T, R), dtype=torch.int32, device=io.dev)
    ws = _workspace(lib.frn_prune_ranges_workspace_bytes(B, T), io.dev)
    check(lib.frn_prune_ranges(_ptr(gx), _ptr(gy), _ptr(bd), B, S, T, T1, int(s_range), _ptr(ranges),
                               _ptr(ws), ws.numel(), _stream(io.dev)), "frn_prune_ranges")
    return io.out(ranges)


class _PruningFn(torch.autograd.Function):
    """do_rnnt_pruning (optionally with the additive joiner) with the gradient TF autodiff derives for
    rnnt_loss.py:802-811: am_grad = sum over the band, lm_grad = scatter-add over ranges."""

    @staticmethod
    def forward(ctx, am, lm, ranges, with_joiner):
        B, T, C = am.shape
        S = lm.shape[1] - 1
        R = ranges.shape[2]
        outs = [torch.empty((B, T, R, C), dtype=torch.float32, device=am.device) for _ in range(3 if with_joiner else 2)]
        if with_joiner:
            check(lib.frn_do_pruning_add_joiner(_ptr(am), _ptr(lm), _ptr(ranges), B, S, T, R, C, _ptr(outs[0]),
                                                _ptr(outs[1]), _ptr(outs[2]), _stream(am.device)),
                  "frn_do_pruning_add_joiner")
        else:
            check(lib.frn_do_pruning(_ptr(am), _ptr(lm), _ptr(ranges), B, S, T, R, C, _ptr(outs[0]), _ptr(outs[1]),
                                     _stream(am.device)), "frn_do_pruning")
        ctx.save_for_backward(ranges)
        ctx.S = S
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        (ranges,) = ctx.saved_tensors
        ga, gl = grads[0], grads[1]
        if len(grads) == 3:                     # logits = am_pruned + lm_pruned feeds both
            ga, gl = ga + grads[2], gl + grads[2]
        am_g, lm_g = do_rnnt_pruning_backward(ga.contiguous(), gl.contiguous(), ranges, ctx.S)
        return am_g, lm_g, None, None


def _wants_grad(*xs) -> bool:
    return torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.is_cuda and x.requires_grad for x in xs)


@_on_device
def do_rnnt_pruning(am: Tensor, lm: Tensor, ranges: Tensor):
    """Reference: rnnt_loss.py:763-812.  CUDA tensors that require grad get gradients w.r.t. am and lm."""
    if _wants_grad(am, lm):
        io = _Io(am, lm)
        return _PruningFn.apply(io.dev_tensor(am, torch.float32), io.dev_tensor(lm, torch.float32),
                                io.dev_tensor(ranges, torch.int32), False)
    io = _Io(am, lm)
    # bf16 / fp16 am, lm on the device keep their type, like the reference's broadcast_to / gather (an even C)
    lp = _low_precision_pair(lm, am, multiple=2)
    lm_d = lp[0] if lp else io.dev_tensor(lm, torch.float32)
    am_d = lp[1] if lp else io.dev_tensor(am, torch.float32)
    rg = io.dev_tensor(ranges, torch.int32)
    B, T, C = am_d.shape
    S = lm_d.shape[1] - 1
    R = rg.shape[2]
    if tuple(rg.shape) != (B, T, R) or tuple(lm_d.shape) != (B, S + 1, C):
        raise ValueError("am [B,T,C], lm [B,S+1,C], ranges [B,T,s_range] expected")
    am_p = torch.empty((B, T, R, C), dtype=am_d.dtype, device=io.dev)
    lm_p = torch.empty((B, T, R, C), dtype=am_d.dtype, device=io.dev)
    if lp:
        check(lib.frn_do_pruning_lp(_ptr(am_d), _ptr(lm_d), lp[2], _ptr(rg), B, S, T, R, C, _ptr(am_p), _ptr(lm_p),
                                    _stream(io.dev)), "frn_do_pruning_lp")
        return am_p, lm_p
    check(lib.frn_do_pruning(_ptr(am_d), _ptr(lm_d), _ptr(rg), B, S, T, R, C, _ptr(am_p), _ptr(lm_p),
                             _stream(io.dev)), "frn_do_pruning")
    return io.out(am_p), io.out(lm_p)


@_on_device
def do_rnnt_pruning_backward(am_pruned_grad: Tensor, lm_pruned_grad: Tensor, ranges: Tensor, S: int):
    """Gradient of do_rnnt_pruning (what TF autodiff derives for rnnt_loss.py:802-811)."""
    io = _Io(am_pruned_grad, lm_pruned_grad)
    ga = io.dev_tensor(am_pruned_grad, torch.float32)
    gl = io.dev_tensor(lm_pruned_grad, torch.float32)
    rg = io.dev_tensor(ranges, torch.int32)
    B, T, R, C = ga.shape
    am_g = torch.empty((B, T, C), dtype=torch.float32, device=io.dev)
    lm_g = torch.empty((B, S + 1, C), dtype=torch.float32, device=io.dev)
    check(lib.frn_do_pruning_bwd(_ptr(ga), _ptr(gl), _ptr(rg), B, S, T, R, C, _ptr(am_g), _ptr(lm_g),
                                 _stream(io.dev)), "frn_do_pruning_bwd")
    return io.out(am_g), io.out(lm_g)


@_on_device
def do_rnnt_pruning_add_joiner(am: Tensor, lm: Tensor, ranges: Tensor):
    """(extension) do_rnnt_pruning (rnnt_loss.py:763-812) plus the additive joiner of the
    reference's tests (simple_rnnt_loss_test.py:120-125) in one pass over the data:
    -> (am_pruned, lm_pruned, am_pruned + lm_pruned)."""
    if _wants_grad(am, lm):
        io = _Io(am, lm)
        return _PruningFn.apply(io.dev_tensor(am, torch.float32), io.dev_tensor(lm, torch.float32),
                                io.dev_tensor(ranges, torch.int32), True)
    io = _Io(am, lm)
    am_d = io.dev_tensor(am, torch.float32)
    lm_d = io.dev_tensor(lm, torch.float32)
    rg = io.dev_tensor(ranges, torch.int32)
    B, T, C = am_d.shape
    S = lm_d.shape[1] - 1
    R = rg.shape[2]
    if tuple(rg.shape) != (B, T, R) or tuple(lm_d.shape) != (B, S + 1, C):
        raise ValueError("am [B,T,C], lm [B,S+1,C], ranges [B,T,s_range] expected")
    am_p, lm_p, lg = (torch.empty((B, T, R, C), dtype=torch.float32, device=io.dev) for _ in range(3))
    check(lib.frn_do_pruning_add_joiner(_ptr(am_d), _ptr(lm_d), _ptr(rg), B, S, T, R, C, _ptr(am_p), _ptr(lm_p),
                                        _ptr(lg), _stream(io.dev)), "frn_do_pruning_add_joiner")
    return io.out(am_p), io.out(lm_p), io.out(lg)


@_on_device
def pruned_add_joiner(am: Tensor, lm: Tensor, ranges: Tensor, dtype=torch.float32):
    """(extension) logits = am_pruned + lm_pruned without
    materialising either."""
    io = _Io(am, lm)
    am_d = io.dev_tensor(am, torch.float32)
    lm_d = io.dev_tensor(lm, torch.float32)
    rg = io.dev_tensor(ranges, torch.int32)
    B, T, C = am_d.shape
    S = lm_d.shape[1] - 1
    R = rg.shape[2]
    out = torch.empty((B, T, R, C), dtype=dtype, device=io.dev)
    check(lib.frn_pruned_add_joiner(_ptr(am_d), _ptr(lm_d), _ptr(rg), B, S, T, R, C, _logits_dtype(out),
                                    _ptr(out), _stream(io.dev)), "frn_pruned_add_joiner")
    return io.out(out) if dtype == torch.float32 else out


# ---------------------------------------------------------------- A7 / A8
class _PrunedLogprobsFn(torch.autograd.Function):
    """(px, py)(logits) with the backward TF autodiff runs through rnnt_loss.py:942-1018 (frn_pruned_logprobs_bwd)."""

    @staticmethod
    def forward(ctx, logits, symbols, ranges, termination_symbol, boundary, rnnt_type):
        px, py = _pruned_logprobs(logits, symbols, ranges, termination_symbol, boundary, rnnt_type)
        ctx.save_for_backward(logits, symbols, ranges, boundary)
        ctx.args = (termination_symbol, rnnt_type)
        return px, py

    @staticmethod
    def backward(ctx, dpx, dpy):
        logits, symbols, ranges, boundary = ctx.saved_tensors
        term, rnnt_type = ctx.args
        B, T, R, C = logits.shape
        S = symbols.shape[1]
        grad = torch.empty_like(logits)
        ws = _workspace(lib.frn_pruned_logprobs_workspace_bytes(B, S, T, R), logits.device)
        check(lib.frn_pruned_logprobs_bwd(_ptr(logits), _logits_dtype(logits), _ptr(symbols), _ptr(ranges),
                                          _ptr(boundary), _ptr(dpx.contiguous().float()), _ptr(dpy.contiguous().float()),
                                          B, S, T, R, C, int(term), _rnnt_type(rnnt_type), _ptr(grad), _ptr(ws),
                                          ws.numel(), _stream(logits.device)), "frn_pruned_logprobs_bwd")
        return grad, None, None, None, None, None


@_on_device
def get_rnnt_logprobs_pruned(logits: Tensor, symbols: Tensor, ranges: Tensor, termination_symbol: int,
                             boundary: Tensor, rnnt_type: str = "regular"):
    """Reference: rnnt_loss.py:853-1020.  CUDA logits that require grad get their gradient."""
    if _wants_grad(logits):
        io = _Io(logits)
        sym_d = io.dev_tensor(symbols, torch.int32)
        B, T = logits.shape[0], logits.shape[1]
        return _PrunedLogprobsFn.apply(logits.contiguous(), sym_d, io.dev_tensor(ranges, torch.int32),
                                       int(termination_symbol), _boundary(io, boundary, B, sym_d.shape[1], T), rnnt_type)
    return _pruned_logprobs(logits, symbols, ranges, termination_symbol, boundary, rnnt_type)


def _pruned_logprobs(logits, symbols, ranges, termination_symbol, boundary, rnnt_type):
    io = _Io(logits)
    lg = logits if isinstance(logits, torch.Tensor) and logits.is_cuda else io.dev_tensor(logits, torch.float32)
    lg = lg.contiguous()
    sym_d = io.dev_tensor(symbols, torch.int32)
    rg = io.dev_tensor(ranges, torch.int32)
    B, T, R, C = lg.shape
    S = sym_d.shape[1]
    rt = _rnnt_type(rnnt_type)
    bd = _boundary(io, boundary, B, S, T)
    T1 = T + 1 if rt == _lib.REGULAR else T
    px = torch.empty((B, S, T1), dtype=torch.float32, device=io.dev)
    py = torch.empty((B, S + 1, T), dtype=torch.float32, device=io.dev)
    ws = _workspace(lib.frn_pruned_logprobs_workspace_bytes(B, S, T, R), io.dev)
    check(lib.frn_pruned_logprobs(_ptr(lg), _logits_dtype(lg), _ptr(sym_d), _ptr(rg), _ptr(bd), B, S, T, R, C,
                                  int(termination_symbol), rt, _ptr(px), _ptr(py), _ptr(ws), ws.numel(),
                                  _stream(io.dev)), "frn_pruned_logprobs")
    return io.out(px), io.out(py)


@_on_device
def pruned_loss_fwd_bwd(logits, symbols, ranges, termination_symbol, boundary, rnnt_type="regular",
                        delay_penalty=0.0, scores_grad=None, want_logits_grad=True):
    """One fused call: scores [B] and d(sum_b scores_grad[b]*scores[b])/d logits
    (scores_grad None = ones).  Device tensors only."""
    dev = logits.device
    io = _Io(logits)
    lg = logits.contiguous()
    sym_d = io.dev_tensor(symbols, torch.int32)
    rg = io.dev_tensor(ranges, torch.int32)
    B, T, R, C = lg.shape
    S = sym_d.shape[1]
    rt = _rnnt_type(rnnt_type)
    bd = _boundary(io, boundary, B, S, T)
    scores = torch.empty(B, dtype=torch.float32, device=dev)
    grad = torch.empty_like(lg) if want_logits_grad else None
    sg = None if scores_grad is None else io.dev_tensor(scores_grad, torch.float32)
    dp = float(delay_penalty) if delay_penalty > 0.0 else 0.0
    ws = _workspace(lib.frn_pruned_loss_min_workspace_bytes(B, S, T, R, dp), dev)
    check(lib.frn_pruned_loss(_ptr(lg), _logits_dtype(lg), _ptr(sym_d), _ptr(rg), _ptr(bd), B, S, T, R, C,
                              int(termination_symbol), rt, dp, _ptr(sg), _ptr(scores), _ptr(grad), _ptr(ws),
                              ws.numel(), _stream(dev)), "frn_pruned_loss")
    return scores, grad


class _PrunedLossFn(torch.autograd.Function):
    """Autograd node standing where TF's GradientTape + _RNNTLossGrad
    (__init__.py:154-162) stand in the reference."""

    @staticmethod
    def forward(ctx, logits, symbols, ranges, termination_symbol, boundary, rnnt_type, delay_penalty):
        scores, _ = pruned_loss_fwd_bwd(logits, symbols, ranges, termination_symbol, boundary, rnnt_type,
                                        delay_penalty, None, want_logits_grad=False)
        ctx.save_for_backward(logits, symbols, ranges, boundary)
        ctx.args = (termination_symbol, rnnt_type, delay_penalty)
        return scores

    @staticmethod
    def backward(ctx, g):
        logits, symbols, ranges, boundary = ctx.saved_tensors
        term, rnnt_type, dp = ctx.args
        _, grad = pruned_loss_fwd_bwd(logits, symbols, ranges, term, boundary, rnnt_type, dp,
                                      g.contiguous(), want_logits_grad=True)
        return grad, None, None, None, None, None, None


@_on_device
def rnnt_loss_pruned(logits: Tensor, symbols: Tensor, ranges: Tensor, termination_symbol: int,
                     boundary: Tensor = None, rnnt_type: str = "regular", delay_penalty: float = 0.0,
                     reduction: Optional[str] = "mean", calc_gradients: bool = False, group=None):
    """Reference: rnnt_loss.py:1022-1130 (returns only the loss, like the
    reference, whatever ``calc_gradients`` is)."""
    if reduction not in _lib.REDUCTIONS:
        raise ValueError(
            f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")
    io = _Io(logits)
    if isinstance(logits, torch.Tensor) and logits.is_cuda:
        lg = logits
    else:
        lg = io.dev_tensor(logits, torch.float32)
    sym_d = io.dev_tensor(symbols, torch.int32)
    rg = io.dev_tensor(ranges, torch.int32)
    B, T, R, C = lg.shape
    S = sym_d.shape[1]
    bd = _boundary(io, boundary, B, S, T)
    if lg.requires_grad and torch.is_grad_enabled():
        scores = _PrunedLossFn.apply(lg, sym_d, rg, int(termination_symbol), bd, rnnt_type,
                                     float(delay_penalty))
        return _reduce_autograd(scores, reduction, group)
    scores, _ = pruned_loss_fwd_bwd(lg, sym_d, rg, termination_symbol, bd, rnnt_type, delay_penalty, None,
                                    want_logits_grad=False)
    return io.out(_reduce(scores, reduction, group))


# ---------------------------------------------------------------- (f1) full joiner
@_on_device
def get_rnnt_logprobs_joint(logits: Tensor, symbols: Tensor, termination_symbol: int,
                            boundary: Optional[Tensor] = None, rnnt_type: str = "regular"):
    """Reference: rnnt_loss.py:340-452.  The full joiner is the pruned case with
    the identity band ranges[b,t,i] = i."""
    io = _Io(logits)
    lg = logits if isinstance(logits, torch.Tensor) and logits.is_cuda else io.dev_tensor(logits, torch.float32)
    B, T, S1, C = lg.shape
    ranges = torch.arange(S1, dtype=torch.int32, device=io.dev).expand(B, T, S1).contiguous()
    sym_d = io.dev_tensor(symbols, torch.int32)
    bd = _boundary(io, boundary, B, S1 - 1, T)
    px, py = get_rnnt_logprobs_pruned(lg, sym_d, ranges, termination_symbol, bd, rnnt_type)
    return io.out(px), io.out(py)


def _joint_loss_call(lg, sym_d, bd, termination_symbol, rnnt_type, delay_penalty, scores_grad, want_grad):
    """frn_joint_loss on device tensors -> (scores, logits_grad or None)."""
    B, T, S1, C = lg.shape
    S = S1 - 1
    scores = torch.empty(B, dtype=torch.float32, device=lg.device)
    grad = torch.empty_like(lg) if want_grad else None
    ws = _workspace(lib.frn_joint_loss_workspace_bytes(B, S, T), lg.device)
    dp = float(delay_penalty) if delay_penalty > 0.0 else 0.0
    check(lib.frn_joint_loss(_ptr(lg), _logits_dtype(lg), _ptr(sym_d), _ptr(bd), B, S, T, C,
                             int(termination_symbol), _rnnt_type(rnnt_type), dp, _ptr(scores_grad), _ptr(scores),
                             _ptr(grad), _ptr(ws), ws.numel(), _stream(lg.device)), "frn_joint_loss")
    return scores, grad


class _JointLossFn(torch.autograd.Function):
    """scores(logits) of the unpruned loss with its logits gradient (what TF autodiff derives through
    rnnt_loss.py:340-551 + _RNNTLossGrad)."""

    @staticmethod
    def forward(ctx, logits, symbols, termination_symbol, boundary, rnnt_type, delay_penalty):
        scores, _ = _joint_loss_call(logits, symbols, boundary, termination_symbol, rnnt_type, delay_penalty, None, False)
        ctx.save_for_backward(logits, symbols, boundary)
        ctx.args = (termination_symbol, rnnt_type, delay_penalty)
        return scores

    @staticmethod
    def backward(ctx, g):
        logits, symbols, boundary = ctx.saved_tensors
        term, rnnt_type, dp = ctx.args
        _, grad = _joint_loss_call(logits, symbols, boundary, term, rnnt_type, dp, g.contiguous().float(), True)
        return grad, None, None, None, None, None


@_on_device
def rnnt_loss(logits: Tensor, symbols: Tensor, termination_symbol: int, boundary: Optional[Tensor] = None,
              rnnt_type: str = "regular", delay_penalty: float = 0.0, reduction: Optional[str] = "mean",
              calc_gradients: bool = False, group=None):
    """Reference: rnnt_loss.py:454-551."""
    if reduction not in _lib.REDUCTIONS:
        raise ValueError(
            f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")
    io = _Io(logits)
    lg = logits if isinstance(logits, torch.Tensor) and logits.is_cuda else io.dev_tensor(logits, torch.float32)
    lg = lg.contiguous()
    sym_d = io.dev_tensor(symbols, torch.int32)
    B, T, S1, C = lg.shape
    S = S1 - 1
    _rnnt_type(rnnt_type)
    bd = _boundary(io, boundary, B, S, T)
    if lg.requires_grad and torch.is_grad_enabled():
        scores = _JointLossFn.apply(lg, sym_d, int(termination_symbol), bd, rnnt_type, float(delay_penalty))
        return _reduce_autograd(scores, reduction, group)
    scores, _ = _joint_loss_call(lg, sym_d, bd, termination_symbol, rnnt_type, delay_penalty, None, False)
    return io.out(_reduce(scores, reduction, group))
