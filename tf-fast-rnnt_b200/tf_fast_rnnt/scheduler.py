"""Batch scheduler for ragged batches: the pruned RNN-T pipeline run per length bucket.

The reference pads a batch to its longest utterance and lets every TensorFlow op stream the padding
(rnnt_loss.py works on [B, T, ...] / [B, S, ...] tensors throughout).  The lattice recursions here already
stop at each utterance's own (S_b, T_b); what padding still costs is bytes - ``do_rnnt_pruning``, the joiner
and the pruned log-softmax touch ``T_max * s_range * C`` elements per utterance whatever ``T_b`` is - and
kernel variants picked for the padded shape instead of the real one.  ``pruned_rnnt_pipeline`` cuts the batch
into at most ``max_buckets`` length buckets (``sharding.plan_buckets``: dynamic programme on the padded
frame count), runs the reference's four calls

    rnnt_loss_simple(calc_gradients=True) -> get_rnnt_prune_ranges -> do_rnnt_pruning -> joiner
    -> rnnt_loss_pruned

on each bucket trimmed to its own (S_max, T_max), and puts the per-utterance results back in batch order.
Every bucket picks its own kernel variants by shape (row scan or wavefront recursion, band recursion, vector
widths).  Results per utterance equal the unbucketed calls (an utterance never sees its neighbours);
autograd flows through every stage.
"""
from __future__ import annotations

import importlib
from typing import Callable, List, Optional

import numpy as np
import torch

from . import _lib
from .sharding import plan_buckets

# (the package re-exports the FUNCTION rnnt_loss under the submodule's name, as the reference's __init__ does)
_rl = importlib.import_module(__package__ + ".rnnt_loss")


class Bucket:
    """One length bucket of a batch: utterance indices (batch order) and the shapes it is trimmed to."""

    def __init__(self, idx: np.ndarray, S_max: int, T_max: int, device):
        self.idx_host = np.asarray(idx, dtype=np.int64)
        self.idx = torch.from_numpy(self.idx_host).to(device)
        self.S_max, self.T_max = int(S_max), int(T_max)

    def take(self, x: torch.Tensor, kind: str) -> torch.Tensor:
        """Sub-batch of `x` trimmed to the bucket: kind 'am' [B,T,C] -> [b,T_max,C]; 'lm' [B,S+1,C] ->
        [b,S_max+1,C]; 'symbols' [B,S] -> [b,S_max]; 'boundary' [B,4] -> [b,4]."""
        sub = x.index_select(0, self.idx)
        if kind == "am":
            return sub[:, :self.T_max].contiguous()
        if kind == "lm":
            return sub[:, :self.S_max + 1].contiguous()
        if kind == "symbols":
            return sub[:, :self.S_max].contiguous()
        if kind == "boundary":
            return sub.contiguous()
        raise ValueError(kind)


def make_buckets(boundary, s_range: int, vocab: int, max_buckets: int = 8, min_bucket: int = 4,
                 device=None) -> List[Bucket]:
    """Length buckets of a batch from its boundary rows [s_begin, t_begin, s_end, t_end] (host copy needed:
    the plan is made on the CPU).  s_end / t_end bound what a bucket is trimmed to."""
    bd = boundary.detach().cpu().numpy() if isinstance(boundary, torch.Tensor) else np.asarray(boundary)
    if device is None:
        device = boundary.device if isinstance(boundary, torch.Tensor) and boundary.is_cuda else _rl._device()
    return [Bucket(p["idx"], p["S_max"], p["T_max"], device)
            for p in plan_buckets(bd, s_range, vocab, max_buckets=max_buckets, min_bucket=min_bucket)]


def _additive_joiner(am, lm, ranges):
    """The joiner of the reference's tests (simple_rnnt_loss_test.py:120-125): one fused pass writes
    am_pruned, lm_pruned and their sum."""
    _, _, logits = _rl.do_rnnt_pruning_add_joiner(am, lm, ranges)
    return logits


@_rl._on_device
def pruned_rnnt_pipeline(lm: torch.Tensor, am: torch.Tensor, symbols: torch.Tensor, termination_symbol: int,
                         boundary: torch.Tensor, s_range: int, joiner: Optional[Callable] = None,
                         rnnt_type: str = "regular", delay_penalty: float = 0.0, reduction: Optional[str] = "sum",
                         max_buckets: int = 8, min_bucket: int = 4, lm_only_scale: float = 0.0,
                         am_only_scale: float = 0.0, group=None, return_ranges: bool = False):
    """The full pruned RNN-T step on a ragged batch, per length bucket.

    lm [B,S+1,C], am [B,T,C], symbols [B,S] int32, boundary [B,4] int32: CUDA tensors, padded like the
    reference's inputs.  ``joiner(am_pruned, lm_pruned) -> logits [b,T',R,C']`` is the user's joiner network
    (None: the additive joiner, fused with the pruning).  ``lm_only_scale`` / ``am_only_scale`` > 0 select
    rnnt_loss_smoothed for the first pass.  Returns ``(simple_loss, pruned_loss)`` with the reference's
    reductions ('none': per-utterance vectors in batch order); with ``group`` the sum / mean is completed
    across the ranks of a torch.distributed group.  max_buckets = 1 is the reference's schedule (one padded
    batch).  ``return_ranges``: also return the prune ranges [B,T,R] in batch order (frames beyond a bucket's
    T_max are 0)."""
    if reduction not in _lib.REDUCTIONS:
        raise ValueError(f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")
    if not (isinstance(am, torch.Tensor) and am.is_cuda):
        raise TypeError("pruned_rnnt_pipeline works on CUDA tensors")
    B, T, C = am.shape
    bd = boundary.to(torch.int32)
    buckets = make_buckets(bd, s_range, C, max_buckets, min_bucket, am.device)
    smoothed = lm_only_scale > 0.0 or am_only_scale > 0.0
    simple_scores = torch.zeros(B, dtype=torch.float32, device=am.device)
    pruned_scores = torch.zeros(B, dtype=torch.float32, device=am.device)
    all_ranges = None
    for bk in buckets:
        lm_k, am_k = bk.take(lm, "lm"), bk.take(am, "am")
        sym_k, bd_k = bk.take(symbols.to(torch.int32), "symbols"), bk.take(bd, "boundary")
        if smoothed:
            loss_k, (gx, gy) = _rl.rnnt_loss_smoothed(lm_k, am_k, sym_k, termination_symbol, lm_only_scale,
                                                      am_only_scale, bd_k, rnnt_type, delay_penalty, "none", True)
        else:
            loss_k, (gx, gy) = _rl.rnnt_loss_simple(lm_k, am_k, sym_k, termination_symbol, bd_k, rnnt_type,
                                                    delay_penalty, "none", True)
        ranges = _rl.get_rnnt_prune_ranges(gx, gy, bd_k, s_range)
        if return_ranges:
            if all_ranges is None:
                all_ranges = torch.zeros((B, T, ranges.shape[2]), dtype=torch.int32, device=am.device)
            all_ranges[bk.idx, :bk.T_max] = ranges
        if joiner is None:
            logits = _additive_joiner(am_k, lm_k, ranges)
        else:
            am_p, lm_p = _rl.do_rnnt_pruning(am_k, lm_k, ranges)
            logits = joiner(am_p, lm_p)
        ploss_k = _rl.rnnt_loss_pruned(logits, sym_k, ranges, termination_symbol, bd_k, rnnt_type, delay_penalty,
                                       "none")
        # per-utterance losses back in batch order (index_put keeps autograd)
        simple_scores = simple_scores.index_put((bk.idx,), -loss_k)
        pruned_scores = pruned_scores.index_put((bk.idx,), -ploss_k)
    if torch.is_grad_enabled() and (simple_scores.requires_grad or pruned_scores.requires_grad):
        out = (_rl._reduce_autograd(simple_scores, reduction, group),
               _rl._reduce_autograd(pruned_scores, reduction, group))
    else:
        out = (_rl._reduce(simple_scores, reduction, group), _rl._reduce(pruned_scores, reduction, group))
    return out + (all_ranges,) if return_ranges else out
