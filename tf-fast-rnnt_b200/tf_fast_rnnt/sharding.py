"""Utterance sharding across the GPUs of one box.

Every kernel on the path indexes its data by utterance, so the batch shards with
no data-path collective: each rank runs the whole pipeline on its own
utterances and only the scalar of ``reduction='sum'|'mean'`` crosses NVLink
(one all-reduce).  The reference has nothing distributed (replicas only).

``partition_batch`` deals utterances to ranks by greedy longest-processing-time
on the lattice size (S_b+1)(T_b+1), which balances the latency-bound recursion
and, to first order, the bandwidth-bound kernels (their cost is ~T_b).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def lattice_cells(boundary: np.ndarray) -> np.ndarray:
    b = np.asarray(boundary, dtype=np.int64)
    return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)


def partition_batch(boundary: np.ndarray, world_size: int) -> List[np.ndarray]:
    """Indices of the utterances each rank processes (length-bucketed, balanced).
    Deterministic; every utterance appears exactly once; ranks may differ in
    count by the LPT imbalance only."""
    cost = lattice_cells(boundary)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i))
        load[r] += int(cost[i])
    # within a rank keep utterances sorted by length so that padding is minimal
    return [np.asarray(sorted(p, key=lambda i: (-int(cost[i]), i)), dtype=np.int64) for p in parts]


def shard_max_shapes(boundary: np.ndarray, idx: Sequence[int]):
    """(S_max, T_max) a rank has to pad its shard to."""
    b = np.asarray(boundary)[np.asarray(idx, dtype=np.int64)]
    if len(b) == 0:
        return 0, 0
    return int(b[:, 2].max()), int(b[:, 3].max())


def _bucket_cost(n_utt: int, t_max: int, s_max: int, s_range: int, vocab: int) -> float:
    """Modelled cost of a bucket in byte-equivalents, padding included: the pruned path touches 20 bytes per
    [t, i, c] element (pruning + joiner + log-softmax + logits gradient); the simple loss works on
    every cell of the padded (S_max + 1) x T_max lattice - arcs, recursion planes, occupation counts, prune-range
    arg-max, and the contraction's 2 * 3 * vocab flops per cell.  The per-cell weight is calibrated on the kernel
    shares of the ragged c5 batch (profiles/r02ai_c5_launches_summary.txt: lattice kernels 56 %, pruned-path
    kernels 41 % of the step at S_max ~ 400, vocab 500: ~170 byte-equivalents per cell; the recursion is latency
    bound, so its cost per cell is well above its bytes)."""
    return float(n_utt) * t_max * (20.0 * s_range * vocab + (s_max + 1) * (130.0 + vocab / 12.0))


def _plan_by_t(b: np.ndarray, members: np.ndarray, s_range: int, vocab: int, max_buckets: int, min_bucket: int):
    """Cut `members` (indices into b), sorted by decreasing T_b, into at most max_buckets contiguous groups that
    minimise the modelled cost (dynamic programming)."""
    n = len(members)
    order = members[np.argsort(-b[members, 3], kind="stable")]      # a group's T_max is its first element
    T_sorted = b[order, 3]
    # S_max of order[i:j] for the cost: running maxima from every start (n is a few hundred at most)
    k_max = max(1, min(max_buckets, n // max(min_bucket, 1) or 1))
    INF = float("inf")
    cost = np.full((k_max + 1, n + 1), INF)
    back = np.zeros((k_max + 1, n + 1), dtype=np.int64)
    cost[0, 0] = 0.0
    S_sorted = b[order, 2]
    smax = np.zeros((n, n + 1), dtype=np.int64)
    for i in range(n):
        smax[i, i + 1:] = np.maximum.accumulate(S_sorted[i:])
    for k in range(1, k_max + 1):
        for j in range(1, n + 1):
            for i in range(0, j):
                if j - i < min_bucket and not (k == 1 and i == 0):
                    continue
                if cost[k - 1, i] == INF:
                    continue
                c = cost[k - 1, i] + _bucket_cost(j - i, int(T_sorted[i]), int(smax[i, j]), s_range, vocab)
                if c < cost[k, j]:
                    cost[k, j] = c
                    back[k, j] = i
    k_best = int(np.argmin(cost[1:, n])) + 1
    cuts, j = [], n
    for k in range(k_best, 0, -1):
        i = int(back[k, j])
        cuts.append((i, j))
        j = i
    return [order[i:j] for i, j in reversed(cuts)], float(cost[k_best, n])


def plan_buckets(boundary: np.ndarray, s_range: int, vocab: int, max_buckets: int = 8,
                 min_bucket: int = 4):
    """Batch scheduler for ragged shards: split a shard into at most
    ``max_buckets`` length buckets so that the kernels do not work on padding.  What padding
    costs: ``do_rnnt_pruning``, the joiner and the pruned log-softmax touch
    ``T_max * s_range * vocab`` elements per utterance whatever T_b is, and the normaliser and
    the read-out of the simple loss work on the padded ``(S_max + 1) x T_max`` lattice
    (``_bucket_cost``).

    Two plans are made and the cheaper (modelled) one is returned: (a) utterances sorted by T_b
    and cut by dynamic programming into contiguous groups; (b) the shard first halved at the
    median S_b - label lengths are nearly independent of frame counts, so sorting by T alone
    leaves S_max ~ the shard maximum in every bucket - and each half cut the same way with half
    the bucket budget.  Every group has ``len(g) >= min_bucket`` (each bucket is one more set of
    kernel launches; tiny buckets cannot fill the GPU).  Returns a list of dicts
    ``{"idx", "S_max", "T_max", "padded_frames", "bytes"}`` ordered by decreasing T_max;
    ``bytes`` estimates the pruned-path traffic of the bucket (SURVEY.md §8d:
    20 bytes per [t, i, c] element for pruning + joiner + log-softmax + logits gradient)."""
    b = np.asarray(boundary, dtype=np.int64)
    n = len(b)
    if n == 0:
        return []
    everyone = np.arange(n)
    groups, best = _plan_by_t(b, everyone, s_range, vocab, max_buckets, min_bucket)
    if max_buckets >= 4 and n >= 4 * max(min_bucket, 1):
        by_s = np.argsort(b[:, 2], kind="stable")
        lo, hi = by_s[: n // 2], by_s[n // 2:]
        g_lo, c_lo = _plan_by_t(b, lo, s_range, vocab, max_buckets // 2, min_bucket)
        g_hi, c_hi = _plan_by_t(b, hi, s_range, vocab, max_buckets - max_buckets // 2, min_bucket)
        if c_lo + c_hi < best:
            groups = g_lo + g_hi
    out = []
    for idx in groups:
        t_max, s_max = int(b[idx, 3].max()), int(b[idx, 2].max())
        frames = int(len(idx) * t_max)
        out.append({"idx": np.sort(idx), "S_max": s_max, "T_max": t_max, "padded_frames": frames,
                    "bytes": int(20 * frames * s_range * vocab)})
    out.sort(key=lambda g: (-g["T_max"], -g["S_max"]))
    return out


def allreduce_loss(local_sum, local_count: int, reduction: str, group=None):
    """Complete 'sum' / 'mean' over the ranks of ``group`` from per-rank partial
    sums of the per-utterance losses.  ``local_sum`` is a 0-d/1-element torch
    tensor on the device the process group's backend communicates from (CUDA for
    NCCL, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    if reduction not in ("sum", "mean"):
        raise ValueError(f"reduction should be ('mean' | 'sum') for a sharded batch, given {reduction}")
    buf = torch.stack([local_sum.reshape(()).to(torch.float64),
                       torch.tensor(float(local_count), dtype=torch.float64, device=local_sum.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, group=group)
    total, count = buf[0], buf[1]
    return (total / count if reduction == "mean" else total).to(local_sum.dtype)
