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


def plan_buckets(boundary: np.ndarray, s_range: int, vocab: int, max_buckets: int = 4,
                 min_bucket: int = 4):
    """Batch scheduler for ragged shards: split a shard into at most
    ``max_buckets`` length buckets so that the bandwidth-bound kernels do not stream
    padding.  The lattice recursions already stop at each utterance's own (S_b, T_b); what
    padding costs is bytes: ``do_rnnt_pruning``, the joiner and the pruned log-softmax touch
    ``T_max * s_range * vocab`` elements per utterance whatever T_b is.

    Utterances are sorted by T_b and cut by dynamic programming into contiguous groups that
    minimise  sum_g |g| * T_max(g)  (the padded frame count, proportional to the bytes of
    those kernels), subject to ``len(g) >= min_bucket`` (each bucket is one more set of kernel
    launches; tiny buckets cannot fill the GPU).  Returns a list of dicts
    ``{"idx", "S_max", "T_max", "padded_frames", "bytes"}`` ordered by decreasing T_max;
    ``bytes`` estimates the pruned-path traffic of the bucket (SURVEY.md §8d:
    20 bytes per [t, i, c] element for pruning + joiner + log-softmax + logits gradient)."""
    b = np.asarray(boundary, dtype=np.int64)
    n = len(b)
    if n == 0:
        return []
    order = np.argsort(-b[:, 3], kind="stable")             # decreasing T_b: a group's T_max is its first element
    T_sorted = b[order, 3]
    k_max = max(1, min(max_buckets, n // max(min_bucket, 1) or 1))
    INF = float("inf")
    # cost[k][j] = minimal padded frames covering the first j utterances with k groups
    cost = np.full((k_max + 1, n + 1), INF)
    back = np.zeros((k_max + 1, n + 1), dtype=np.int64)
    cost[0, 0] = 0.0
    for k in range(1, k_max + 1):
        for j in range(1, n + 1):
            for i in range(0, j):
                if j - i < min_bucket and not (k == 1 and i == 0):
                    continue
                if cost[k - 1, i] == INF:
                    continue
                c = cost[k - 1, i] + (j - i) * T_sorted[i]
                if c < cost[k, j]:
                    cost[k, j] = c
                    back[k, j] = i
    k_best = int(np.argmin(cost[1:, n])) + 1
    cuts, j = [], n
    for k in range(k_best, 0, -1):
        i = int(back[k, j])
        cuts.append((i, j))
        j = i
    out = []
    for i, j in reversed(cuts):
        idx = order[i:j]
        t_max, s_max = int(b[idx, 3].max()), int(b[idx, 2].max())
        frames = int(len(idx) * t_max)
        out.append({"idx": np.sort(idx), "S_max": s_max, "T_max": t_max, "padded_frames": frames,
                    "bytes": int(20 * frames * s_range * vocab)})
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
