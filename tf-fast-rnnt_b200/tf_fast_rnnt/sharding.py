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
