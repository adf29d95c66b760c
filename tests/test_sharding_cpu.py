"""Host-side multi-GPU logic on CPU: utterance sharding + the scalar all-reduce
of reduction='sum'/'mean', 2 ranks over gloo.  The per-rank compute is the oracle
here (no GPU); on the B200 the same code path calls the CUDA kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tf-fast-rnnt_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from tests.helpers import make_inputs  # noqa: E402


def test_partition_is_a_balanced_exact_cover():
    from tf_fast_rnnt.sharding import lattice_cells, partition_batch, shard_max_shapes
    rng = np.random.default_rng(0)
    B = 256
    bd = np.zeros((B, 4), np.int32)
    bd[:, 3] = rng.integers(200, 1501, B)                      # config c5: T 200-1500
    bd[:, 2] = np.minimum(rng.integers(20, 401, B), bd[:, 3])  # S 20-400
    for world in (1, 2, 4, 8):
        parts = partition_batch(bd, world)
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(B))
        loads = np.array([lattice_cells(bd[p]).sum() for p in parts], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.05, (world, loads)
        for p in parts:
            s, t = shard_max_shapes(bd, p)
            assert s == bd[p, 2].max() and t == bd[p, 3].max()
    assert partition_batch(bd, 2)[0].tolist() == partition_batch(bd, 2)[0].tolist()   # deterministic


def test_bucket_plan_covers_the_shard_and_cuts_padding():
    from tf_fast_rnnt.sharding import plan_buckets
    rng = np.random.default_rng(1)
    B = 64
    bd = np.zeros((B, 4), np.int32)
    bd[:, 3] = rng.integers(200, 1501, B)                      # config c5: T 200-1500
    bd[:, 2] = np.minimum(rng.integers(20, 401, B), bd[:, 3])
    one = plan_buckets(bd, 5, 500, max_buckets=1)
    assert len(one) == 1 and one[0]["T_max"] == bd[:, 3].max() and one[0]["padded_frames"] == B * bd[:, 3].max()
    plan = plan_buckets(bd, 5, 500, max_buckets=4, min_bucket=4)
    assert 1 <= len(plan) <= 4
    allidx = np.concatenate([g["idx"] for g in plan])
    assert sorted(allidx.tolist()) == list(range(B))          # exact cover
    for g in plan:
        assert len(g["idx"]) >= 4
        assert g["T_max"] == bd[g["idx"], 3].max() and g["S_max"] == bd[g["idx"], 2].max()
        assert g["bytes"] == 20 * g["padded_frames"] * 5 * 500
    # buckets are contiguous in T: every utterance of a later bucket is no longer than any of an earlier one
    for a, b in zip(plan, plan[1:]):
        assert bd[a["idx"], 3].min() >= bd[b["idx"], 3].max()
    padded = sum(g["padded_frames"] for g in plan)
    assert padded < 0.75 * one[0]["padded_frames"]             # uniform T in [200,1500]: 4 buckets save > 25 %
    assert padded >= int(bd[:, 3].sum())                       # never below the unpadded frame count
    # optimality against brute force on a small case
    small = bd[:10]
    best = plan_buckets(small, 5, 500, max_buckets=2, min_bucket=1)
    Ts = np.sort(small[:, 3])[::-1]
    brute = min(i * Ts[0] + (10 - i) * Ts[i] for i in range(1, 10))
    assert sum(g["padded_frames"] for g in best) == min(brute, 10 * Ts[0])
    assert plan_buckets(np.zeros((0, 4), np.int32), 5, 500) == []


def test_bucket_plan_over_frames_and_labels():
    """With enough buckets the planner may first halve the shard at the median label length (label lengths are
    nearly independent of frame counts): exact cover, bucket size floor, and never a higher modelled cost than the
    frames-only plan."""
    from tf_fast_rnnt.sharding import _bucket_cost, _plan_by_t, plan_buckets
    rng = np.random.default_rng(7)
    B, R, C = 128, 5, 500
    bd = np.zeros((B, 4), np.int32)
    bd[:, 3] = rng.integers(200, 1501, B)
    bd[:, 2] = np.minimum(rng.integers(20, 401, B), bd[:, 3])
    plan = plan_buckets(bd, R, C, max_buckets=8, min_bucket=4)
    assert 1 <= len(plan) <= 8
    assert sorted(np.concatenate([g["idx"] for g in plan]).tolist()) == list(range(B))
    cost = 0.0
    for g in plan:
        assert len(g["idx"]) >= 4
        assert g["T_max"] == bd[g["idx"], 3].max() and g["S_max"] == bd[g["idx"], 2].max()
        cost += _bucket_cost(len(g["idx"]), g["T_max"], g["S_max"], R, C)
    _, frames_only = _plan_by_t(bd.astype(np.int64), np.arange(B), R, C, 8, 4)
    assert cost <= frames_only * (1 + 1e-12)
    # uniform, independent S_b: the two-way plan wins, and half of the buckets hold the short label sequences only
    med = np.median(bd[:, 2])
    assert sum(g["S_max"] <= med for g in plan) >= len(plan) // 2 - 1
    assert cost < 0.97 * frames_only
    # a shard whose label length follows its frame count gains nothing from the split: frames-only plan
    tied = bd.copy()
    tied[:, 2] = tied[:, 3] // 4
    plan2 = plan_buckets(tied, R, C, max_buckets=8, min_bucket=4)
    for a, b in zip(plan2, plan2[1:]):
        assert tied[a["idx"], 3].min() >= tied[b["idx"], 3].max()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, reduction, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import rnnt_oracle as orc
    from tf_fast_rnnt.sharding import allreduce_loss, partition_batch
    am, lm, sym, term, bd = make_inputs(11, 6, 30, 8, 12, ragged=True)
    mine = partition_batch(bd, world)[rank]
    losses = orc.rnnt_loss_simple(lm[mine], am[mine], sym[mine], term, bd[mine], "regular", 0.0, "none",
                                  dtype=np.float64)
    total = allreduce_loss(torch.tensor(float(losses.sum()), dtype=torch.float64), len(mine), reduction)
    if rank == 0:
        out.put(float(total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("reduction", ["sum", "mean"])
def test_two_rank_reduction_equals_unsharded(reduction):
    from oracle import rnnt_oracle as orc
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, reduction, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = out.get()
    am, lm, sym, term, bd = make_inputs(11, 6, 30, 8, 12, ragged=True)
    want = orc.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, reduction, dtype=np.float64)
    np.testing.assert_allclose(got, float(want), rtol=1e-12)


def _grad_worker(rank, world, port, reduction, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tf_fast_rnnt.rnnt_loss import _reduce_autograd
    sizes = [3, 5]                                     # unequal shards (partition_batch gives those)
    scores = torch.arange(sizes[rank], dtype=torch.float64) + 10.0 * rank + 1.0
    scores.requires_grad_(True)
    loss = _reduce_autograd(scores, reduction, dist.group.WORLD)
    loss.backward()
    out.put((rank, float(loss), scores.grad.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("reduction", ["sum", "mean"])
def test_group_reduction_on_the_autograd_path(reduction):
    """`group=` with gradients enabled: every rank reports the GLOBAL sum / mean and its utterances
    get the gradient of that global loss (-1 for 'sum', -1 / N_global for 'mean'), whatever the shard sizes."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, reduction, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(out.get() for _ in range(2))
    all_scores = np.concatenate([np.arange(3) + 1.0, np.arange(5) + 11.0])
    want = -all_scores.sum() if reduction == "sum" else -all_scores.mean()
    scale = -1.0 if reduction == "sum" else -1.0 / 8
    for rank, loss, grad in got:
        np.testing.assert_allclose(loss, want, rtol=1e-12)
        np.testing.assert_allclose(grad, [scale] * (3 if rank == 0 else 5), rtol=1e-12)


def _smoothed_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import rnnt_oracle as orc
    from tf_fast_rnnt.sharding import partition_batch
    am, lm, sym, term, bd = make_inputs(23, 6, 30, 8, 12, ragged=True)
    mine = partition_batch(bd, world)[rank]
    # the protocol of include/fast_rnnt_b200.h "batch sharded by utterance": local sums -> all-reduce -> loss
    sums, count = orc.smoothed_unigram_sums(lm[mine])
    buf = torch.from_numpy(np.concatenate([sums, [count]]))
    dist.all_reduce(buf)
    losses = orc.rnnt_loss_smoothed(lm[mine], am[mine], sym[mine], term, 0.25, 0.2, bd[mine], "regular", 0.0, "none",
                                    dtype=np.float64, unigram_sums=(buf[:-1].numpy(), float(buf[-1])))
    out.put((mine.tolist(), losses.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_smoothed_loss_with_am_only_scale_equals_unsharded():
    """rnnt_loss_smoothed's unigram is batch-global (rnnt_loss.py:1279-1280): with the C+1 unigram sums
    all-reduced before the am-only terms, two ranks reproduce the unsharded per-utterance losses
    (am_only_scale = 0.2) to 1e-6."""
    from oracle import rnnt_oracle as orc
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_smoothed_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    am, lm, sym, term, bd = make_inputs(23, 6, 30, 8, 12, ragged=True)
    want = orc.rnnt_loss_smoothed(lm, am, sym, term, 0.25, 0.2, bd, "regular", 0.0, "none", dtype=np.float64)
    got = np.zeros(6)
    for _ in range(2):
        idx, losses = out.get()
        got[idx] = losses
    np.testing.assert_allclose(got, want, rtol=1e-6)
    # and without the exchange the shards do NOT agree with the batch (the coupling is real)
    half = orc.rnnt_loss_smoothed(lm[:3], am[:3], sym[:3], term, 0.25, 0.2, bd[:3], "regular", 0.0, "none", dtype=np.float64)
    assert np.abs(half - want[:3]).max() > 1e-6 * np.abs(want[:3]).max()
