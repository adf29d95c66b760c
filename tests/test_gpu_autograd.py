"""Autograd through the building blocks (the reference composes them under TF autodiff: rnnt_loss_pruned is
get_rnnt_logprobs_pruned followed by mutual_information_recursion, rnnt_loss.py:1088-1119, with the gradient of
FastRNNTLoss registered at __init__.py:154-162), the length-bucket scheduler, and the small entry points added
in round 2."""
import numpy as np
import pytest

from oracle import rnnt_oracle as orc
from tests.helpers import GRAD_ATOL, GRAD_RTOL, LOSS_RTOL, assert_close, make_inputs

pytestmark = pytest.mark.gpu


def _torch_logprobs64(lm, am, sym, term, bd, rnnt_type):
    """float64 torch restatement of rnnt_loss.py:175-221 (test reference for arbitrary cotangents)."""
    import torch
    B, T, C = am.shape
    S = lm.shape[1] - 1
    norm = torch.logsumexp(lm[:, :, None, :] + am[:, None, :, :], dim=3)            # [B,S+1,T]
    idx = sym.long()[:, :, None].expand(B, S, T)
    px = am.transpose(1, 2).gather(1, idx) + lm[:, :S].gather(2, sym.long()[:, :, None]) - norm[:, :S]
    py = am[:, :, term][:, None, :] + lm[:, :, term][:, :, None] - norm
    if rnnt_type == "regular":
        px = torch.cat([px, torch.full((B, S, 1), float("-inf"), dtype=px.dtype)], dim=2)
        mask = torch.zeros(B, 1, T + 1, dtype=torch.bool)
        mask[torch.arange(B), 0, bd[:, 3].long()] = True
        px = torch.where(mask, torch.full_like(px, float("-inf")), px)
    elif rnnt_type == "constrained":
        px = px + py[:, 1:, :]
    return px, py


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
def test_simple_logprobs_backward_with_arbitrary_cotangents(rnnt_type):
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C = 2, 37, 11, 20
    am, lm, sym, term, bd = make_inputs(41, B, T, S, C, ragged=True)
    rng = np.random.default_rng(7)
    T1 = T + 1 if rnnt_type == "regular" else T
    cx = rng.standard_normal((B, S, T1)).astype(np.float32)
    cy = rng.standard_normal((B, S + 1, T)).astype(np.float32)
    lm_d = torch.from_numpy(lm).cuda().requires_grad_(True)
    am_d = torch.from_numpy(am).cuda().requires_grad_(True)
    px, py = frn.get_rnnt_logprobs(lm_d, am_d, torch.from_numpy(sym).cuda(), term, rnnt_type, torch.from_numpy(bd).cuda())
    fin = torch.isfinite(px)
    obj = (torch.where(fin, px, torch.zeros_like(px)) * torch.from_numpy(cx).cuda()).sum() + (py * torch.from_numpy(cy).cuda()).sum()
    obj.backward()
    lm64 = torch.from_numpy(lm).double().requires_grad_(True)
    am64 = torch.from_numpy(am).double().requires_grad_(True)
    rx, ry = _torch_logprobs64(lm64, am64, torch.from_numpy(sym), term, torch.from_numpy(bd), rnnt_type)
    assert_close(px.detach().cpu().numpy(), rx.detach().numpy(), 1e-5, 1e-5, "px")
    rfin = torch.isfinite(rx)
    robj = (torch.where(rfin, rx, torch.zeros_like(rx)) * torch.from_numpy(cx).double()).sum() + (ry * torch.from_numpy(cy).double()).sum()
    robj.backward()
    assert_close(am_d.grad.cpu().numpy(), am64.grad.numpy(), 2e-4, 2e-5, "am grad")
    assert_close(lm_d.grad.cpu().numpy(), lm64.grad.numpy(), 2e-4, 2e-5, "lm grad")


@pytest.mark.parametrize("rnnt_type", ["regular", "modified"])
def test_logprobs_composed_with_recursion_equals_fused_loss(rnnt_type):
    """get_rnnt_logprobs -> mutual_information_recursion under autograd == rnnt_loss_simple (fused) gradients."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C = 3, 61, 17, 24
    am, lm, sym, term, bd = make_inputs(5, B, T, S, C, ragged=True)
    sym_d, bd_d = torch.from_numpy(sym).cuda(), torch.from_numpy(bd).cuda()
    w = torch.tensor([1.0, -0.5, 2.0], device="cuda")

    def grads(fused):
        lm_d = torch.from_numpy(lm).cuda().requires_grad_(True)
        am_d = torch.from_numpy(am).cuda().requires_grad_(True)
        if fused:
            loss = frn.rnnt_loss_simple(lm_d, am_d, sym_d, term, bd_d, rnnt_type, 0.0, "none")
        else:
            px, py = frn.get_rnnt_logprobs(lm_d, am_d, sym_d, term, rnnt_type, bd_d)
            loss = -frn.mutual_information_recursion(px, py, bd_d)
        (loss * w).sum().backward()
        return loss.detach().cpu().numpy(), am_d.grad.cpu().numpy(), lm_d.grad.cpu().numpy()

    l0, a0, m0 = grads(True)
    l1, a1, m1 = grads(False)
    assert_close(l1, l0, LOSS_RTOL, 0, "loss")
    assert_close(a1, a0, 2e-4, 2e-6, "am grad")
    assert_close(m1, m0, 2e-4, 2e-6, "lm grad")
    # and against the float64 oracle
    o_loss, (o_gx, o_gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True, dtype=np.float64)
    assert_close(l1, o_loss, LOSS_RTOL, 0, "loss vs oracle")


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
def test_pruned_logprobs_composed_with_recursion_equals_fused_loss(rnnt_type):
    """get_rnnt_logprobs_pruned -> mutual_information_recursion under autograd == rnnt_loss_pruned gradients
    (the reference's own composition, rnnt_loss.py:1088-1119)."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C, R = 3, 47, 13, 18, 4
    am, lm, sym, term, bd = make_inputs(9, B, T, S, C, ragged=True)
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
    logits = (am_p + lm_p).astype(np.float32)
    sym_d, bd_d, rg_d = torch.from_numpy(sym).cuda(), torch.from_numpy(bd).cuda(), torch.from_numpy(ranges).cuda()
    w = torch.tensor([1.0, 0.25, -3.0], device="cuda")

    def grads(fused):
        lg = torch.from_numpy(logits).cuda().requires_grad_(True)
        if fused:
            loss = frn.rnnt_loss_pruned(lg, sym_d, rg_d, term, bd_d, rnnt_type, 0.0, "none")
        else:
            px, py = frn.get_rnnt_logprobs_pruned(lg, sym_d, rg_d, term, bd_d, rnnt_type)
            loss = -frn.mutual_information_recursion(px, py, bd_d)
        (loss * w).sum().backward()
        return loss.detach().cpu().numpy(), lg.grad.cpu().numpy()

    l0, g0 = grads(True)
    l1, g1 = grads(False)
    assert_close(l1, l0, LOSS_RTOL, 0, "loss")
    assert_close(g1, g0, 2e-4, 2e-6, "logits grad")


def test_broadcast_am_pruned_and_joiner_without_am_output():
    """frn_broadcast_am_pruned (copy-engine am half of do_rnnt_pruning) + frn_do_pruning_add_joiner(am_pruned = NULL)
    == frn_do_pruning_add_joiner, bit for bit."""
    import torch
    import tf_fast_rnnt as frn
    lib, chk = frn._lib.lib, frn._lib.check
    B, T, S, C, R = 3, 70, 20, 64, 5
    rng = np.random.default_rng(2)
    am = torch.from_numpy(rng.standard_normal((B, T, C), dtype=np.float32)).cuda()
    lm = torch.from_numpy(rng.standard_normal((B, S + 1, C), dtype=np.float32)).cuda()
    r0 = rng.integers(0, S + 2 - R, (B, T, 1))
    ranges = torch.from_numpy((r0 + np.arange(R)).astype(np.int32)).cuda()
    ref = [torch.empty(B, T, R, C, device="cuda") for _ in range(3)]
    st = torch.cuda.current_stream().cuda_stream
    chk(lib.frn_do_pruning_add_joiner(am.data_ptr(), lm.data_ptr(), ranges.data_ptr(), B, S, T, R, C, ref[0].data_ptr(),
                                      ref[1].data_ptr(), ref[2].data_ptr(), st), "ref")
    for ctas in (1, 3, 20, 1000):
        out = [torch.full((B, T, R, C), float("nan"), device="cuda") for _ in range(3)]
        chk(lib.frn_broadcast_am_pruned(am.data_ptr(), B, T, R, C, out[0].data_ptr(), ctas, st), "broadcast")
        chk(lib.frn_do_pruning_add_joiner(am.data_ptr(), lm.data_ptr(), ranges.data_ptr(), B, S, T, R, C, 0,
                                          out[1].data_ptr(), out[2].data_ptr(), st), "lm+logits")
        for a, b in zip(out, ref):
            assert torch.equal(a, b)
    assert lib.frn_broadcast_am_pruned(am.data_ptr(), B, T, R, 63, ref[0].data_ptr(), 4, st) == -4   # C % 4 != 0


@pytest.mark.parametrize("shape", [(3, 70, 20, 64), (2, 150, 100, 128), (2, 600, 40, 36)])
def test_simple_loss_with_the_am_broadcast_beside_the_recursion(shape):
    """frn_simple_loss_bcast (the am half of do_rnnt_pruning forked onto a second stream behind the normaliser, joined
    after the read-out) == frn_simple_loss followed by frn_broadcast_am_pruned, bit for bit - on the fused arc-plane
    path (first two shapes) and on the sequential fall-back (third: long lattice on the row-scan kernel); also
    captured in a CUDA graph."""
    import torch
    import tf_fast_rnnt as frn
    lib, chk = frn._lib.lib, frn._lib.check
    B, T, S, C = shape
    R = 5
    am_h, lm_h, sym_h, term, bd_h = make_inputs(5, B, T, S, C, ragged=True)
    am, lm = torch.from_numpy(am_h).cuda(), torch.from_numpy(lm_h).cuda()
    sym, bd = torch.from_numpy(sym_h).cuda(), torch.from_numpy(bd_h).cuda()
    main, side = torch.cuda.current_stream(), torch.cuda.Stream()
    fork, join = torch.cuda.Event(), torch.cuda.Event()
    fork.record(); join.record()
    torch.cuda.synchronize()
    ws = torch.empty(int(lib.frn_simple_loss_workspace_bytes(B, S, T, C)), dtype=torch.uint8, device="cuda")

    def outputs():
        return (torch.full((B,), float("nan"), device="cuda"), torch.full((B, S, T + 1), float("nan"), device="cuda"),
                torch.full((B, S + 1, T), float("nan"), device="cuda"), torch.full((B, T, R, C), float("nan"), device="cuda"))

    ref = outputs()
    chk(lib.frn_simple_loss(lm.data_ptr(), am.data_ptr(), sym.data_ptr(), bd.data_ptr(), B, S, T, C, term, 0, 0, 0.0, 0.0,
                            0.1, 1, ref[0].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr(), ws.data_ptr(), ws.numel(),
                            main.cuda_stream), "simple_loss")
    chk(lib.frn_broadcast_am_pruned(am.data_ptr(), B, T, R, C, ref[3].data_ptr(), 8, main.cuda_stream), "broadcast")

    def fused(out):
        chk(lib.frn_simple_loss_bcast(lm.data_ptr(), am.data_ptr(), sym.data_ptr(), bd.data_ptr(), B, S, T, C, term, 0, 0,
                                      0.0, 0.0, 0.1, 1, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), R,
                                      out[3].data_ptr(), 8, side.cuda_stream, fork.cuda_event, join.cuda_event,
                                      ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
            "simple_loss_bcast")

    out = outputs()
    fused(out)
    torch.cuda.synchronize()
    for a, b in zip(out, ref):
        assert torch.equal(a, b)
    out2 = outputs()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fused(out2)
    g.replay()
    torch.cuda.synchronize()
    for a, b in zip(out2, ref):
        assert torch.equal(a, b)


def test_pruning_backward_with_arbitrary_index_patterns():
    """frn_do_pruning_bwd's lm side is a scatter-add over ranges: any index pattern, also non-consecutive and
    repeated indices (hand-made ranges); wide bands take the per-element path."""
    import torch
    import tf_fast_rnnt as frn
    rng = np.random.default_rng(4)
    for (B, T, S, C, R) in [(2, 33, 9, 12, 4), (2, 17, 40, 8, 37)]:
        ranges = rng.integers(0, S + 1, (B, T, R)).astype(np.int32)
        ga = rng.standard_normal((B, T, R, C)).astype(np.float32)
        gl = rng.standard_normal((B, T, R, C)).astype(np.float32)
        am_g, lm_g = frn.do_rnnt_pruning_backward(ga, gl, ranges, S)
        want = np.zeros((B, S + 1, C), np.float64)
        for b in range(B):
            np.add.at(want[b], ranges[b].reshape(-1), gl[b].reshape(-1, C).astype(np.float64))
        assert_close(lm_g, want, 1e-5, 1e-5, "lm grad")
        assert_close(am_g, ga.astype(np.float64).sum(2), 1e-5, 1e-5, "am grad")


def test_pruned_loss_beyond_1024_rows_and_with_the_small_workspace():
    """The band recursion does not depend on S: S + 1 > 1024 works for the pruned loss (the dense wavefront does
    not hold that many rows), and it needs only frn_pruned_loss_min_workspace_bytes."""
    import torch
    import tf_fast_rnnt as frn
    lib = frn._lib.lib
    B, T, S, C, R = 1, 1300, 1100, 8, 4
    rng = np.random.default_rng(3)
    logits = rng.standard_normal((B, T, R, C)).astype(np.float32)
    sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
    bd = np.array([[0, 0, S, T]], np.int32)
    # a monotone band that reaches S at the last frame
    r0 = np.minimum(np.arange(T) * (S - R + 1) // (T - 1), S - R + 1).astype(np.int32)
    ranges = (r0[None, :, None] + np.arange(R)[None, None, :]).astype(np.int32)
    assert lib.frn_pruned_loss_min_workspace_bytes(B, S, T, R, 0.0) < (8 << 20)
    loss = frn.rnnt_loss_pruned(logits, sym, ranges, C - 1, bd, "regular", 0.0, "none")
    want = orc.rnnt_loss_pruned(logits, sym, ranges, C - 1, bd, "regular", 0.0, "none", dtype=np.float64)
    assert_close(loss, want, LOSS_RTOL, 0, "pruned loss at S = 1100")
    # c2 shape: the band path's workspace is a few MB, the dense planes ~100 MB
    assert lib.frn_pruned_loss_min_workspace_bytes(32, 100, 500, 5, 0.0) < (16 << 20)
    assert lib.frn_pruned_loss_workspace_bytes(32, 100, 500, 5) > (64 << 20)


def test_simple_loss_beyond_1024_rows_runs_the_row_scan():
    import tf_fast_rnnt as frn
    B, T, S, C = 1, 1500, 1030, 8
    am, lm, sym, term, bd = make_inputs(8, B, T, S, C, ragged=False)
    loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    o_loss, (o_gx, o_gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True, dtype=np.float64)
    assert_close(loss, o_loss, LOSS_RTOL, 0, "loss")
    assert_close(gy, o_gy, 4 * GRAD_RTOL, GRAD_ATOL, "py_grad")


@pytest.mark.parametrize("rnnt_type", ["regular", "modified"])
@pytest.mark.parametrize("smoothed", [False, True])
def test_bucketed_pipeline_equals_one_padded_batch(rnnt_type, smoothed):
    """pruned_rnnt_pipeline (length buckets) against the same calls on the one padded batch and
    against the float64 oracle run on every utterance alone (no padding at all).  Buckets pick their own kernel
    variants by shape, so occupation counts differ in the last bits and a near-tie in the arg-max of
    get_rnnt_prune_ranges may fall the other way: the pruned loss is therefore checked against the oracle on
    the ranges each run actually used."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C, R = 12, 90, 30, 32, 4
    rng = np.random.default_rng(11)
    am, lm, sym, term, bd = make_inputs(13, B, T, S, C, ragged=True)
    bd[:, 3] = rng.integers(25, T + 1, B)
    bd[:, 2] = np.minimum(rng.integers(5, S + 1, B), bd[:, 3])
    sym_d, bd_d = torch.from_numpy(sym).cuda(), torch.from_numpy(bd).cuda()
    kw = dict(lm_only_scale=0.2, am_only_scale=0.0) if smoothed else {}
    w = torch.from_numpy(rng.standard_normal(B).astype(np.float32)).cuda()

    def run(max_buckets):
        lm_d = torch.from_numpy(lm).cuda().requires_grad_(True)
        am_d = torch.from_numpy(am).cuda().requires_grad_(True)
        sl, pl, rg = frn.pruned_rnnt_pipeline(lm_d, am_d, sym_d, term, bd_d, R, None, rnnt_type, 0.0, "none",
                                              max_buckets=max_buckets, min_bucket=2, return_ranges=True, **kw)
        ((sl + pl) * w).sum().backward()
        return (sl.detach().cpu().numpy(), pl.detach().cpu().numpy(), am_d.grad.cpu().numpy(), lm_d.grad.cpu().numpy(),
                rg.cpu().numpy())

    one = run(1)
    many = run(4)
    assert len(frn.make_buckets(bd, R, C, 4, 2)) > 1
    same_ranges = [np.array_equal(one[4][b, :bd[b, 3]], many[4][b, :bd[b, 3]]) for b in range(B)]
    assert sum(same_ranges) >= B - 2
    for res in (one, many):
        for b in range(B):
            Sb, Tb = int(bd[b, 2]), int(bd[b, 3])
            bd1 = np.array([[0, 0, Sb, Tb]], np.int32)
            rg = res[4][b:b + 1, :Tb]
            assert rg.max() <= Sb and rg.min() >= 0
            logits = am[b:b + 1, :Tb, None, :] + lm[b, rg[0]][None]
            o = orc.rnnt_loss_pruned(logits, sym[b:b + 1, :Sb], rg, term, bd1, rnnt_type, 0.0, "none", dtype=np.float64)
            assert_close(res[1][b:b + 1], o, LOSS_RTOL, 0, f"pruned loss of utterance {b}")
            if not smoothed:
                o = orc.rnnt_loss_simple(lm[b:b + 1, :Sb + 1], am[b:b + 1, :Tb], sym[b:b + 1, :Sb], term, bd1, rnnt_type,
                                         0.0, "none", False, dtype=np.float64)
                assert_close(res[0][b:b + 1], o, LOSS_RTOL, 0, f"simple loss of utterance {b}")
    if not smoothed:        # (the smoothed loss's unigram is batch-global: buckets change it by design)
        assert_close(many[0], one[0], LOSS_RTOL, 0, "simple loss")
        ok = np.array(same_ranges)
        assert_close(many[1][ok], one[1][ok], LOSS_RTOL, 0, "pruned loss")
        if ok.all():
            assert_close(many[3], one[3], 5e-4, 5e-5, "lm grad")
            assert_close(many[2], one[2], 5e-4, 5e-5, "am grad")


def test_sum_and_mean_over_a_process_group_on_the_autograd_path():
    """`group=` on the gradient path: value = global sum / mean, gradient at the global scale.
    One rank here (NCCL, world size 1); the two-rank arithmetic is covered on gloo in test_sharding_cpu.py."""
    import os
    import torch
    import torch.distributed as dist
    import tf_fast_rnnt as frn
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29571")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        B, T, S, C = 3, 40, 9, 16
        am, lm, sym, term, bd = make_inputs(2, B, T, S, C, ragged=True)
        for red in ("sum", "mean"):
            out = []
            for group in (None, dist.group.WORLD):
                lm_d = torch.from_numpy(lm).cuda().requires_grad_(True)
                am_d = torch.from_numpy(am).cuda().requires_grad_(True)
                loss = frn.rnnt_loss_simple(lm_d, am_d, sym, term, bd, "regular", 0.0, red, group=group)
                loss.backward()
                out.append((float(loss.detach()), am_d.grad.clone()))
            assert abs(out[0][0] - out[1][0]) <= 1e-6 * abs(out[0][0])
            assert torch.equal(out[0][1], out[1][1])
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("rnnt_type", ["regular", "constrained"])
def test_sharded_smoothed_loss_and_gradients_equal_the_unsharded_batch(rnnt_type):
    """SURVEY.md 8e / VERDICT r1 item 7: rnnt_loss_smoothed with am_only_scale > 0 on a batch sharded by utterance.
    Two shards on one GPU, the two all-reduces (unigram sums: C+1 floats, d loss / d unigram: C floats) done by
    hand: scores, occupation counts and am / lm gradients of every shard equal those of the whole batch."""
    import torch
    import tf_fast_rnnt as frn
    lib, chk = frn._lib.lib, frn._lib.check
    B, T, S, C = 6, 44, 12, 24
    lms, ams = 0.25, 0.2
    am, lm, sym, term, bd = make_inputs(31, B, T, S, C, ragged=True)
    rt = {"regular": 0, "constrained": 2}[rnnt_type]
    T1 = T + 1 if rnnt_type == "regular" else T
    g = np.random.default_rng(1).standard_normal(B).astype(np.float32)
    dev = torch.device("cuda", 0)
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    shards = [np.arange(0, 2), np.arange(2, 6)]                # unequal on purpose
    # whole batch (one rank)
    w_loss, (w_gx, w_gy) = frn.rnnt_loss_smoothed(lm, am, sym, term, lms, ams, bd, rnnt_type, 0.0, "none", True)
    w_am, w_lm = frn.smoothed_loss_backward(lm, am, sym, term, bd, w_gx, w_gy, -g, lms, ams, rnnt_type)
    # shards: phase A - local unigram sums, "all-reduce" = add
    sums = []
    for idx in shards:
        lm_d = cu(lm[idx])
        s_ = torch.empty(C + 1, device=dev)
        ws = torch.empty(max(int(lib.frn_simple_logprobs_workspace_bytes(len(idx), S, 1, C)), 256), dtype=torch.uint8, device=dev)
        chk(lib.frn_smoothed_unigram_sums(lm_d.data_ptr(), len(idx), S, C, s_.data_ptr(), ws.data_ptr(), ws.numel(), st), "sums")
        sums.append(s_)
    usums = sums[0] + sums[1]
    assert float(usums[C]) == B * (S + 1)
    # phase B - forward per shard with the global sums; backward phase 1 -> add du -> phase 2
    state = []
    for idx in shards:
        n = len(idx)
        t = dict(lm=cu(lm[idx]), am=cu(am[idx]), sym=cu(sym[idx]), bd=cu(bd[idx]), g=cu(-g[idx]),
                 scores=torch.empty(n, device=dev), gx=torch.empty(n, S, T1, device=dev), gy=torch.empty(n, S + 1, T, device=dev),
                 am_g=torch.empty(n, T, C, device=dev), lm_g=torch.empty(n, S + 1, C, device=dev), du=torch.empty(C, device=dev))
        ws = torch.empty(int(lib.frn_simple_loss_workspace_bytes(n, S, T, C)), dtype=torch.uint8, device=dev)
        chk(lib.frn_simple_loss_sharded(t["lm"].data_ptr(), t["am"].data_ptr(), t["sym"].data_ptr(), t["bd"].data_ptr(), n, S, T, C,
                                        term, rt, 1, lms, ams, usums.data_ptr(), 0.0, 1, t["scores"].data_ptr(),
                                        t["gx"].data_ptr(), t["gy"].data_ptr(), ws.data_ptr(), ws.numel(), st), "loss")
        t["ws"] = torch.empty(int(lib.frn_simple_loss_bwd_workspace_bytes(n, S, T, C)), dtype=torch.uint8, device=dev)
        state.append((idx, n, t))

    def bwd(t, n, phase):
        chk(lib.frn_smoothed_loss_bwd_sharded(t["lm"].data_ptr(), t["am"].data_ptr(), t["sym"].data_ptr(), t["bd"].data_ptr(),
                                              t["gx"].data_ptr(), t["gy"].data_ptr(), t["g"].data_ptr(), n, S, T, C, term, rt,
                                              lms, ams, usums.data_ptr(), t["du"].data_ptr(), phase, t["am_g"].data_ptr(),
                                              t["lm_g"].data_ptr(), t["ws"].data_ptr(), t["ws"].numel(), st), "bwd")
    for idx, n, t in state:
        bwd(t, n, 1)
    du = state[0][2]["du"] + state[1][2]["du"]
    for idx, n, t in state:
        t["du"].copy_(du)
        bwd(t, n, 2)
    for idx, n, t in state:
        assert_close(-t["scores"].cpu().numpy(), w_loss[idx], 2e-6, 0, "scores")
        assert_close(t["gx"].cpu().numpy(), w_gx[idx], 1e-5, 1e-7, "px_grad")
        assert_close(t["am_g"].cpu().numpy(), w_am[idx], 2e-5, 1e-6, "am grad")
        assert_close(t["lm_g"].cpu().numpy(), w_lm[idx], 2e-5, 1e-6, "lm grad")
    # the whole-batch numbers themselves are pinned by the float64 oracle elsewhere; here also directly:
    o_am, o_lm = orc.smoothed_am_lm_grad(lm, am, sym, term, bd, lms, ams, rnnt_type, 0.0, g, np.float64)
    for idx, n, t in state:
        assert_close(t["lm_g"].cpu().numpy(), o_lm[idx], 2 * GRAD_RTOL, 2e-5, "lm grad vs oracle")
        assert_close(t["am_g"].cpu().numpy(), o_am[idx], 2 * GRAD_RTOL, 2e-6, "am grad vs oracle")
    # and a shard evaluated WITHOUT the exchange differs (the coupling is real)
    alone = frn.rnnt_loss_smoothed(lm[:2], am[:2], sym[:2], term, lms, ams, bd[:2], rnnt_type, 0.0, "none")
    assert np.abs(alone - w_loss[:2]).max() > 1e-5 * np.abs(w_loss[:2]).max()


def test_in_library_allreduce_over_a_caller_owned_nccl_communicator():
    """frn_allreduce_sum binds the NCCL already loaded in the process; one-rank communicator made with ctypes."""
    import ctypes
    import torch
    import tf_fast_rnnt as frn
    lib = frn._lib.lib
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    try:
        nccl = ctypes.CDLL("libnccl.so.2", mode=ctypes.RTLD_GLOBAL)
    except OSError:
        pytest.skip("no libnccl.so.2 on the loader path")
    uid = ctypes.create_string_buffer(128)
    assert nccl.ncclGetUniqueId(uid) == 0

    class Uid(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_char * 128)]
    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, Uid, ctypes.c_int]
    u = Uid()
    ctypes.memmove(ctypes.byref(u), uid, 128)
    assert nccl.ncclCommInitRank(ctypes.byref(comm), 1, u, 0) == 0
    x = torch.arange(501, dtype=torch.float32, device="cuda")
    want = x.clone()
    frn._lib.check(lib.frn_allreduce_sum(x.data_ptr(), x.numel(), comm, torch.cuda.current_stream().cuda_stream), "allreduce")
    torch.cuda.synchronize()
    assert torch.equal(x, want)
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    nccl.ncclCommDestroy(comm)
