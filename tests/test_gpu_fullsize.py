"""BASELINE.json configs at full size on the GPU: comparison with the float64
oracle where the oracle finishes in seconds (c2, c3, sub-batches of c4/c5) and
size-independent properties everywhere else (occupation counts sum to one per
frame / per symbol, zero outside the boundary box, range monotonicity, pruned loss
>= unpruned loss, linearity of the logits gradient in the upstream gradient)."""
import numpy as np
import pytest

from oracle import rnnt_oracle as orc
from tests.helpers import GRAD_ATOL, GRAD_RTOL, LOSS_RTOL, assert_close, make_inputs

pytestmark = pytest.mark.gpu


def _occupation_properties(gx, gy, bd, regular=True):
    gx = gx.astype(np.float64); gy = gy.astype(np.float64)
    for b in range(gx.shape[0]):
        s_end, t_end = bd[b, 2], bd[b, 3]
        if regular:
            np.testing.assert_allclose(gy[b, :, :t_end].sum(axis=0), 1.0, rtol=2e-4)
            np.testing.assert_allclose(gx[b, :s_end, :].sum(axis=1), 1.0, rtol=2e-4)
        else:
            np.testing.assert_allclose(gx[b, :, :t_end].sum(axis=0) + gy[b, :, :t_end].sum(axis=0), 1.0, rtol=2e-4)
        assert gx[b, s_end:, :].sum() == 0 and gy[b, :, t_end:].sum() == 0
    assert not np.isnan(gx).any() and not np.isnan(gy).any()


def _range_properties(ranges, bd, R):
    r0 = ranges[:, :, 0].astype(np.int64)
    assert (ranges == r0[:, :, None] + np.arange(ranges.shape[2])).all()
    assert (np.diff(r0, axis=1) >= 0).all() and (np.diff(r0, axis=1) < R).all() and (r0[:, 0] == 0).all()
    for b in range(len(bd)):
        assert r0[b, bd[b, 3] - 1] == max(bd[b, 2] - R + 1, 0)


def test_c2_full_pipeline_against_oracle():
    """configs[1]: B=32 T=500 S=100 C=500 s_range=5 fp32 — every stage against the float64 oracle."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C, R = 32, 500, 100, 500, 5
    am, lm, sym, term, bd = make_inputs(1234, B, T, S, C, ragged=True)
    loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    o_loss, (o_gx, o_gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True, dtype=np.float64)
    assert_close(loss, o_loss, LOSS_RTOL, 0, "simple loss")
    # End to end the float32 log-probs themselves limit the occupation counts: every
    # arc score carries >= 3e-7 of representation error and a path has 600 arcs, so the
    # worst of the 1.6 M counts sits at ~1e-4 relative whatever the recursion does ...
    assert_close(gx, o_gx, GRAD_RTOL, GRAD_ATOL, "px_grad")
    assert_close(gy, o_gy, GRAD_RTOL, GRAD_ATOL, "py_grad")
    # ... while the recursion alone (same float32 px/py into the float64 oracle) meets
    # the 1e-4 of the north star with margin.
    px, py = frn.get_rnnt_logprobs(lm, am, sym, term, "regular", bd)
    ans, (mgx, mgy) = frn.mutual_information_recursion(px, py, bd, True)
    o_ans, (o_mgx, o_mgy) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    assert_close(ans, o_ans, LOSS_RTOL, 0, "recursion score")
    assert_close(mgx, o_mgx, 0.5 * GRAD_RTOL, GRAD_ATOL, "recursion px_grad")
    assert_close(mgy, o_mgy, 0.5 * GRAD_RTOL, GRAD_ATOL, "recursion py_grad")
    _occupation_properties(gx, gy, bd)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    assert np.array_equal(ranges, orc.get_rnnt_prune_ranges(gx, gy, bd, R))      # bit-exact
    _range_properties(ranges, bd, R)
    am_d, lm_d, rg_d = (torch.from_numpy(x).cuda() for x in (am, lm, ranges))
    am_p, lm_p = frn.do_rnnt_pruning(am_d, lm_d, rg_d)
    o_am_p, o_lm_p = orc.do_rnnt_pruning(am, lm, ranges)
    assert torch.equal(am_p.cpu(), torch.from_numpy(o_am_p)) and torch.equal(lm_p.cpu(), torch.from_numpy(o_lm_p))
    logits = am_p + lm_p
    w = np.linspace(0.5, 1.5, B).astype(np.float32)
    scores, grad = frn.pruned_loss_fwd_bwd(logits, sym, ranges, term, bd, "regular", 0.0, -w)
    lg = logits.cpu().numpy()
    o_grad, o_scores = orc.pruned_logits_grad(lg, sym, ranges, term, bd, "regular", 0.0, w, np.float64, True)
    assert_close(-scores.cpu().numpy(), -o_scores, LOSS_RTOL, 0, "pruned loss")
    assert_close(grad.cpu().numpy(), o_grad, GRAD_RTOL, 2e-6, "logits grad")
    assert (-scores.cpu().numpy() >= loss * (1 - 1e-6)).all()                  # pruning only removes paths


@pytest.mark.parametrize("rnnt_type", ["modified", "constrained"])
def test_c3_smoothed_types_delay_penalty(rnnt_type):
    """configs[2]: smoothed lm_only_scale=0.25 am_only_scale=0.0, delay_penalty=0.2, c2 shape."""
    import tf_fast_rnnt as frn
    B, T, S, C = 32, 500, 100, 500
    am, lm, sym, term, bd = make_inputs(77, B, T, S, C, ragged=True)
    loss, (gx, gy) = frn.rnnt_loss_smoothed(lm, am, sym, term, 0.25, 0.0, bd, rnnt_type, 0.2, "none", True)
    o_loss, (o_gx, o_gy) = orc.rnnt_loss_smoothed(lm, am, sym, term, 0.25, 0.0, bd, rnnt_type, 0.2, "none", True,
                                                  dtype=np.float64)
    # with delay_penalty the loss is a near-cancellation of +-thousands (penalties of up to
    # +-50 per arc): the tolerance is relative to the un-penalised loss, the natural scale
    scale = np.abs(orc.rnnt_loss_smoothed(lm, am, sym, term, 0.25, 0.0, bd, rnnt_type, 0.0, "none", dtype=np.float64))
    err = np.abs(loss.astype(np.float64) - o_loss)
    assert (err <= LOSS_RTOL * scale).all(), (err / scale).max()
    assert_close(gx, o_gx, GRAD_RTOL, GRAD_ATOL, "px_grad")
    assert_close(gy, o_gy, GRAD_RTOL, GRAD_ATOL, "py_grad")
    _occupation_properties(gx, gy, bd, regular=False)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, 5)
    assert np.array_equal(ranges, orc.get_rnnt_prune_ranges(gx, gy, bd, 5))
    assert (np.diff(ranges[:, :, 0], axis=1) >= 0).all() and (np.diff(ranges[:, :, 0], axis=1) < 2).all()


def test_c4_large_vocab_bf16():
    """configs[3]: B=16 T=1500 S=400 C=5000 s_range=5, bf16 joiner logits into an fp32 loss.
    Full batch: properties; first 2 utterances: float64 oracle ('upcast then reference math')."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C, R = 16, 1500, 400, 5000, 5
    rng = np.random.default_rng(4)
    am = torch.from_numpy(rng.standard_normal((B, T, C), dtype=np.float32)).cuda()
    lm = torch.from_numpy(rng.standard_normal((B, S + 1, C), dtype=np.float32)).cuda()
    sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
    bd = np.tile(np.array([0, 0, S, T], np.int32), (B, 1))
    bd[1] = [0, 0, 250, 1100]
    term = C - 1
    loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    gx_h, gy_h = gx.cpu().numpy(), gy.cpu().numpy()
    _occupation_properties(gx_h, gy_h, bd)
    nb = 2
    o_loss, (o_gx, o_gy) = orc.rnnt_loss_simple(lm[:nb].cpu().numpy(), am[:nb].cpu().numpy(), sym[:nb], term, bd[:nb],
                                                "regular", 0.0, "none", True, dtype=np.float64)
    assert_close(loss[:nb].cpu().numpy(), o_loss, LOSS_RTOL, 0, "c4 simple loss")
    assert_close(gx_h[:nb], o_gx, GRAD_RTOL, GRAD_ATOL, "c4 px_grad")
    assert_close(gy_h[:nb], o_gy, GRAD_RTOL, GRAD_ATOL, "c4 py_grad")
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    rg_h = ranges.cpu().numpy()
    assert np.array_equal(rg_h[:nb], orc.get_rnnt_prune_ranges(gx_h[:nb], gy_h[:nb], bd[:nb], R))
    _range_properties(rg_h, bd, R)
    logits = frn.pruned_add_joiner(am, lm, ranges, dtype=torch.bfloat16)
    scores, grad = frn.pruned_loss_fwd_bwd(logits, sym, ranges, term, bd, "regular", 0.0, None)
    assert grad.dtype == torch.bfloat16 and torch.isfinite(scores).all()
    up = logits[:nb].float().cpu().numpy()
    o_grad, o_scores = orc.pruned_logits_grad(up, sym[:nb], rg_h[:nb], term, bd[:nb], "regular", 0.0, -np.ones(nb),
                                              np.float64, True)
    assert_close(scores[:nb].cpu().numpy(), o_scores, LOSS_RTOL, 0, "c4 pruned scores (bf16 logits)")
    assert_close(grad[:nb].float().cpu().numpy(), o_grad, 1e-2, 1e-4, "c4 logits grad (bf16 rounding)")
    # each row of the logits gradient sums to zero (softmax Jacobian), checked on the full batch
    row_sums = grad.float().sum(dim=3)
    assert float(row_sums.abs().max()) < 2e-2


def test_c5_pruned_half_and_bucketed_path_at_ragged_shapes():
    """configs[4] shapes (T_b 200-1500, S_b 20-400, C=500, s_range=5), 40 utterances: prune ranges bit-exact,
    pruned loss and logits gradient against the float64 oracle on picked utterances, range properties on all -
    and the length-bucketed schedule bench.py runs for c5 (plan_buckets -> pipeline per bucket) against the one
    padded batch."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C, R = 40, 1500, 400, 500, 5
    rng = np.random.default_rng(55)
    bd = np.zeros((B, 4), np.int32)
    bd[:, 3] = rng.integers(200, T + 1, B)
    bd[:, 2] = np.minimum(rng.integers(20, S + 1, B), bd[:, 3])
    bd[0] = [0, 0, S, T]                      # the batch maxima are present
    am_h = rng.standard_normal((B, T, C), dtype=np.float32)
    lm_h = rng.standard_normal((B, S + 1, C), dtype=np.float32)
    sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
    am, lm = torch.from_numpy(am_h).cuda(), torch.from_numpy(lm_h).cuda()
    sym_d, bd_d = torch.from_numpy(sym).cuda(), torch.from_numpy(bd).cuda()
    term = C - 1
    loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym_d, term, bd_d, "regular", 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd_d, R)
    rg_h = ranges.cpu().numpy()
    _range_properties(rg_h, bd, R)
    logits = frn.pruned_add_joiner(am, lm, ranges)
    scores, grad = frn.pruned_loss_fwd_bwd(logits, sym_d, ranges, term, bd_d, "regular", 0.0, None)
    scores_h = scores.cpu().numpy()
    assert np.isfinite(scores_h).all()
    assert (-scores_h >= loss.cpu().numpy() * (1 - 1e-5)).all()      # pruning can only lose probability mass
    pick = [0, 7, int(np.argmin(bd[:, 3])), int(np.argmin(bd[:, 2]))]
    gx_h, gy_h = gx[pick].cpu().numpy(), gy[pick].cpu().numpy()
    assert np.array_equal(rg_h[pick], orc.get_rnnt_prune_ranges(gx_h, gy_h, bd[pick], R))
    o_grad, o_scores = orc.pruned_logits_grad(logits[pick].cpu().numpy(), sym[pick], rg_h[pick], term, bd[pick],
                                              "regular", 0.0, -np.ones(len(pick)), np.float64, True)
    assert_close(scores_h[pick], o_scores, LOSS_RTOL, 0, "c5 pruned scores")
    assert_close(grad[pick].cpu().numpy(), o_grad, GRAD_RTOL, 2e-6, "c5 logits grad")
    del logits, grad
    # the bucketed schedule: same per-utterance losses as the padded batch
    one = frn.pruned_rnnt_pipeline(lm, am, sym_d, term, bd_d, R, None, "regular", 0.0, "none", max_buckets=1,
                                   return_ranges=True)
    many = frn.pruned_rnnt_pipeline(lm, am, sym_d, term, bd_d, R, None, "regular", 0.0, "none", max_buckets=4,
                                    return_ranges=True)
    assert len(frn.make_buckets(bd_d, R, C, 4, 4)) > 1
    assert_close(one[0].cpu().numpy(), loss.cpu().numpy(), LOSS_RTOL, 0, "simple loss, padded pipeline vs direct call")
    assert_close(many[0].cpu().numpy(), one[0].cpu().numpy(), LOSS_RTOL, 0, "simple loss, bucketed vs padded")
    r1, r4 = one[2].cpu().numpy(), many[2].cpu().numpy()
    same = np.array([np.array_equal(r1[b, :bd[b, 3]], r4[b, :bd[b, 3]]) for b in range(B)])
    assert same.sum() >= B - 3                # a near-tie of the arg-max may fall the other way (other kernel variant)
    assert_close(many[1].cpu().numpy()[same], one[1].cpu().numpy()[same], LOSS_RTOL, 0, "pruned loss, bucketed vs padded")
    assert_close(one[1].cpu().numpy(), -scores_h, LOSS_RTOL, 0, "pruned loss, padded pipeline vs direct calls")


def test_c5_ragged_batch_sharded_sum():
    """configs[4]: ragged B=256 (T 200-1500, S 20-400), C=500, reduction=sum, sharded over 1/2/4/8 'ranks'
    (emulated on one GPU: the per-rank partial sums must add up to the unsharded sum)."""
    import torch
    import tf_fast_rnnt as frn
    from tf_fast_rnnt.sharding import partition_batch
    B, T, S, C = 256, 1500, 400, 500
    rng = np.random.default_rng(5)
    bd = np.zeros((B, 4), np.int32)
    bd[:, 3] = rng.integers(200, T + 1, B)
    bd[:, 2] = np.minimum(rng.integers(20, S + 1, B), bd[:, 3])
    am = torch.from_numpy(rng.standard_normal((B, T, C), dtype=np.float32)).cuda()
    lm = torch.from_numpy(rng.standard_normal((B, S + 1, C), dtype=np.float32)).cuda()
    sym = torch.from_numpy(rng.integers(0, C - 1, (B, S)).astype(np.int32)).cuda()
    bd_d = torch.from_numpy(bd).cuda()
    term = C - 1
    loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd_d, "regular", 0.0, "none", True)
    loss_h = loss.cpu().numpy().astype(np.float64)
    assert np.isfinite(loss_h).all()
    pick = [0, 17, 101, 255]
    o_loss = orc.rnnt_loss_simple(lm[pick].cpu().numpy(), am[pick].cpu().numpy(), sym[pick].cpu().numpy(), term,
                                  bd[pick], "regular", 0.0, "none", dtype=np.float64)
    assert_close(loss_h[pick], o_loss, LOSS_RTOL, 0, "c5 per-utterance loss")
    _occupation_properties(gx[pick].cpu().numpy(), gy[pick].cpu().numpy(), bd[pick])
    total = float(frn.rnnt_loss_simple(lm, am, sym, term, bd_d, "regular", 0.0, "sum"))
    np.testing.assert_allclose(total, loss_h.sum(), rtol=1e-5)
    for world in (2, 4, 8):
        parts = partition_batch(bd, world)
        partial = 0.0
        for idx in parts:
            i = torch.from_numpy(idx).cuda()
            s_max, t_max = int(bd[idx, 2].max()), int(bd[idx, 3].max())      # each rank pads to its own maxima
            partial += float(frn.rnnt_loss_simple(lm[i, :s_max + 1].contiguous(), am[i, :t_max].contiguous(),
                                                  sym[i, :s_max].contiguous(), term, bd_d[i], "regular", 0.0, "sum"))
        np.testing.assert_allclose(partial, total, rtol=2e-6)


def test_c2_am_lm_gradients_on_tensor_cores():
    """A9 at the c2 shape (B=4 of the 32: the float64 oracle holds [B,S+1,T] x C products): am / lm gradients of
    rnnt_loss_simple from the tcgen05 contraction kernel (simple_bwd_tc.cu) against the float64 oracle."""
    import tf_fast_rnnt as frn
    B, T, S, C = 4, 500, 100, 500
    am, lm, sym, term, bd = make_inputs(77, B, T, S, C, ragged=True)
    w = np.array([1.0, 0.5, -2.0, 1.5], np.float32)
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    am_g, lm_g = frn.simple_loss_backward(lm, am, sym, term, bd, gx, gy, -w, "regular")
    o_am, o_lm = orc.simple_am_lm_grad(lm, am, sym, term, bd, "regular", 0.0, w, np.float64)
    assert_close(am_g, o_am, GRAD_RTOL, 2e-6, "am grad")
    assert_close(lm_g, o_lm, GRAD_RTOL, 2e-5, "lm grad")


def test_c4_am_lm_gradients_on_tensor_cores():
    """A9 at the c4 shape, two utterances (T=1500 S=400 C=5000: 4 x 20 + 12 x 40 tiles per utterance)."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C = 2, 1500, 400, 5000
    rng = np.random.default_rng(44)
    am = rng.standard_normal((B, T, C), dtype=np.float32)
    lm = rng.standard_normal((B, S + 1, C), dtype=np.float32)
    sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
    bd = np.array([[0, 0, S, T], [0, 0, 250, 1100]], np.int32)
    term = C - 1
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    am_g, lm_g = frn.simple_loss_backward(lm, am, sym, term, bd, gx, gy, None, "regular")
    o_am, o_lm = orc.simple_am_lm_grad(lm, am, sym, term, bd, "regular", 0.0, -np.ones(B), np.float64)
    # the occupation counts themselves carry ~2e-4 at this lattice size (test_c4_large_vocab_bf16)
    assert_close(am_g, o_am, 4 * GRAD_RTOL, 2e-6, "c4 am grad")
    assert_close(lm_g, o_lm, 4 * GRAD_RTOL, 2e-5, "c4 lm grad")
