"""Pins of the oracle that do not need the reference tree: torchaudio's RNN-T
loss (CPU), and the invariants of SURVEY.md §8(c)(iii)."""
import numpy as np
import pytest

from oracle import rnnt_oracle as orc
from tests.helpers import make_inputs


def test_simple_loss_matches_torchaudio():
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    B, T, S, C = 3, 40, 12, 18
    am, lm, sym, term, bd = make_inputs(5, B, T, S, C, ragged=True)
    full = torch.from_numpy(am[:, :, None, :] + lm[:, None, :, :])
    want = ta.functional.rnnt_loss(full, torch.from_numpy(sym), torch.from_numpy(bd[:, 3].copy()),
                                   torch.from_numpy(bd[:, 2].copy()), blank=term, reduction="none",
                                   fused_log_softmax=True).numpy()
    got = orc.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", dtype=np.float64)
    np.testing.assert_allclose(got, want, rtol=2e-6)
    joint = orc.rnnt_loss(full.numpy(), sym, term, bd, "regular", 0.0, "none", dtype=np.float64)
    np.testing.assert_allclose(joint, want, rtol=2e-6)


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
def test_occupation_invariants(rnnt_type):
    B, T, S, C = 4, 50, 10, 16
    am, lm, sym, term, bd = make_inputs(1234, B, T, S, C, ragged=True)
    _, (gx, gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True, dtype=np.float64)
    for b in range(B):
        s_end, t_end = bd[b, 2], bd[b, 3]
        if rnnt_type == "regular":
            np.testing.assert_allclose(gy[b, :, :t_end].sum(axis=0), 1.0, rtol=1e-9)   # one blank per frame
            np.testing.assert_allclose(gx[b, :s_end, :].sum(axis=1), 1.0, rtol=1e-9)   # one arc per symbol
        else:
            np.testing.assert_allclose((gx[b, :, :t_end].sum(axis=0) + gy[b, :, :t_end].sum(axis=0)), 1.0, rtol=1e-9)
        assert gx[b, s_end:, :].sum() == 0 and gy[b, :, t_end:].sum() == 0            # nothing outside the box


def test_prune_range_invariants_and_full_band():
    B, T, S, C, R = 3, 60, 14, 20, 5
    am, lm, sym, term, bd = make_inputs(2, B, T, S, C, ragged=True)
    loss, (gx, gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    ranges = orc.get_rnnt_prune_ranges(gx, gy, bd, R)
    r0 = ranges[:, :, 0]
    assert (np.diff(r0, axis=1) >= 0).all() and (np.diff(r0, axis=1) < R).all() and (r0[:, 0] == 0).all()
    for b in range(B):
        assert r0[b, bd[b, 3] - 1] == max(bd[b, 2] - R + 1, 0)
    # pruned loss with the full band equals the unpruned loss
    full_ranges = orc.get_rnnt_prune_ranges(gx, gy, bd, S + 5)
    am_p, lm_p = orc.do_rnnt_pruning(am, lm, full_ranges)
    pl = orc.rnnt_loss_pruned(am_p + lm_p, sym, full_ranges, term, bd, "regular", 0.0, "none", dtype=np.float64)
    np.testing.assert_allclose(pl, loss, rtol=2e-6)


def test_logprobs_invariant_to_row_constants():
    B, T, S, C = 2, 20, 6, 9
    am, lm, sym, term, bd = make_inputs(3, B, T, S, C, ragged=False)
    px, py = orc.get_rnnt_logprobs(lm, am, sym, term, "regular", bd, np.float64)
    rng = np.random.default_rng(0)
    px2, py2 = orc.get_rnnt_logprobs(lm + rng.standard_normal((B, S + 1, 1)), am + rng.standard_normal((B, T, 1)),
                                     sym, term, "regular", bd, np.float64)
    fin = np.isfinite(px)
    np.testing.assert_allclose(px2[fin], px[fin], atol=1e-9)
    np.testing.assert_allclose(py2, py, atol=1e-9)


def test_analytic_gradients_against_finite_differences():
    B, T, S, C, R = 1, 8, 3, 5, 2
    am, lm, sym, term, bd = make_inputs(7, B, T, S, C, ragged=False)
    _, (gx, gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    ranges = orc.get_rnnt_prune_ranges(gx, gy, bd, R)
    am_p, lm_p = orc.do_rnnt_pruning(am, lm, ranges)
    logits = (am_p + lm_p).astype(np.float64)
    g = orc.pruned_logits_grad(logits, sym, ranges, term, bd, "regular", 0.2, None, np.float64)
    f = lambda x: orc.rnnt_loss_pruned(x, sym, ranges, term, bd, "regular", 0.2, "sum", dtype=np.float64)
    rng = np.random.default_rng(0)
    for _ in range(10):
        idx = tuple(rng.integers(0, n) for n in logits.shape)
        e = np.zeros_like(logits); e[idx] = 1e-6
        fd = (f(logits + e) - f(logits - e)) / 2e-6
        np.testing.assert_allclose(g[idx], fd, rtol=1e-5, atol=1e-8)
    ga, gl = orc.simple_am_lm_grad(lm, am, sym, term, bd, "regular", 0.0, None, np.float64)
    h = lambda a, l: orc.rnnt_loss_simple(l, a, sym, term, bd, "regular", 0.0, "sum", dtype=np.float64)
    for _ in range(6):
        idx = tuple(rng.integers(0, n) for n in am.shape)
        e = np.zeros(am.shape); e[idx] = 1e-6
        np.testing.assert_allclose(ga[idx], (h(am + e, lm) - h(am - e, lm)) / 2e-6, rtol=1e-5, atol=1e-8)
        idx = tuple(rng.integers(0, n) for n in lm.shape)
        e = np.zeros(lm.shape); e[idx] = 1e-6
        np.testing.assert_allclose(gl[idx], (h(am, lm + e) - h(am, lm - e)) / 2e-6, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
def test_smoothed_am_lm_grad_matches_finite_differences(rnnt_type):
    """The analytic A9 gradient of the smoothed loss (incl. the batch-global unigram path)
    against central differences of the oracle's own float64 loss."""
    rng = np.random.default_rng(5)
    B, T, S, C = 2, 5, 3, 4
    am = rng.standard_normal((B, T, C))
    lm = rng.standard_normal((B, S + 1, C))
    sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
    bd = np.array([[0, 0, S, T], [0, 0, S - 1, T - 1]], np.int32)
    w = np.array([1.0, -0.7])
    args = dict(lm_only_scale=0.25, am_only_scale=0.15, rnnt_type=rnnt_type, delay_penalty=0.1)

    def total(lm_, am_):
        loss = orc.rnnt_loss_smoothed(lm_, am_, sym, C - 1, boundary=bd, reduction="none",
                                      dtype=np.float64, **args)
        return float((w * loss).sum())

    am_g, lm_g = orc.smoothed_am_lm_grad(lm, am, sym, C - 1, bd, loss_grad=w, dtype=np.float64, **args)
    eps = 1e-6
    for arr, grad, which in ((am, am_g, "am"), (lm, lm_g, "lm")):
        num = np.zeros_like(arr)
        for idx in np.ndindex(arr.shape):
            old = arr[idx]
            arr[idx] = old + eps
            hi = total(lm, am)
            arr[idx] = old - eps
            lo = total(lm, am)
            arr[idx] = old
            num[idx] = (hi - lo) / (2 * eps)
        np.testing.assert_allclose(grad, num, rtol=2e-5, atol=2e-7, err_msg=which)
