"""GPU parity of the rest of the hot path through the public API / C ABI:
simple & smoothed log-probs and losses (A1-A3), prune ranges (A5), pruning
(A6), pruned log-probs, loss and logits gradient (A7, A8), full joiner (f1)."""
import numpy as np
import pytest

from oracle import rnnt_oracle as orc
from tests.helpers import (GRAD_ATOL, GRAD_RTOL, LOSS_RTOL, assert_close, load_golden, make_inputs)

pytestmark = pytest.mark.gpu
LOGP_RTOL, LOGP_ATOL = 2e-6, 2e-5     # log-probs: float32 vs float64 oracle


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", ["c1_readme", "stress_b2_t200_s50_c50"])
def test_golden_simple_and_smoothed(name):
    import tf_fast_rnnt as frn
    g = load_golden(name)
    term = int(g["termination_symbol"])
    px, py = frn.get_rnnt_logprobs(g["lm"], g["am"], g["symbols"], term, "regular", g["boundary"])
    assert_close(px, g["simple_px"], LOGP_RTOL, LOGP_ATOL, "px")
    assert_close(py, g["simple_py"], LOGP_RTOL, LOGP_ATOL, "py")
    for dp, tag in ((0.0, "dp0"), (0.2, "dp2")):
        loss, (gx, gy) = frn.rnnt_loss_simple(g["lm"], g["am"], g["symbols"], term, g["boundary"],
                                              "regular", dp, "none", True)
        assert_close(loss, g[f"simple_loss_{tag}"], LOSS_RTOL, 0, "loss")
        # the golden grads are the reference's float32 recursion: 1e-3 is its own accuracy
        assert_close(gx, g[f"simple_px_grad_{tag}"], 2e-3, 1e-6, "px_grad")
        assert_close(gy, g[f"simple_py_grad_{tag}"], 2e-3, 1e-6, "py_grad")
    s = frn.rnnt_loss_simple(g["lm"], g["am"], g["symbols"], term, g["boundary"], reduction="sum")
    assert_close(s, g["simple_loss_sum"], LOSS_RTOL, 0, "sum")
    for lms, ams in ((0.1, 0.2), (0.25, 0.0)):
        tag = f"l{int(lms * 100)}_a{int(ams * 100)}"
        spx, spy = frn.get_rnnt_logprobs_smoothed(g["lm"], g["am"], g["symbols"], term, lms, ams,
                                                  g["boundary"], "regular")
        assert_close(spx, g[f"smoothed_px_{tag}"], LOGP_RTOL, LOGP_ATOL, "smoothed px")
        assert_close(spy, g[f"smoothed_py_{tag}"], LOGP_RTOL, LOGP_ATOL, "smoothed py")
        loss = frn.rnnt_loss_smoothed(g["lm"], g["am"], g["symbols"], term, lms, ams, g["boundary"],
                                      "regular", 0.2, "none")
        assert_close(loss, g[f"smoothed_loss_{tag}"], LOSS_RTOL, 0, "smoothed loss")


@pytest.mark.parametrize("name", ["c1_readme", "stress_b2_t200_s50_c50"])
def test_golden_prune_and_pruned(name):
    import tf_fast_rnnt as frn
    g = load_golden(name)
    term = int(g["termination_symbol"])
    for r in g["s_ranges"]:
        ranges = frn.get_rnnt_prune_ranges(g["simple_px_grad_dp2"], g["simple_py_grad_dp2"],
                                           g["boundary"], int(r))
        assert ranges.dtype == np.int32
        assert np.array_equal(ranges, g[f"ranges_r{r}"]), f"s_range={r}"          # bit-exact
        am_p, lm_p = frn.do_rnnt_pruning(g["am"], g["lm"], ranges)
        o_am_p, o_lm_p = orc.do_rnnt_pruning(g["am"], g["lm"], ranges)
        assert np.array_equal(am_p, o_am_p) and np.array_equal(lm_p, o_lm_p)      # bit-exact copy
        logits = (1.0 / (1.0 + np.exp(-(am_p + lm_p)))).astype(np.float32)
        for rt in ("regular", "modified", "constrained"):
            px, py = frn.get_rnnt_logprobs_pruned(logits, g["symbols"], ranges, term, g["boundary"], rt)
            assert_close(px, g[f"pruned_px_r{r}_{rt}"], LOGP_RTOL, 5e-6, f"pruned px {rt}")
            assert_close(py, g[f"pruned_py_r{r}_{rt}"], LOGP_RTOL, 5e-6, f"pruned py {rt}")
            for dp in (0.0, 0.2):
                loss = frn.rnnt_loss_pruned(logits, g["symbols"], ranges, term, g["boundary"], rt, dp, "none")
                assert_close(loss, g[f"pruned_loss_r{r}_{rt}_dp{int(dp * 10)}"], LOSS_RTOL, 0,
                             f"pruned loss r={r} {rt} dp={dp}")
    r0 = int(g["s_ranges"][0])
    am_p, lm_p = frn.do_rnnt_pruning(g["am"], g["lm"], g[f"ranges_r{r0}"])
    assert np.array_equal(am_p, g[f"am_pruned_r{r0}"]) and np.array_equal(lm_p, g[f"lm_pruned_r{r0}"])


@pytest.mark.parametrize("name", ["c1_readme", "stress_b2_t200_s50_c50"])
def test_golden_joint(name):
    import tf_fast_rnnt as frn
    g = load_golden(name)
    term = int(g["termination_symbol"])
    full = (1.0 / (1.0 + np.exp(-(g["am"][:, :, None, :] + g["lm"][:, None, :, :])))).astype(np.float32)
    px, py = frn.get_rnnt_logprobs_joint(full, g["symbols"], term, g["boundary"])
    assert_close(px, g["joint_px"], LOGP_RTOL, 5e-6, "joint px")
    assert_close(py, g["joint_py"], LOGP_RTOL, 5e-6, "joint py")
    loss = frn.rnnt_loss(full, g["symbols"], term, g["boundary"], "regular", 0.2, "none")
    assert_close(loss, g["joint_loss_dp2"], LOSS_RTOL, 0, "joint loss")


# ------------------------------------------------------------------ seeded inputs vs float64 oracle
@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
@pytest.mark.parametrize("shape", [(2, 50, 10, 16), (3, 70, 33, 37), (2, 200, 130, 50)])
@pytest.mark.parametrize("dp", [0.0, 0.2])
def test_simple_loss_types(rnnt_type, shape, dp):
    import tf_fast_rnnt as frn
    B, T, S, C = shape
    am, lm, sym, term, bd = make_inputs(100 + S, B, T, S, C, ragged=True)
    loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, dp, "none", True)
    o_loss, (o_gx, o_gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, dp, "none", True,
                                                dtype=np.float64)
    assert_close(loss, o_loss, LOSS_RTOL, 0, "loss")
    assert_close(gx, o_gx, GRAD_RTOL, GRAD_ATOL, "px_grad")
    assert_close(gy, o_gy, GRAD_RTOL, GRAD_ATOL, "py_grad")
    for red in ("mean", "sum"):
        v = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, dp, red)
        o = orc.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, dp, red, dtype=np.float64)
        assert_close(v, o, LOSS_RTOL, 0, red)


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
def test_smoothed_loss_types(rnnt_type):
    import tf_fast_rnnt as frn
    B, T, S, C = 3, 60, 20, 30
    am, lm, sym, term, bd = make_inputs(7, B, T, S, C, ragged=True)
    for lms, ams in ((0.25, 0.0), (0.1, 0.2)):
        px, py = frn.get_rnnt_logprobs_smoothed(lm, am, sym, term, lms, ams, bd, rnnt_type)
        o_px, o_py = orc.get_rnnt_logprobs_smoothed(lm, am, sym, term, lms, ams, bd, rnnt_type, np.float64)
        assert_close(px, o_px, LOGP_RTOL, LOGP_ATOL, "px")
        assert_close(py, o_py, LOGP_RTOL, LOGP_ATOL, "py")
        loss, (gx, gy) = frn.rnnt_loss_smoothed(lm, am, sym, term, lms, ams, bd, rnnt_type, 0.2, "none", True)
        o_loss, (o_gx, o_gy) = orc.rnnt_loss_smoothed(lm, am, sym, term, lms, ams, bd, rnnt_type, 0.2,
                                                      "none", True, dtype=np.float64)
        assert_close(loss, o_loss, LOSS_RTOL, 0, "loss")
        assert_close(gx, o_gx, GRAD_RTOL, GRAD_ATOL, "px_grad")
        assert_close(gy, o_gy, GRAD_RTOL, GRAD_ATOL, "py_grad")


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
@pytest.mark.parametrize("s_range", [2, 5, 9, 100])
def test_pipeline_vs_oracle(rnnt_type, s_range):
    """simple -> prune ranges -> pruning -> pruned loss (+ logits gradient)."""
    import torch
    import tf_fast_rnnt as frn
    if rnnt_type == "regular" and s_range < 2:
        pytest.skip("regular needs s_range >= 2")
    B, T, S, C = 3, 80, 24, 20
    am, lm, sym, term, bd = make_inputs(31, B, T, S, C, ragged=True)
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, s_range)
    assert np.array_equal(ranges, orc.get_rnnt_prune_ranges(gx, gy, bd, s_range))
    am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
    logits = np.tanh(am_p + lm_p).astype(np.float32) * 3
    for dp in (0.0, 0.2):
        loss = frn.rnnt_loss_pruned(logits, sym, ranges, term, bd, rnnt_type, dp, "none")
        o_loss = orc.rnnt_loss_pruned(logits, sym, ranges, term, bd, rnnt_type, dp, "none", dtype=np.float64)
        assert_close(loss, o_loss, LOSS_RTOL, 0, f"pruned loss dp={dp}")
    # logits gradient of sum_b w_b * loss_b
    w = np.array([1.0, -0.5, 2.0], np.float32)
    scores, grad = frn.pruned_loss_fwd_bwd(torch.from_numpy(logits).cuda(), sym, ranges, term, bd, rnnt_type,
                                           0.2, -w)
    o_grad = orc.pruned_logits_grad(logits, sym, ranges, term, bd, rnnt_type, 0.2, w, np.float64)
    assert_close(grad.cpu().numpy(), o_grad, GRAD_RTOL, 2e-6, "logits grad")
    # autograd wrapper
    lg = torch.from_numpy(logits).cuda().requires_grad_(True)
    l = frn.rnnt_loss_pruned(lg, sym, ranges, term, bd, rnnt_type, 0.2, "sum")
    l.backward()
    o_grad1 = orc.pruned_logits_grad(logits, sym, ranges, term, bd, rnnt_type, 0.2, None, np.float64)
    assert_close(lg.grad.cpu().numpy(), o_grad1, GRAD_RTOL, 2e-6, "autograd logits grad")


def test_pruned_full_band_equals_unpruned():
    import tf_fast_rnnt as frn
    B, T, S, C = 2, 40, 12, 18
    am, lm, sym, term, bd = make_inputs(3, B, T, S, C, ragged=True)
    full = am[:, :, None, :] + lm[:, None, :, :]
    simple = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none")
    joint = frn.rnnt_loss(full, sym, term, bd, "regular", 0.0, "none")
    assert_close(joint, simple, LOSS_RTOL, 0, "joint == simple for an additive joiner")
    try:
        import torchaudio
        import torch
        ta = torchaudio.functional.rnnt_loss(
            torch.from_numpy(full), torch.from_numpy(sym), torch.from_numpy(bd[:, 3].copy()),
            torch.from_numpy(bd[:, 2].copy()), blank=term, reduction="none", fused_log_softmax=True)
        assert_close(joint, ta.numpy(), LOSS_RTOL, 0, "torchaudio cross-check")
    except ImportError:
        pass


def test_bf16_logits():
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C, R = 2, 60, 20, 64, 5
    am, lm, sym, term, bd = make_inputs(9, B, T, S, C, ragged=True)
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    lg = frn.pruned_add_joiner(torch.from_numpy(am).cuda(), torch.from_numpy(lm).cuda(),
                               torch.from_numpy(ranges).cuda(), dtype=torch.bfloat16)
    assert lg.dtype == torch.bfloat16
    up = lg.float().cpu().numpy()        # oracle: upcast then reference math
    scores, grad = frn.pruned_loss_fwd_bwd(lg, sym, ranges, term, bd, "regular", 0.0, None)
    o_loss = orc.rnnt_loss_pruned(up, sym, ranges, term, bd, "regular", 0.0, "none", dtype=np.float64)
    assert_close(-scores.cpu().numpy(), o_loss, LOSS_RTOL, 0, "bf16 loss")
    o_grad = orc.pruned_logits_grad(up, sym, ranges, term, bd, "regular", 0.0, -np.ones(B), np.float64)
    assert grad.dtype == torch.bfloat16
    assert_close(grad.float().cpu().numpy(), o_grad, 1e-2, 1e-4, "bf16 logits grad (bf16 rounding)")


def test_pruning_backward():
    import tf_fast_rnnt as frn
    B, T, S, C, R = 2, 30, 9, 12, 4
    am, lm, sym, term, bd = make_inputs(4, B, T, S, C, ragged=True)
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    rng = np.random.default_rng(1)
    ga = rng.standard_normal((B, T, R, C), dtype=np.float32)
    gl = rng.standard_normal((B, T, R, C), dtype=np.float32)
    am_g, lm_g = frn.do_rnnt_pruning_backward(ga, gl, ranges, S)
    o_am, o_lm = orc.do_rnnt_pruning_bwd(ga, gl, ranges, S + 1)
    assert_close(am_g, o_am, 1e-5, 1e-5, "am grad")
    assert_close(lm_g, o_lm, 1e-5, 1e-5, "lm grad")


def test_argument_errors():
    import tf_fast_rnnt as frn
    am, lm, sym, term, bd = make_inputs(1, 2, 20, 5, 8)
    with pytest.raises(ValueError):
        frn.rnnt_loss_simple(lm, am, sym, term, bd, reduction="bogus")
    with pytest.raises(ValueError):
        frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type="bogus")
    with pytest.raises(ValueError):
        frn.mutual_information_recursion(np.zeros((2, 3, 9), np.float32), np.zeros((2, 4, 5), np.float32), None)


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
def test_simple_loss_am_lm_gradients(rnnt_type):
    """A9: gradients w.r.t. am / lm (C-ABI frn_simple_loss_bwd and the autograd wrapper)."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C = 3, 70, 21, 28
    am, lm, sym, term, bd = make_inputs(17, B, T, S, C, ragged=True)
    w = np.array([1.0, 0.5, -2.0], np.float32)
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.2, "none", True)
    am_g, lm_g = frn.simple_loss_backward(lm, am, sym, term, bd, gx, gy, -w, rnnt_type)
    o_am, o_lm = orc.simple_am_lm_grad(lm, am, sym, term, bd, rnnt_type, 0.2, w, np.float64)
    assert_close(am_g, o_am, GRAD_RTOL, 2e-6, "am grad")
    assert_close(lm_g, o_lm, GRAD_RTOL, 2e-5, "lm grad")       # lm rows sum ~T terms
    lm_t = torch.from_numpy(lm).cuda().requires_grad_(True)
    am_t = torch.from_numpy(am).cuda().requires_grad_(True)
    loss = frn.rnnt_loss_simple(lm_t, am_t, sym, term, bd, rnnt_type, 0.2, "sum")
    loss.backward()
    o_am1, o_lm1 = orc.simple_am_lm_grad(lm, am, sym, term, bd, rnnt_type, 0.2, None, np.float64)
    assert_close(am_t.grad.cpu().numpy(), o_am1, GRAD_RTOL, 2e-6, "autograd am grad")
    assert_close(lm_t.grad.cpu().numpy(), o_lm1, GRAD_RTOL, 2e-5, "autograd lm grad")


@pytest.mark.parametrize("C", [500, 37, 504])     # vector path (C % 4 == 0), scalar path, bf16 vector path (C % 8 == 0)
def test_fused_joiner_equals_pruning_plus_add(C):
    """frn_pruned_add_joiner == do_rnnt_pruning followed by the addition, bit for bit;
    out-of-range band entries contribute zeros like tf.gather on GPU."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, R = 3, 40, 12, 5
    rng = np.random.default_rng(C)
    am = rng.standard_normal((B, T, C), dtype=np.float32)
    lm = rng.standard_normal((B, S + 1, C), dtype=np.float32)
    r0 = np.sort(rng.integers(0, S + 2 - R, (B, T)), axis=1)
    ranges = (r0[:, :, None] + np.arange(R)[None, None, :]).astype(np.int32)
    ranges[0, 3, 4] = S + 5                        # out of range: zeros
    am_d, lm_d, rg_d = (torch.from_numpy(x).cuda() for x in (am, lm, ranges))
    am_p, lm_p = frn.do_rnnt_pruning(am_d, lm_d, rg_d)
    fused = frn.pruned_add_joiner(am_d, lm_d, rg_d)
    assert torch.equal(fused, am_p + lm_p)
    assert torch.equal(fused[0, 3, 4], am_d[0, 3])
    # bf16 output (BASELINE configs[3]): the float32 sum rounded to nearest even, vector (C % 8 == 0) or scalar kernel
    fused16 = frn.pruned_add_joiner(am_d, lm_d, rg_d, dtype=torch.bfloat16)
    assert fused16.dtype == torch.bfloat16 and torch.equal(fused16, (am_p + lm_p).to(torch.bfloat16))


@pytest.mark.parametrize("C,R", [(500, 5), (37, 5), (64, 11)])   # one-pass kernel; scalar and wide-band fall-backs
def test_pruning_with_joiner_equals_separate_ops(C, R):
    """frn_do_pruning_add_joiner == do_rnnt_pruning + addition, all three tensors bit for bit."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S = 3, 40, 12
    rng = np.random.default_rng([C, R])
    am = rng.standard_normal((B, T, C), dtype=np.float32)
    lm = rng.standard_normal((B, S + 1, C), dtype=np.float32)
    r0 = np.sort(rng.integers(0, S + 2 - R, (B, T)), axis=1)
    ranges = (r0[:, :, None] + np.arange(R)[None, None, :]).astype(np.int32)
    am_d, lm_d, rg_d = (torch.from_numpy(x).cuda() for x in (am, lm, ranges))
    am_p, lm_p = frn.do_rnnt_pruning(am_d, lm_d, rg_d)
    am_q, lm_q, logits = frn.do_rnnt_pruning_add_joiner(am_d, lm_d, rg_d)
    assert torch.equal(am_q, am_p) and torch.equal(lm_q, lm_p)
    assert torch.equal(logits, am_p + lm_p)
    # halves of do_rnnt_pruning on their own (the am broadcast does not need the ranges)
    lib, ptr = frn._lib.lib, lambda t: t.data_ptr()
    if C % 4 == 0 and R <= 8:
        a2, l2 = torch.zeros_like(am_p), torch.zeros_like(lm_p)
        st = torch.cuda.current_stream().cuda_stream
        frn._lib.check(lib.frn_do_pruning(ptr(am_d), 0, 0, B, S, T, R, C, ptr(a2), 0, st), "am half")
        frn._lib.check(lib.frn_do_pruning(0, ptr(lm_d), ptr(rg_d), B, S, T, R, C, 0, ptr(l2), st), "lm half")
        assert torch.equal(a2, am_p) and torch.equal(l2, lm_p)


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
@pytest.mark.parametrize("scales", [(0.25, 0.15), (0.25, 0.0)])
def test_smoothed_loss_am_lm_gradients(rnnt_type, scales):
    """A9 for rnnt_loss_smoothed: C-ABI frn_smoothed_loss_bwd and the autograd wrapper against the
    float64 oracle (itself checked against finite differences on the CPU)."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C = 3, 70, 21, 28
    lms, ams = scales
    am, lm, sym, term, bd = make_inputs(23, B, T, S, C, ragged=True)
    w = np.array([1.0, 0.5, -2.0], np.float32)
    _, (gx, gy) = frn.rnnt_loss_smoothed(lm, am, sym, term, lms, ams, bd, rnnt_type, 0.2, "none", True)
    am_g, lm_g = frn.smoothed_loss_backward(lm, am, sym, term, bd, gx, gy, -w, lms, ams, rnnt_type)
    o_am, o_lm = orc.smoothed_am_lm_grad(lm, am, sym, term, bd, lms, ams, rnnt_type, 0.2, w, np.float64)
    assert_close(am_g, o_am, GRAD_RTOL, 2e-6, "am grad")
    assert_close(lm_g, o_lm, GRAD_RTOL, 2e-5, "lm grad")
    lm_t = torch.from_numpy(lm).cuda().requires_grad_(True)
    am_t = torch.from_numpy(am).cuda().requires_grad_(True)
    loss = frn.rnnt_loss_smoothed(lm_t, am_t, sym, term, lms, ams, bd, rnnt_type, 0.2, "sum")
    loss.backward()
    o_am1, o_lm1 = orc.smoothed_am_lm_grad(lm, am, sym, term, bd, lms, ams, rnnt_type, 0.2, None, np.float64)
    assert_close(am_t.grad.cpu().numpy(), o_am1, GRAD_RTOL, 2e-6, "autograd am grad")
    assert_close(lm_t.grad.cpu().numpy(), o_lm1, GRAD_RTOL, 2e-5, "autograd lm grad")


def test_fuzz_band_recursion_equals_dense_recursions(monkeypatch):
    """Fuzz: rnnt_loss_pruned through the band recursion (band_dp.cu) against the same call forced onto the
    dense-lattice kernels (FRN_BAND_DENSE=1, wavefront and row scan): three independent implementations of
    the recursion on random shapes, ragged boundaries, all rnnt types, delay penalty and some -inf logits."""
    import torch
    import tf_fast_rnnt as frn
    rng = np.random.default_rng(77)
    for case in range(36):
        rnnt_type = ["regular", "modified", "constrained"][case % 3]
        B = int(rng.integers(1, 4)); S = int(rng.integers(2, 40)); T = int(rng.integers(max(S, 4), 300))
        C = int(rng.integers(3, 20)); R = int(rng.integers(1, min(8, S + 1) + 1))
        am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=False)
        dp = [0.0, 0.3][case % 2]
        monkeypatch.delenv("FRN_BAND_DENSE", raising=False)
        monkeypatch.delenv("FRN_DP_CHAIN", raising=False)
        monkeypatch.delenv("FRN_DP_SCAN", raising=False)
        _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True)
        ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
        am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
        logits = am_p + lm_p
        if case % 4 == 0:
            logits[rng.random(logits.shape) < 0.01] = -np.inf
        w = rng.standard_normal(B).astype(np.float32)
        res = []
        for env in ({}, {"FRN_BAND_DENSE": "1", "FRN_DP_CHAIN": "1"}, {"FRN_BAND_DENSE": "1", "FRN_DP_SCAN": "1"}):
            for k in ("FRN_BAND_DENSE", "FRN_DP_CHAIN", "FRN_DP_SCAN"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            sc, gr = frn.pruned_loss_fwd_bwd(torch.from_numpy(logits).cuda(), sym, ranges, term, bd, rnnt_type, dp,
                                             torch.from_numpy(w).cuda())
            res.append((sc.cpu().numpy(), gr.cpu().numpy()))
        tag = f"case {case}: {rnnt_type} B={B} S={S} T={T} C={C} R={R} dp={dp}"
        (s0, g0) = res[0]
        for which, (s1, g1) in zip(("wavefront", "row scan"), res[1:]):
            assert np.array_equal(np.isfinite(s0), np.isfinite(s1)), tag + " " + which
            ok = np.isfinite(s0)
            assert_close(s1[ok], s0[ok], 2e-6, 1e-5, tag + f" scores vs {which}")
            assert_close(np.nan_to_num(g1[ok]), np.nan_to_num(g0[ok]), 5e-5, 2e-6, tag + f" logits grad vs {which}")


def test_fuzz_tensor_core_normaliser_equals_simt(monkeypatch):
    """Fuzz: the tcgen05 normaliser (two-term float16 split) against the exact-FP32 SIMT kernel of the same op on
    random shapes (C % 4 == 0 so that both paths exist), simple and smoothed, all rnnt types."""
    import tf_fast_rnnt as frn
    rng = np.random.default_rng(78)
    for case in range(24):
        rnnt_type = ["regular", "modified", "constrained"][case % 3]
        B = int(rng.integers(1, 4)); S = int(rng.integers(1, 150)); T = int(rng.integers(1, 300))
        C = 4 * int(rng.integers(1, 160))
        am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=False)
        am *= 3.0
        out = []
        for simt in ("0", "1"):
            monkeypatch.setenv("FRN_SIMPLE_SIMT", simt)
            if case % 2:
                out.append(frn.get_rnnt_logprobs_smoothed(lm, am, sym, term, 0.25, 0.1, bd, rnnt_type))
            else:
                out.append(frn.get_rnnt_logprobs(lm, am, sym, term, rnnt_type, bd))
        tag = f"case {case}: {rnnt_type} B={B} S={S} T={T} C={C} smoothed={case % 2}"
        assert_close(out[0][0], out[1][0], 2e-6, 2e-6, tag + " px")
        assert_close(out[0][1], out[1][1], 2e-6, 2e-6, tag + " py")


@pytest.mark.parametrize("scale", [8.0, 14.0])
def test_normaliser_on_peaky_distributions(scale):
    """The two-term float16 operands of the tensor-core normaliser on the inputs they are weakest on: peaky am / lm
    rows (probabilities down to e^-100 relative to the row maximum), row maxima far from zero, am and lm peaking
    at DIFFERENT classes (Z = sum_c p_am p_lm is then made of small products only), against the float64 oracle."""
    import tf_fast_rnnt as frn
    rng = np.random.default_rng(int(scale))
    B, T, S, C = 2, 150, 40, 256
    am = (scale * rng.standard_normal((B, T, C))).astype(np.float32)
    lm = (scale * rng.standard_normal((B, S + 1, C))).astype(np.float32)
    am += rng.uniform(-30, 30, (B, T, 1)).astype(np.float32)            # row offsets cancel in px / py
    lm += rng.uniform(-30, 30, (B, S + 1, 1)).astype(np.float32)
    am[0, :, 7] += 4 * scale                                           # utterance 0: am peaks at class 7 ...
    lm[0, :, 11] += 4 * scale                                          # ... lm at class 11
    sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
    term = C - 1
    bd = np.tile(np.array([0, 0, S, T], np.int32), (B, 1))
    px, py = frn.get_rnnt_logprobs(lm, am, sym, term, "regular", bd)
    o_px, o_py = orc.get_rnnt_logprobs(lm, am, sym, term, "regular", bd, dtype=np.float64)
    # px = am[sym] + lm[sym] - log Z - ammax - lmmax in float32: a few ulp of the largest operand, + the 2^-21 of Z
    atol = 4e-7 * float(np.abs(am).max() + np.abs(lm).max())
    # the float32 reference saturates at log(0 + tiny) = -103.3 once Z underflows (rnnt_loss.py:181), and within
    # e^10 of the underflow edge the terms of Z themselves flush to zero one by one; compare where Z and its leading
    # terms are representable in float32 (log Z > -70) - at the larger scale some cells of utterance 0 are not
    lm64, am64 = lm.astype(np.float64), am.astype(np.float64)
    z = np.einsum("bsc,btc->bst", np.exp(lm64 - lm64.max(2, keepdims=True)), np.exp(am64 - am64.max(2, keepdims=True)))
    ok = np.log(z) > -70.0
    assert ok.mean() > 0.6 and (np.log(z[0]) < -25.0).mean() > 0.5       # the hard regime is what is being tested
    assert_close(np.where(ok[:, :S, :], px[:, :, :T], 0), np.where(ok[:, :S, :], o_px[:, :, :T], 0), 3e-7, atol,
                 f"px, scale {scale}")
    assert_close(np.where(ok, py, 0), np.where(ok, o_py, 0), 3e-7, atol, f"py, scale {scale}")
    if not ok.all():
        return                                   # a saturated cell changes the float32 reference's loss: nothing to compare
    loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
    o_loss, (o_gx, o_gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True, dtype=np.float64)
    assert_close(loss, o_loss, LOSS_RTOL, 0, f"loss, scale {scale}")
    assert_close(gx, o_gx, GRAD_RTOL, GRAD_ATOL, f"px_grad, scale {scale}")
    assert_close(gy, o_gy, GRAD_RTOL, GRAD_ATOL, f"py_grad, scale {scale}")


def test_fuzz_pipeline_against_float64_oracle(monkeypatch):
    """Fuzz of the whole path against the float64 oracle: random shapes, ragged boundaries with offsets, all
    rnnt types, delay penalties, s_range 1..9 (band recursion and dense fall-back), both dense-lattice kernels."""
    import torch
    import tf_fast_rnnt as frn
    rng = np.random.default_rng(4242)
    for case in range(30):
        rnnt_type = ["regular", "modified", "constrained"][case % 3]
        B = int(rng.integers(1, 4)); S = int(rng.integers(1, 26)); T = int(rng.integers(max(S, 2), 130))
        C = int(rng.integers(2, 12)) * 4 if case % 2 else int(rng.integers(3, 30))
        R = int(rng.integers(2 if rnnt_type == "regular" else 1, 10))
        am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=bool(case % 5 == 0))
        dp = float([0.0, 0.25, 0.6][case % 3 if case % 2 else 0])
        for k in ("FRN_DP_CHAIN", "FRN_DP_SCAN"):
            monkeypatch.delenv(k, raising=False)
        monkeypatch.setenv("FRN_DP_SCAN" if case % 2 else "FRN_DP_CHAIN", "1")
        tag = f"case {case}: {rnnt_type} B={B} S={S} T={T} C={C} R={R} dp={dp} bd={bd.tolist()}"
        loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, dp, "none", True)
        o_loss, (o_gx, o_gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, dp, "none", True, dtype=np.float64)
        assert_close(loss, o_loss, LOSS_RTOL, 1e-5, tag + " simple loss")
        assert_close(gx, o_gx, GRAD_RTOL, GRAD_ATOL, tag + " px_grad")
        assert_close(gy, o_gy, GRAD_RTOL, GRAD_ATOL, tag + " py_grad")
        ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
        assert np.array_equal(ranges, orc.get_rnnt_prune_ranges(gx, gy, bd, R)), tag + " ranges"
        am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
        o_am_p, o_lm_p = orc.do_rnnt_pruning(am, lm, ranges)
        assert np.array_equal(am_p, o_am_p) and np.array_equal(lm_p, o_lm_p), tag + " pruning"
        logits = (am_p + lm_p).astype(np.float32)
        w = rng.standard_normal(B).astype(np.float32)
        scores, grad = frn.pruned_loss_fwd_bwd(torch.from_numpy(logits).cuda(), sym, ranges, term, bd, rnnt_type, dp,
                                               torch.from_numpy(-w).cuda())
        o_grad, o_scores = orc.pruned_logits_grad(logits, sym, ranges, term, bd, rnnt_type, dp, w, np.float64,
                                                  return_scores=True)
        assert_close(scores.cpu().numpy(), o_scores, LOSS_RTOL, 1e-5, tag + " pruned scores")
        ok = np.isfinite(o_scores)        # no path inside the band (score -inf): the gradient is undefined there
        assert_close(grad.cpu().numpy()[ok], o_grad[ok], GRAD_RTOL, 2e-6, tag + " logits grad")
        assert not np.isnan(grad.cpu().numpy()).any(), tag


def test_fuzz_remaining_entry_points_against_oracle():
    """Fuzz of the entry points the pipeline fuzz does not reach: smoothed loss, am/lm gradients of the simple and
    smoothed losses (A9), the joint loss on the full joiner output, the pruning gradient and bf16 logits - random
    shapes, ragged boundaries, all rnnt types, against the float64 oracle."""
    import torch
    import tf_fast_rnnt as frn
    rng = np.random.default_rng(909)
    for case in range(18):
        rnnt_type = ["regular", "modified", "constrained"][case % 3]
        B = int(rng.integers(1, 4)); S = int(rng.integers(1, 20)); T = int(rng.integers(max(S, 2), 70))
        C = 4 * int(rng.integers(1, 9)) if case % 2 else int(rng.integers(2, 25))
        am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=bool(case % 4 == 0))
        dp = float([0.0, 0.3][case % 2])
        lms, ams = float(rng.uniform(0.0, 0.4)), float([0.0, 0.2][case % 2])
        w = rng.standard_normal(B).astype(np.float32)
        tag = f"case {case}: {rnnt_type} B={B} S={S} T={T} C={C} dp={dp} scales=({lms:.2f},{ams:.2f}) bd={bd.tolist()}"
        # smoothed loss forward + occupation counts
        loss, (gx, gy) = frn.rnnt_loss_smoothed(lm, am, sym, term, lms, ams, bd, rnnt_type, dp, "none", True)
        o_loss, (o_gx, o_gy) = orc.rnnt_loss_smoothed(lm, am, sym, term, lms, ams, bd, rnnt_type, dp, "none", True,
                                                      dtype=np.float64)
        assert_close(loss, o_loss, LOSS_RTOL, 1e-5, tag + " smoothed loss")
        assert_close(gx, o_gx, GRAD_RTOL, GRAD_ATOL, tag + " smoothed px_grad")
        assert_close(gy, o_gy, GRAD_RTOL, GRAD_ATOL, tag + " smoothed py_grad")
        # A9, smoothed and simple
        am_g, lm_g = frn.smoothed_loss_backward(lm, am, sym, term, bd, gx, gy, -w, lms, ams, rnnt_type)
        o_am, o_lm = orc.smoothed_am_lm_grad(lm, am, sym, term, bd, lms, ams, rnnt_type, dp, w, np.float64)
        assert_close(am_g, o_am, GRAD_RTOL, 5e-6, tag + " smoothed am grad")
        assert_close(lm_g, o_lm, GRAD_RTOL, 3e-5, tag + " smoothed lm grad")
        _, (sgx, sgy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, dp, "none", True)
        am_g, lm_g = frn.simple_loss_backward(lm, am, sym, term, bd, sgx, sgy, -w, rnnt_type)
        o_am, o_lm = orc.simple_am_lm_grad(lm, am, sym, term, bd, rnnt_type, dp, w, np.float64)
        assert_close(am_g, o_am, GRAD_RTOL, 5e-6, tag + " simple am grad")
        assert_close(lm_g, o_lm, GRAD_RTOL, 3e-5, tag + " simple lm grad")
        # joint loss on the full joiner output
        full = (am[:, :, None, :] + lm[:, None, :, :]).astype(np.float32)
        jl = frn.rnnt_loss(full, sym, term, bd, rnnt_type, dp, "none")
        o_jl = orc.rnnt_loss(full, sym, term, bd, rnnt_type, dp, "none", dtype=np.float64)
        assert_close(jl, o_jl, LOSS_RTOL, 1e-5, tag + " joint loss")
        # pruning gradient
        R = int(rng.integers(1, min(S, 6) + 1)) if rnnt_type != "regular" else int(rng.integers(2, 7))
        ranges = frn.get_rnnt_prune_ranges(sgx, sgy, bd, R)
        Rw = ranges.shape[2]
        ga = rng.standard_normal((B, T, Rw, C), dtype=np.float32)
        gl = rng.standard_normal((B, T, Rw, C), dtype=np.float32)
        a_g, l_g = frn.do_rnnt_pruning_backward(ga, gl, ranges, S)
        o_a, o_l = orc.do_rnnt_pruning_bwd(ga, gl, ranges, S + 1)
        assert_close(a_g, o_a, 1e-5, 1e-5, tag + " pruning am grad")
        assert_close(l_g, o_l, 1e-5, 1e-4, tag + " pruning lm grad")
        # bf16 logits into the float32 pruned loss
        lg = frn.pruned_add_joiner(torch.from_numpy(am).cuda(), torch.from_numpy(lm).cuda(),
                                   torch.from_numpy(ranges).cuda(), dtype=torch.bfloat16)
        scores, _ = frn.pruned_loss_fwd_bwd(lg, sym, ranges, term, bd, rnnt_type, dp, None)
        o_pl = orc.rnnt_loss_pruned(lg.float().cpu().numpy(), sym, ranges, term, bd, rnnt_type, dp, "none",
                                    dtype=np.float64)
        assert_close(-scores.cpu().numpy(), o_pl, LOSS_RTOL, 1e-5, tag + " bf16 pruned loss")


@pytest.mark.parametrize("T,S,R,dp", [(500, 100, 5, 0.2), (700, 60, 3, 0.25), (400, 40, 8, 0.15), (650, 80, 4, 0.28)])
@pytest.mark.parametrize("rnnt_type", ["regular", "modified"])
def test_band_recursion_with_large_delay_penalty(monkeypatch, T, S, R, dp, rnnt_type):
    """The band recursion close to the delay-penalty spread at which frn_pruned_loss hands over to the dense
    kernels (band_delay_ok: 250-393 of 400 bits here, incl. the c3 setting): it is the band kernel that runs
    (4 launches against 6), and it agrees with the dense kernels."""
    import torch
    import tf_fast_rnnt as frn
    B, C = 2, 8
    am, lm, sym, term, bd = make_inputs(T + R, B, T, S, C, ragged=True)
    monkeypatch.delenv("FRN_BAND_DENSE", raising=False)
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
    logits = torch.from_numpy((am_p + lm_p).astype(np.float32)).cuda()
    lib = frn._lib.lib
    out, launches = [], []
    for dense in ("0", "1"):
        monkeypatch.setenv("FRN_BAND_DENSE", dense)
        n0 = lib.frn_kernel_launches()
        sc, gr = frn.pruned_loss_fwd_bwd(logits, sym, ranges, term, bd, rnnt_type, dp, None)
        launches.append(lib.frn_kernel_launches() - n0)
        out.append((sc.cpu().numpy(), gr.cpu().numpy()))
    assert launches == [4, 6], launches
    assert_close(out[0][0], out[1][0], 2e-6, 1e-4, "scores band vs dense")
    assert_close(out[0][1], out[1][1], 5e-5, 2e-6, "logits grad band vs dense")
    # beyond the limit the same call runs the dense kernels by itself
    monkeypatch.delenv("FRN_BAND_DENSE", raising=False)
    n0 = lib.frn_kernel_launches()
    frn.pruned_loss_fwd_bwd(logits, sym, ranges, term, bd, rnnt_type, 4.0 * dp, None)
    assert lib.frn_kernel_launches() - n0 == 6


@pytest.mark.parametrize("reduction", ["none", "mean", "sum"])
def test_reduce_and_reduce_pair(reduction):
    """frn_reduce / frn_reduce_pair (the reduction branches of rnnt_loss.py:327-338, 1121-1130): -scores, -sum,
    -mean, for one vector and for the two vectors of a step in one launch."""
    import torch
    import tf_fast_rnnt as frn
    lib, chk = frn._lib.lib, frn._lib.check
    code = {"none": 0, "mean": 1, "sum": 2}[reduction]
    rng = np.random.default_rng(3)
    for B in (1, 31, 32, 257, 1000):
        a = rng.standard_normal(B).astype(np.float32) * 100
        b = rng.standard_normal(B).astype(np.float32) * 100
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        n = B if reduction == "none" else 1
        oa, ob, oc = (torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3))
        st = torch.cuda.current_stream().cuda_stream
        chk(lib.frn_reduce(da.data_ptr(), B, code, 0.0, oc.data_ptr(), st), "reduce")
        chk(lib.frn_reduce_pair(da.data_ptr(), db.data_ptr(), B, code, 0.0, oa.data_ptr(), ob.data_ptr(), st), "pair")
        want = {"none": lambda x: -x, "mean": lambda x: np.array([-x.astype(np.float64).mean()]),
                "sum": lambda x: np.array([-x.astype(np.float64).sum()])}[reduction]
        assert torch.equal(oa, oc)
        np.testing.assert_allclose(oa.cpu().numpy(), want(a), rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(ob.cpu().numpy(), want(b), rtol=1e-5, atol=1e-3)


@pytest.mark.debug_hooks
@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_half_precision_am_lm_inputs(dtype):
    """ bf16 / fp16 am and lm on the device are consumed as they are (frn_simple_loss_lp /
    frn_simple_logprobs_lp: the row-statistics kernel widens every element in registers) - bit-identical to feeding
    the up-cast values, with NO extra kernel; shapes the tensor-core path cannot take (C % 4 != 0) are widened by
    frn_cast_to_f32 (two more kernels) and then take the float32 path."""
    import torch
    import tf_fast_rnnt as frn
    lib = frn._lib.lib
    td = getattr(torch, dtype)
    for C, extra in ((36, 0), (35, 2)):
        B, T, S = 2, 61, 17
        am, lm, sym, term, bd = make_inputs(12, B, T, S, C, ragged=True)
        am_h, lm_h = torch.from_numpy(am).cuda().to(td), torch.from_numpy(lm).cuda().to(td)
        for kind in ("simple", "smoothed", "modified", "logprobs"):
            def run(l, a):
                if kind == "simple":
                    loss, (gx, gy) = frn.rnnt_loss_simple(l, a, sym, term, bd, "regular", 0.1, "none", True)
                elif kind == "smoothed":
                    loss, (gx, gy) = frn.rnnt_loss_smoothed(l, a, sym, term, 0.2, 0.1, bd, "regular", 0.0, "none", True)
                elif kind == "modified":
                    loss, (gx, gy) = frn.rnnt_loss_simple(l, a, sym, term, bd, "modified", 0.0, "none", True)
                else:
                    gx, gy = frn.get_rnnt_logprobs(l, a, sym, term, "regular", bd)
                    loss = gx[:, 0, 0]
                return loss, gx, gy
            n0 = lib.frn_kernel_launches()
            loss_h, gx_h, gy_h = run(lm_h, am_h)
            launched = lib.frn_kernel_launches() - n0
            n1 = lib.frn_kernel_launches()
            loss_f, gx_f, gy_f = run(lm_h.float(), am_h.float())
            assert torch.equal(loss_h, loss_f) and torch.equal(gx_h, gx_f) and torch.equal(gy_h, gy_f), (C, kind)
            assert launched == (lib.frn_kernel_launches() - n1) + extra, (C, kind)
    # against the float64 oracle on the up-cast values ("upcast then reference math")
    am, lm, sym, term, bd = make_inputs(12, 2, 61, 17, 36, ragged=True)
    am_h, lm_h = torch.from_numpy(am).cuda().to(td), torch.from_numpy(lm).cuda().to(td)
    loss, (gx, gy) = frn.rnnt_loss_simple(lm_h, am_h, sym, term, bd, "regular", 0.0, "none", True)
    o_loss, (o_gx, o_gy) = orc.rnnt_loss_simple(lm_h.float().cpu().numpy(), am_h.float().cpu().numpy(), sym, term, bd,
                                                "regular", 0.0, "none", True, dtype=np.float64)
    assert_close(loss.cpu().numpy(), o_loss, LOSS_RTOL, 0, "loss of bf16 / fp16 inputs")
    assert_close(gx.cpu().numpy(), o_gx, GRAD_RTOL, GRAD_ATOL, "px_grad of bf16 / fp16 inputs")
    assert_close(gy.cpu().numpy(), o_gy, GRAD_RTOL, GRAD_ATOL, "py_grad of bf16 / fp16 inputs")
    # do_rnnt_pruning keeps the type of bf16 / fp16 inputs (rnnt_loss.py:802-811 are a broadcast and a gather)
    rng = np.random.default_rng(3)
    for C in (36, 34):
        am, lm, sym, term, bd = make_inputs(12, 2, 61, 17, C, ragged=True)
        am_h, lm_h = torch.from_numpy(am).cuda().to(td), torch.from_numpy(lm).cuda().to(td)
        rg = torch.from_numpy(np.sort(rng.integers(0, 14, (2, 61, 1)), axis=1).astype(np.int32) + np.arange(4, dtype=np.int32)).cuda()
        am_p, lm_p = frn.do_rnnt_pruning(am_h, lm_h, rg)
        assert am_p.dtype == td and lm_p.dtype == td
        ref_a, ref_l = frn.do_rnnt_pruning(am_h.float(), lm_h.float(), rg)
        assert torch.equal(am_p.float(), ref_a) and torch.equal(lm_p.float(), ref_l)
    # odd element counts take the scalar tail of the widening kernel
    x = torch.randn(1003, device="cuda").to(td)
    y = torch.empty(1003, dtype=torch.float32, device="cuda")
    frn._lib.check(lib.frn_cast_to_f32(x.data_ptr(), 1 if dtype == "bfloat16" else 2, 1003, y.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream), "cast")
    assert torch.equal(y, x.float())


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
def test_joint_loss_autograd(rnnt_type):
    """rnnt_loss on the full joiner output with CUDA logits that require grad: the logits gradient TF autodiff
    derives through rnnt_loss.py:340-551 (frn_joint_loss), against the float64 oracle."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C = 2, 37, 11, 6
    am, lm, sym, term, bd = make_inputs(51, B, T, S, C, ragged=True)
    full = (am[:, :, None, :] + lm[:, None, :, :]).astype(np.float32)
    lg = torch.from_numpy(full).cuda().requires_grad_(True)
    w = np.array([1.5, -0.25], np.float32)
    loss = frn.rnnt_loss(lg, sym, term, bd, rnnt_type, 0.15, "none")
    (loss * torch.from_numpy(w).cuda()).sum().backward()
    o_loss = orc.rnnt_loss(full, sym, term, bd, rnnt_type, 0.15, "none", dtype=np.float64)
    ranges = np.broadcast_to(np.arange(S + 1, dtype=np.int32)[None, None, :], (B, T, S + 1)).copy()
    o_grad = orc.pruned_logits_grad(full, sym, ranges, term, bd, rnnt_type, 0.15, w, np.float64)
    assert_close(loss.detach().cpu().numpy(), o_loss, LOSS_RTOL, 0, "joint loss")
    assert_close(lg.grad.cpu().numpy(), o_grad, GRAD_RTOL, 2e-6, "joint logits grad")
    lg.grad = None
    frn.rnnt_loss(lg, sym, term, bd, rnnt_type, 0.15, "mean").backward()
    o_mean = orc.pruned_logits_grad(full, sym, ranges, term, bd, rnnt_type, 0.15, np.full(B, 1.0 / B), np.float64)
    assert_close(lg.grad.cpu().numpy(), o_mean, GRAD_RTOL, 2e-6, "joint logits grad (mean)")


@pytest.mark.parametrize("fused_joiner", [False, True])
def test_training_step_gradients_end_to_end(fused_joiner):
    """A whole training step through the public API with autograd (what a TF user gets from GradientTape):
    loss = 0.5 * simple + pruned, gradients w.r.t. am and lm flow back through the pruned loss, the joiner,
    do_rnnt_pruning and the simple loss; checked against the float64 oracle composed by hand."""
    import torch
    import tf_fast_rnnt as frn
    B, T, S, C, R = 2, 45, 13, 12, 4
    am, lm, sym, term, bd = make_inputs(61, B, T, S, C, ragged=True)
    am_t = torch.from_numpy(am).cuda().requires_grad_(True)
    lm_t = torch.from_numpy(lm).cuda().requires_grad_(True)
    simple, (gx, gy) = frn.rnnt_loss_simple(lm_t, am_t, sym, term, bd, "regular", 0.0, "sum", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    if fused_joiner:
        _, _, logits = frn.do_rnnt_pruning_add_joiner(am_t, lm_t, ranges)
    else:
        am_p, lm_p = frn.do_rnnt_pruning(am_t, lm_t, ranges)
        logits = am_p + lm_p
    pruned = frn.rnnt_loss_pruned(logits, sym, ranges, term, bd, "regular", 0.0, "sum")
    (0.5 * simple + pruned).backward()
    # oracle: d simple / d(am, lm) + pruning-backward of d pruned / d logits
    rg = ranges.cpu().numpy()
    o_am, o_lm = orc.simple_am_lm_grad(lm, am, sym, term, bd, "regular", 0.0, None, np.float64)
    o_am_p, o_lm_p = orc.do_rnnt_pruning(am, lm, rg)
    o_lg = orc.pruned_logits_grad((o_am_p + o_lm_p).astype(np.float32), sym, rg, term, bd, "regular", 0.0, None, np.float64)
    p_am, p_lm = orc.do_rnnt_pruning_bwd(o_lg, o_lg, rg, S + 1)
    assert_close(am_t.grad.cpu().numpy(), 0.5 * o_am + p_am, GRAD_RTOL, 5e-6, "am grad")
    assert_close(lm_t.grad.cpu().numpy(), 0.5 * o_lm + p_lm, GRAD_RTOL, 5e-5, "lm grad")


def test_concurrent_streams_are_independent():
    """The library keeps no hidden device state: two pipelines enqueued on two CUDA streams at once (own
    workspaces) give the same bits as the same pipelines run one after the other."""
    import torch
    import tf_fast_rnnt as frn
    shapes = [(3, 120, 30, 24, 4), (2, 333, 41, 16, 5)]
    inputs = []
    for k, (B, T, S, C, R) in enumerate(shapes):
        am, lm, sym, term, bd = make_inputs(70 + k, B, T, S, C, ragged=True)
        inputs.append((torch.from_numpy(am).cuda(), torch.from_numpy(lm).cuda(), sym, term, bd, R))

    def pipeline(am, lm, sym, term, bd, R):
        loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.1, "none", True)
        ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
        am_p, lm_p, logits = frn.do_rnnt_pruning_add_joiner(am, lm, ranges)
        scores, grad = frn.pruned_loss_fwd_bwd(logits, sym, ranges, term, bd, "regular", 0.1, None)
        return loss, gx, gy, ranges, scores, grad

    serial = [pipeline(*inp) for inp in inputs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in inputs]
    for rep in range(3):
        conc = []
        for st, inp in zip(streams, inputs):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                conc.append(pipeline(*inp))
        torch.cuda.synchronize()
        for a, b in zip(serial, conc):
            for x, y in zip(a, b):
                assert torch.equal(x, y)


def test_fuzz_tensor_core_gradient_contraction_equals_simt(monkeypatch):
    """simple_bwd_tc.cu (tcgen05, MN-major / K-major swizzled operands) against the exact-FP32 SIMT tiles of
    simple_bwd.cu on shapes that exercise every tile edge: partial K slices, several M tiles on either side,
    partial and multiple N tiles, all rnnt types, smoothed and not."""
    import tf_fast_rnnt as frn
    rng = np.random.default_rng(21)
    shapes = [(2, 70, 21, 28), (1, 129, 5, 260), (2, 64, 130, 36), (1, 300, 200, 516), (3, 37, 11, 8), (1, 257, 63, 132)]
    for n, (B, T, S, C) in enumerate(shapes):
        rnnt_type = ["regular", "modified", "constrained"][n % 3]
        smoothed = n % 2 == 1
        am, lm, sym, term, bd = make_inputs(100 + n, B, T, S, C, ragged=True, begin=(n % 2 == 0))
        sg = rng.standard_normal(B).astype(np.float32)
        if smoothed:
            _, (gx, gy) = frn.rnnt_loss_smoothed(lm, am, sym, term, 0.2, 0.1, bd, rnnt_type, 0.0, "none", True)
        else:
            _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True)
        out = []
        for simt in ("0", "1"):
            monkeypatch.setenv("FRN_BWD_SIMT", simt)
            out.append(frn.simple_loss_backward(lm, am, sym, term, bd, gx, gy, sg, rnnt_type, smoothed, 0.2, 0.1))
        assert_close(out[0][0], out[1][0], 2e-5, 1e-6, f"am grad {B, T, S, C}")
        assert_close(out[0][1], out[1][1], 2e-5, 1e-6, f"lm grad {B, T, S, C}")
