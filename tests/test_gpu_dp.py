"""GPU parity of the lattice recursion (A4) and cummin: CUDA path through the
C ABI vs the float64 oracle, and the oracle vs the reference's own CUDA kernels
(oracle/_ref) — the run that pins the oracle."""
import numpy as np
import pytest

from oracle import rnnt_oracle as orc
from tests.helpers import (GRAD_ATOL, GRAD_RTOL, LOSS_RTOL, RefKernels, assert_close,
                           load_golden, random_pxpy)

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["chain", "scan"], autouse=True)
def dp_kernel(request, monkeypatch):
    """Every test of this module runs with both dense-lattice recursions: the wavefront chain
    (mi_dp.cu) and the row scan (mi_scan.cu); the library picks by shape otherwise."""
    monkeypatch.delenv("FRN_DP_CHAIN", raising=False)
    monkeypatch.delenv("FRN_DP_SCAN", raising=False)
    monkeypatch.setenv("FRN_DP_CHAIN" if request.param == "chain" else "FRN_DP_SCAN", "1")
    return request.param


def _boundaries(rng, B, S, T, kind):
    bd = np.zeros((B, 4), np.int32)
    bd[:, 2], bd[:, 3] = S, T
    if kind in ("ragged", "begin"):
        bd[:, 3] = rng.integers(max(1, T // 2), T + 1, B)
        bd[:, 2] = np.minimum(rng.integers(0, S + 1, B), bd[:, 3])
        bd[0, 2], bd[0, 3] = S, T
    if kind == "begin":
        bd[:, 0] = np.minimum(rng.integers(0, 3, B), bd[:, 2])
        bd[:, 1] = np.minimum(rng.integers(0, 4, B), bd[:, 3])
    return bd


@pytest.mark.parametrize("modified", [False, True])
@pytest.mark.parametrize("shape", [(2, 10, 50), (3, 1, 1), (4, 33, 70), (2, 127, 129), (2, 130, 300),
                                   (1, 300, 40), (5, 64, 64)])
@pytest.mark.parametrize("kind", ["full", "ragged", "begin"])
def test_mi_against_float64_oracle(modified, shape, kind):
    import tf_fast_rnnt
    B, S, T = shape
    rng = np.random.default_rng([int(modified), *shape, len(kind)])
    px, py = random_pxpy(rng.integers(1 << 30), B, S, T, modified)
    bd = _boundaries(rng, B, S, T, kind)
    ans, (gx, gy) = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
    ans64, (gx64, gy64) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    # scores here are sums of ~(S+T) O(1) terms of both signs that may cancel to ~0,
    # hence the absolute term (1e-5 in log-probability)
    assert_close(ans, ans64, LOSS_RTOL, 1e-5, "ans")
    ok = np.isfinite(ans64)      # no path (score -inf): the reference's grads are meaningless there
    assert_close(gx[ok], gx64[ok], GRAD_RTOL, GRAD_ATOL, "px_grad")
    assert_close(gy[ok], gy64[ok], GRAD_RTOL, GRAD_ATOL, "py_grad")
    assert not np.isnan(gx).any() and not np.isnan(gy).any()
    only = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=False)
    assert np.array_equal(only, ans)


@pytest.mark.parametrize("modified", [False, True])
@pytest.mark.parametrize("cluster,cols", [(1, 2), (1, 4), (2, 1), (2, 2), (4, 1), (8, 1), (8, 4)])
def test_row_scan_variants(monkeypatch, dp_kernel, modified, cluster, cols):
    """Row-scan kernel with the chain of warps spread over a thread-block cluster and with
    several lattice columns per thread (the variants long / many utterances select)."""
    if dp_kernel != "scan":
        pytest.skip("row-scan variants only")
    import tf_fast_rnnt
    monkeypatch.setenv("FRN_SCAN_CLUSTER", str(cluster))
    monkeypatch.setenv("FRN_SCAN_K", str(cols))
    B, S, T = 3, 45, 900
    rng = np.random.default_rng([int(modified), cluster, cols])
    px, py = random_pxpy(rng.integers(1 << 30), B, S, T, modified)
    bd = _boundaries(rng, B, S, T, "begin")
    bd[1] = [2, 3, 30, 333]          # ends inside the first CTA of a cluster
    ans, (gx, gy) = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
    ans64, (gx64, gy64) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    assert_close(ans, ans64, LOSS_RTOL, 1e-5, "ans")
    assert_close(gx, gx64, GRAD_RTOL, GRAD_ATOL, "px_grad")
    assert_close(gy, gy64, GRAD_RTOL, GRAD_ATOL, "py_grad")


def test_mi_with_minus_inf_inputs():
    """px/py holding -inf (what get_rnnt_logprobs produces at t_end) and an
    unreachable end state."""
    import tf_fast_rnnt
    B, S, T = 3, 6, 9
    px, py = random_pxpy(7, B, S, T, False)
    px[:, :, T] = -np.inf
    px[1, 2, :] = -np.inf          # utterance 1 cannot pass symbol 2: score = -inf
    py[2, 3, 4] = -np.inf
    bd = np.tile(np.array([0, 0, S, T], np.int32), (B, 1))
    ans, (gx, gy) = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
    ans64, (gx64, gy64) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    assert np.isneginf(ans[1]) and np.isneginf(ans64[1])
    assert_close(ans[[0, 2]], ans64[[0, 2]], LOSS_RTOL, 1e-6, "ans")
    assert_close(gx[[0, 2]], gx64[[0, 2]], GRAD_RTOL, GRAD_ATOL, "px_grad")
    assert_close(gy[[0, 2]], gy64[[0, 2]], GRAD_RTOL, GRAD_ATOL, "py_grad")
    assert not np.isnan(gx).any() and not np.isnan(gy).any()


@pytest.mark.parametrize("modified", [False, True])
def test_oracle_pinned_by_reference_kernels(modified):
    """The plain-C oracle (float32) against the REFERENCE's CUDA kernels."""
    ref = RefKernels()
    rng = np.random.default_rng(11 + modified)
    for (B, S, T) in [(2, 10, 50), (4, 33, 70), (2, 100, 200)]:
        px, py = random_pxpy(rng.integers(1 << 30), B, S, T, modified)
        bd = _boundaries(rng, B, S, T, "ragged")
        if modified:
            bd[:, 2] = np.minimum(bd[:, 2], bd[:, 3])
        r_ans, r_gx, r_gy, r_ag = ref.fast_rnnt_loss(px, py, bd)
        o_ans, (o_gx, o_gy) = orc.mutual_information_recursion(px, py, bd, True, np.float32)
        assert_close(o_ans, r_ans, 2e-6, 1e-5, "oracle ans vs reference kernel")
        assert_close(o_gx, r_gx, 2e-4, 1e-6, "oracle px_grad vs reference kernel")
        assert_close(o_gy, r_gy, 2e-4, 1e-6, "oracle py_grad vs reference kernel")
        # the reference's built-in self check: p_grad[s_begin,t_begin] == ans_grad == 1
        np.testing.assert_allclose(r_ag, 1.0, rtol=1e-3)


def test_cuda_path_against_reference_kernels():
    import tf_fast_rnnt
    ref = RefKernels()
    g = load_golden("stress_b2_t200_s50_c50")
    px, py, bd = g["simple_px"], g["simple_py"], g["boundary"]
    r_ans, r_gx, r_gy, _ = ref.fast_rnnt_loss(px, py, bd)
    ans, (gx, gy) = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
    assert_close(ans, r_ans, LOSS_RTOL, 0, "ans vs reference kernel")
    # the reference itself is float32 on |p| ~ 1e3 here: 1e-3 is its own accuracy
    assert_close(gx, r_gx, 2e-3, 1e-6, "px_grad vs reference kernel")
    assert_close(gy, r_gy, 2e-3, 1e-6, "py_grad vs reference kernel")


def test_cummin_bit_exact():
    import tf_fast_rnnt
    ref = RefKernels()
    rng = np.random.default_rng(5)
    for rows, n in [(1, 1), (2, 31), (3, 32), (5, 33), (32, 500), (7, 1500)]:
        x = rng.integers(-1000, 1000, (rows, n)).astype(np.int32)
        out = tf_fast_rnnt.cummin(x)
        assert out.dtype == np.int32
        assert np.array_equal(out, np.minimum.accumulate(x, axis=1))
        assert np.array_equal(out, orc.cummin(x))
        assert np.array_equal(out, ref.cummin(x))


@pytest.mark.parametrize("rnnt_type", ["regular", "modified", "constrained"])
@pytest.mark.parametrize("R", [1, 3, 5, 8])
@pytest.mark.parametrize("jumpy", [False, True])
def test_band_recursion_equals_dense_lattice(rnnt_type, R, jumpy):
    """frn_band_mi_fwd_bwd (the recursion on the [B,T,R] band, what rnnt_loss_pruned runs) against the
    float64 oracle on the dense lattice the reference builds from the band (rnnt_loss.py:968-1018):
    same scores, and band occupation counts = dense occupation counts gathered back onto the band.
    `jumpy` bands move up by R-1 rows in single frames (every row but one leaves the band)."""
    import torch
    from tf_fast_rnnt import _lib
    lib = _lib.lib
    B, S, T = 3, 40, 90
    rng = np.random.default_rng([R, int(jumpy), len(rnnt_type)])
    bd = np.array([[0, 0, S, T], [0, 0, S - 7, T - 11], [0, 3, S - 1, T - 1]], np.int32)[:B]
    # bands as get_rnnt_prune_ranges guarantees them: monotone, starting at s_begin, steps < R (one
    # utterance deliberately violates that when R == 1), parked at s_end - R + 1 from frame t_end - 1 on
    r0 = np.zeros((B, T), np.int64)
    for b in range(B):
        t_b, s_e, t_e = bd[b, 1], bd[b, 2], bd[b, 3]
        target = max(s_e - R + 1, 0)
        step = max(R - 1, 1) if jumpy else 1
        njump = -(-target // step)
        frames = np.arange(t_b + 1, t_e - 1)
        inc = np.zeros(T, np.int64)
        if njump <= len(frames):
            inc[rng.choice(frames, njump, replace=False)] = step
        else:
            inc[frames] = step
        r0[b] = np.minimum(np.cumsum(inc), target)
        r0[b, t_e - 1:] = target
    ranges = (r0[:, :, None] + np.arange(R)[None, None, :]).astype(np.int32)
    pxc = (rng.standard_normal((B, T, R)) * 2 - 3).astype(np.float32)
    pyc = (rng.standard_normal((B, T, R)) * 2 - 1).astype(np.float32)
    rt = _lib.RNNT_TYPES[rnnt_type]
    dev = torch.device("cuda")
    d = lambda a: torch.from_numpy(a).to(dev)
    pxc_d, pyc_d, rg_d, bd_d = d(pxc), d(pyc), d(ranges), d(bd)
    n = lib.frn_band_mi_workspace_bytes(B, S, T, R)
    assert n > 0
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    ans = torch.empty(B, device=dev)
    gxc, gyc = torch.empty(B, T, R, device=dev), torch.empty(B, T, R, device=dev)
    _lib.check(lib.frn_band_mi_fwd_bwd(pxc_d.data_ptr(), pyc_d.data_ptr(), rg_d.data_ptr(), bd_d.data_ptr(), B, S, T, R,
                                       rt, 0.0, 1, ans.data_ptr(), gxc.data_ptr(), gyc.data_ptr(), ws.data_ptr(), n,
                                       torch.cuda.current_stream().cuda_stream), "band_mi")
    # dense lattice from the band, in numpy (the scatter of get_rnnt_logprobs_pruned)
    T1 = T + 1 if rnnt_type == "regular" else T
    px = np.full((B, S, T1), -np.inf, np.float32)
    py = np.full((B, S + 1, T), -np.inf, np.float32)
    for b in range(B):
        for t in range(T):
            for i in range(R):
                s = ranges[b, t, i]
                if s <= S:
                    py[b, s, t] = pyc[b, t, i]
                if s < S:
                    v = pxc[b, t, i]
                    if rnnt_type == "constrained":
                        v = v + (pyc[b, t, i + 1] if i + 1 < R else -np.inf)
                    px[b, s, t] = v
        if rnnt_type == "regular":
            px[b, :, bd[b, 3]] = -np.inf
    o_ans, (o_gx, o_gy) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    assert_close(ans.cpu().numpy(), o_ans, LOSS_RTOL, 1e-5, "band score")
    ok = np.isfinite(o_ans)
    gxc, gyc = gxc.cpu().numpy(), gyc.cpu().numpy()
    for b in np.nonzero(ok)[0]:
        for t in range(T):
            for i in range(R):
                s = ranges[b, t, i]
                want_x = o_gx[b, s, t] if s < S else 0.0
                want_y = o_gy[b, s, t] if s <= S else 0.0
                if rnnt_type == "constrained" and i >= 1:
                    want_y += o_gx[b, s - 1, t]        # that px arc borrowed this py entry
                assert abs(gxc[b, t, i] - want_x) <= GRAD_ATOL + GRAD_RTOL * abs(want_x), (b, t, i, "px")
                assert abs(gyc[b, t, i] - want_y) <= GRAD_ATOL + GRAD_RTOL * abs(want_y), (b, t, i, "py")
    only = torch.empty(B, device=dev)
    _lib.check(lib.frn_band_mi_fwd_bwd(pxc_d.data_ptr(), pyc_d.data_ptr(), rg_d.data_ptr(), bd_d.data_ptr(), B, S, T, R,
                                       rt, 0.0, 0, only.data_ptr(), None, None, ws.data_ptr(), n,
                                       torch.cuda.current_stream().cuda_stream), "band_mi")
    assert torch.equal(only, ans)


def test_row_scan_equals_wavefront_on_random_lattices(monkeypatch, dp_kernel):
    """Fuzz: the two dense-lattice recursions (independent kernels, different summation orders) agree on
    random shapes, boundaries with offsets, both recursion types and inputs with -inf arcs."""
    if dp_kernel != "scan":
        pytest.skip("runs once")
    import tf_fast_rnnt
    rng = np.random.default_rng(2024)
    for case in range(48):
        modified = bool(case & 1)
        B = int(rng.integers(1, 5))
        S = int(rng.integers(1, 70))
        T = int(rng.integers(max(S // 8, 1), 750))
        px, py = random_pxpy(int(rng.integers(1 << 30)), B, S, T, modified)
        if case % 3 == 0:                       # a few dead arcs
            px[rng.random(px.shape) < 0.02] = -np.inf
            py[rng.random(py.shape) < 0.01] = -np.inf
        bd = _boundaries(rng, B, S, T, ["full", "ragged", "begin"][case % 3])
        res = {}
        for which in ("FRN_DP_CHAIN", "FRN_DP_SCAN"):
            monkeypatch.delenv("FRN_DP_CHAIN", raising=False)
            monkeypatch.delenv("FRN_DP_SCAN", raising=False)
            monkeypatch.setenv(which, "1")
            res[which] = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
        (a0, (gx0, gy0)), (a1, (gx1, gy1)) = res["FRN_DP_CHAIN"], res["FRN_DP_SCAN"]
        tag = f"case {case}: B={B} S={S} T={T} modified={modified}"
        assert np.array_equal(np.isfinite(a0), np.isfinite(a1)), tag
        ok = np.isfinite(a0)
        assert_close(a1[ok], a0[ok], 2e-6, 1e-5, tag + " ans")
        assert_close(gx1[ok], gx0[ok], 2e-5, 1e-6, tag + " px_grad")
        assert_close(gy1[ok], gy0[ok], 2e-5, 1e-6, tag + " py_grad")
