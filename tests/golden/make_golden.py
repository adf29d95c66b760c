"""Generate tests/golden/*.npz by executing the REFERENCE's own rnnt_loss.py.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference file is imported unmodified, by path, on top of oracle/tf_emu.py
(a NumPy stand-in for the TensorFlow ops it calls).  Its two compiled custom ops
(lattice recursion, cummin) are served by the plain-C oracle, which is pinned
separately against the reference's CUDA kernels on the B200.  Input recipes
follow the reference's tests (simple_rnnt_loss_test.py:51-66, 274-289) and
README.md:85-152.  The resulting fixtures are committed; nothing at test time
reads /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tf_emu  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def recipe(seed, B, T, S, C):
    """Input recipe of the reference's tests (simple_rnnt_loss_test.py:274-289)."""
    np.random.seed(seed)
    frames = np.random.randint(S, T, (B,))
    seq_length = np.random.randint(3, S - 1, (B,))
    T = int(np.amax(frames))
    S = int(np.amax(seq_length))
    am = np.random.randn(B, T, C).astype("f")
    lm = np.random.randn(B, S + 1, C).astype("f")
    symbols = np.random.randint(0, C - 1, (B, S)).astype(np.int32)
    boundary = np.zeros((B, 4), dtype=np.int32)
    boundary[:, 2] = seq_length
    boundary[:, 3] = frames
    return am, lm, symbols, C - 1, boundary


def run_case(ref, name, seed, B, T, S, C, s_ranges):
    am, lm, symbols, term, boundary = recipe(seed, B, T, S, C)
    out = dict(am=am, lm=lm, symbols=symbols, boundary=boundary,
               termination_symbol=np.int32(term), s_ranges=np.asarray(s_ranges, np.int32))
    # A1 + A3 + A4
    px, py = ref.get_rnnt_logprobs(lm=lm, am=am, symbols=symbols,
                                   termination_symbol=term, boundary=boundary)
    out["simple_px"], out["simple_py"] = px, py
    for dp in (0.0, 0.2):
        loss, (gx, gy) = ref.rnnt_loss_simple(
            lm=lm, am=am, symbols=symbols, termination_symbol=term,
            boundary=boundary, rnnt_type="regular", calc_gradients=True,
            reduction="none", delay_penalty=dp)
        tag = f"dp{int(dp * 10)}"
        out[f"simple_loss_{tag}"] = loss
        out[f"simple_px_grad_{tag}"] = gx
        out[f"simple_py_grad_{tag}"] = gy
    out["simple_loss_sum"] = np.asarray(ref.rnnt_loss_simple(
        lm=lm, am=am, symbols=symbols, termination_symbol=term,
        boundary=boundary, reduction="sum"))
    # A2
    for lms, ams in ((0.1, 0.2), (0.25, 0.0)):
        tag = f"l{int(lms * 100)}_a{int(ams * 100)}"
        spx, spy = ref.get_rnnt_logprobs_smoothed(
            lm=lm, am=am, symbols=symbols, termination_symbol=term,
            lm_only_scale=lms, am_only_scale=ams, boundary=boundary)
        out[f"smoothed_px_{tag}"], out[f"smoothed_py_{tag}"] = spx, spy
        loss, (sgx, sgy) = ref.rnnt_loss_smoothed(
            lm=lm, am=am, symbols=symbols, termination_symbol=term,
            lm_only_scale=lms, am_only_scale=ams, boundary=boundary,
            rnnt_type="regular", calc_gradients=True, reduction="none",
            delay_penalty=0.2)
        out[f"smoothed_loss_{tag}"] = loss
        out[f"smoothed_px_grad_{tag}"] = sgx
        out[f"smoothed_py_grad_{tag}"] = sgy
    # A5..A8 driven by the dp=0.2 simple grads, as in the reference's stress test
    gx, gy = out["simple_px_grad_dp2"], out["simple_py_grad_dp2"]
    for r in s_ranges:
        ranges = ref.get_rnnt_prune_ranges(px_grad=gx, py_grad=gy,
                                           boundary=boundary, s_range=int(r))
        out[f"ranges_r{r}"] = ranges
        am_p, lm_p = ref.do_rnnt_pruning(am=am, lm=lm, ranges=ranges)
        logits = (1.0 / (1.0 + np.exp(-(am_p + lm_p)))).astype(np.float32)  # sigmoid joiner
        for rt in ("regular", "modified", "constrained"):
            ppx, ppy = ref.get_rnnt_logprobs_pruned(
                logits=logits, symbols=symbols, ranges=ranges,
                termination_symbol=term, boundary=boundary, rnnt_type=rt)
            out[f"pruned_px_r{r}_{rt}"], out[f"pruned_py_r{r}_{rt}"] = ppx, ppy
            for dp in (0.0, 0.2):
                out[f"pruned_loss_r{r}_{rt}_dp{int(dp * 10)}"] = ref.rnnt_loss_pruned(
                    logits=logits, symbols=symbols, ranges=ranges,
                    termination_symbol=term, boundary=boundary, rnnt_type=rt,
                    reduction="none", delay_penalty=dp)
        if r == s_ranges[0]:
            out["am_pruned_r%d" % r] = am_p
            out["lm_pruned_r%d" % r] = lm_p
    # (f) rnnt_loss on the full joiner
    full = (1.0 / (1.0 + np.exp(-(am[:, :, None, :] + lm[:, None, :, :])))).astype(np.float32)
    out["joint_loss_dp2"] = ref.rnnt_loss(
        logits=full, symbols=symbols, termination_symbol=term, boundary=boundary,
        rnnt_type="regular", reduction="none", delay_penalty=0.2)
    jpx, jpy = ref.get_rnnt_logprobs_joint(
        logits=full, symbols=symbols, termination_symbol=term, boundary=boundary)
    out["joint_px"], out["joint_py"] = jpx, jpy
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
    print(name, {k: np.asarray(v).shape for k, v in out.items() if k.endswith("loss_dp2") or k.startswith("ranges")},
          os.path.getsize(path) // 1024, "KiB")


def broken_in_reference(ref):
    """Record which documented options raise in the reference itself."""
    am, lm, symbols, term, boundary = recipe(1234, 2, 50, 10, 16)
    status = {}
    for rt in ("modified", "constrained"):
        try:
            ref.rnnt_loss_simple(lm=lm, am=am, symbols=symbols, termination_symbol=term,
                                 boundary=boundary, rnnt_type=rt, reduction="none")
            status[rt] = "ok"
        except Exception as e:  # noqa: BLE001
            status[rt] = type(e).__name__
    try:
        ref.rnnt_loss_simple(lm=lm, am=am, symbols=symbols, termination_symbol=term,
                             boundary=boundary, reduction="mean")
        status["mean"] = "ok"
    except Exception as e:  # noqa: BLE001
        status["mean"] = type(e).__name__
    return status


if __name__ == "__main__":
    ref = tf_emu.load_reference()
    print("reference defects reproduced:", broken_in_reference(ref))
    # c1: README example shape (README.md:85-152), seed of simple_rnnt_loss_test.py:73
    run_case(ref, "c1_readme", 1234, 2, 50, 10, 16, [2, 3, 5, 20])
    # the reference's live stress test (simple_rnnt_loss_test.py:256-369)
    run_case(ref, "stress_b2_t200_s50_c50", 12345, 2, 200, 50, 50, [5, 7, 50])
