"""Ad-hoc fuzz: tensor-core normaliser against the exact-FP32 SIMT kernel at large shapes (many k slices,
several symbol tiles), simple and smoothed."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from tests.helpers import make_inputs
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    rnnt_type = ["regular", "modified", "constrained"][case % 3]
    B = int(rng.integers(1, 3)); S = int(rng.integers(100, 450)); T = int(rng.integers(200, 1500)); C = 4 * int(rng.integers(100, 1300))
    am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True)
    am *= float(rng.uniform(0.5, 4.0)); lm *= float(rng.uniform(0.5, 4.0))
    out = []
    for simt in ("0", "1"):
        os.environ["FRN_SIMPLE_SIMT"] = simt
        if case % 2:
            out.append(frn.get_rnnt_logprobs_smoothed(lm, am, sym, term, 0.25, 0.1, bd, rnnt_type))
        else:
            out.append(frn.get_rnnt_logprobs(lm, am, sym, term, rnnt_type, bd))
    f = np.isfinite(out[1][0]); g = np.isfinite(out[1][1])
    same_inf = np.array_equal(np.isfinite(out[0][0]), f) and np.array_equal(np.isfinite(out[0][1]), g)
    ex = np.abs(out[0][0][f] - out[1][0][f]).max(); ey = np.abs(out[0][1][g] - out[1][1][g]).max()
    flag = "  <<<<" if (ex > 2e-5 or ey > 2e-5 or not same_inf) else ""
    bad += bool(flag)
    print(f"case {case:2d} {rnnt_type:11s} B={B} S={S} T={T} C={C} smoothed={case % 2} max|px diff| {ex:.1e} max|py diff| {ey:.1e}{flag}")
print("bad", bad)
