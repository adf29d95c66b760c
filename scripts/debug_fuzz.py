"""Reproduce a case of tests/test_gpu_dp.py::test_row_scan_equals_wavefront_on_random_lattices and compare
both recursion kernels with the float64 oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt
from oracle import rnnt_oracle as orc
from tests.helpers import random_pxpy
from tests.test_gpu_dp import _boundaries
want = int(sys.argv[1]) if len(sys.argv) > 1 else 21
rng = np.random.default_rng(2024)
for case in range(48):
    modified = bool(case & 1)
    B = int(rng.integers(1, 5)); S = int(rng.integers(1, 70)); T = int(rng.integers(max(S // 8, 1), 750))
    px, py = random_pxpy(int(rng.integers(1 << 30)), B, S, T, modified)
    if case % 3 == 0:
        px[rng.random(px.shape) < 0.02] = -np.inf
        py[rng.random(py.shape) < 0.01] = -np.inf
    bd = _boundaries(rng, B, S, T, ["full", "ragged", "begin"][case % 3])
    if case != want:
        continue
    o_ans, (o_gx, o_gy) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    print("case", case, "B S T", B, S, T, "modified", modified, "oracle ans", o_ans)
    for which in ("FRN_DP_CHAIN", "FRN_DP_SCAN"):
        os.environ.pop("FRN_DP_CHAIN", None); os.environ.pop("FRN_DP_SCAN", None)
        os.environ[which] = "1"
        a, (gx, gy) = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
        ok = np.isfinite(o_ans)
        print(which, "ans", a, "max|gx-o|", np.abs(gx[ok] - o_gx[ok]).max(), "max|gy-o|", np.abs(gy[ok] - o_gy[ok]).max(),
              "sum gx", gx.sum(axis=(1, 2)), "oracle sum gx", o_gx.sum(axis=(1, 2)))
        bad = np.argwhere(np.abs(gx - o_gx) > 1e-3)
        print("  first bad gx idx", bad[:5].tolist(), "n bad", len(bad))
