import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt
from oracle import rnnt_oracle as orc
from tests.helpers import random_pxpy
from tests.test_gpu_dp import _boundaries
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 2024)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 48
for case in range(n):
    modified = bool(case & 1)
    if os.environ.get("FUZZ_BIG"):
        B = int(rng.integers(1, 4)); S = int(rng.integers(120, 700)); T = int(rng.integers(max(S // 2, 1), 1500))
    else:
        B = int(rng.integers(1, 5)); S = int(rng.integers(1, 70)); T = int(rng.integers(max(S // 8, 1), 750))
    px, py = random_pxpy(int(rng.integers(1 << 30)), B, S, T, modified)
    if os.environ.get("FUZZ_DP"):
        tt = np.arange(px.shape[2], dtype=np.float64)
        px = (px + ((T - 1) / 2.0 - tt)[None, None, :] * float(os.environ["FUZZ_DP"])).astype(np.float32)
    dead = case % 3 == 0
    if dead:
        px[rng.random(px.shape) < 0.02] = -np.inf
        py[rng.random(py.shape) < 0.01] = -np.inf
    kind = ["full", "ragged", "begin"][case % 3]
    bd = _boundaries(rng, B, S, T, kind)
    o_ans, (o_gx, o_gy) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    ok = np.isfinite(o_ans)
    out = []
    for which in ("FRN_DP_CHAIN", "FRN_DP_SCAN"):
        os.environ.pop("FRN_DP_CHAIN", None); os.environ.pop("FRN_DP_SCAN", None)
        os.environ[which] = "1"
        a, (gx, gy) = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
        fin_ok = np.array_equal(np.isfinite(a), ok)
        e = max(np.abs(gx[ok] - o_gx[ok]).max(initial=0), np.abs(gy[ok] - o_gy[ok]).max(initial=0))
        ea = np.abs(a[ok] - o_ans[ok]).max(initial=0)
        out.append((fin_ok, ea, e))
    flag = "  <<<<" if (out[0][2] > 1e-4 or out[1][2] > 1e-4 or not out[0][0] or not out[1][0]) else ""
    print(f"case {case:2d} B={B} S={S:2d} T={T:3d} mod={int(modified)} dead={int(dead)} {kind:6s} n_ok={int(ok.sum())} "
          f"chain: fin={out[0][0]} ans {out[0][1]:.1e} grad {out[0][2]:.1e} | scan: fin={out[1][0]} ans {out[1][1]:.1e} grad {out[1][2]:.1e}{flag}")
