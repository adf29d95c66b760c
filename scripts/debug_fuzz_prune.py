"""Ad-hoc fuzz: get_rnnt_prune_ranges (bit-exact) and do_rnnt_pruning at larger shapes against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from oracle import rnnt_oracle as orc
from tests.helpers import random_pxpy
from tests.test_gpu_dp import _boundaries
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    modified = bool(case & 1)
    B = int(rng.integers(1, 4)); S = int(rng.integers(1, 450)); T = int(rng.integers(max(S // 3, 2), 1500))
    R = int(rng.integers(1, 12))
    px, py = random_pxpy(int(rng.integers(1 << 30)), B, S, T, modified)
    bd = _boundaries(rng, B, S, T, ["full", "ragged", "begin"][case % 3])
    _, (gx, gy) = frn.mutual_information_recursion(px, py, bd, calc_gradients=True)
    if case % 4 == 0:       # ties and zeros: quantise the occupation counts
        gx = np.round(gx * 8) / 8; gy = np.round(gy * 8) / 8
    r = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    o = orc.get_rnnt_prune_ranges(gx, gy, bd, R)
    ok = np.array_equal(r, o)
    C = 8
    am = rng.standard_normal((B, T, C), dtype=np.float32); lm = rng.standard_normal((B, S + 1, C), dtype=np.float32)
    a, l = frn.do_rnnt_pruning(am, lm, r)
    oa, ol = orc.do_rnnt_pruning(am, lm, o)
    ok2 = np.array_equal(a, oa) and np.array_equal(l, ol)
    bad += (not ok) + (not ok2)
    print(f"case {case:2d} B={B} S={S:3d} T={T:4d} R={R:2d} mod={int(modified)} ranges {'ok' if ok else 'DIFF ' + str(int((r != o).sum()))} pruning {'ok' if ok2 else 'DIFF'}")
print("bad", bad)
