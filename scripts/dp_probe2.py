import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from oracle import rnnt_oracle as orc
from tests.helpers import random_pxpy
from tests.test_gpu_dp import _boundaries

def case(modified, shape, kind, seed):
    B, S, T = shape
    rng = np.random.default_rng([int(modified), *shape, len(kind), seed])
    px, py = random_pxpy(rng.integers(1 << 30), B, S, T, modified)
    bd = _boundaries(rng, B, S, T, kind)
    return px, py, bd
import itertools
worst = {}
for modified, shape, kind, seed in itertools.product([False, True], [(4, 33, 70), (2, 127, 129), (5, 64, 64), (2, 130, 200)], ["full", "ragged", "begin"], range(6)):
    px, py, bd = case(modified, shape, kind, seed)
    ans, (gx, gy) = frn.mutual_information_recursion(px, py, bd, True)
    a64, (gx64, gy64) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    ok = np.isfinite(a64)
    for nm, g, g64 in (("gx", gx, gx64), ("gy", gy, gy64)):
        d = np.abs(g[ok] - g64[ok]); i = np.unravel_index(np.argmax(d), d.shape)
        bi = np.flatnonzero(ok)[i[0]]
        key = (modified, shape, nm); worst[key] = max(worst.get(key, 0), d.max())
        if d.max() > 5e-5: print(f"mod={modified} {shape} {kind}: {nm} max abs {d.max():.2e} at b={bi} s={i[1]} t={i[2]} truth {g64[ok][i]:.4f} bd={bd[bi]} score {a64[bi]:.3f} err {abs(ans[bi]-a64[bi]):.2e}")

for k, v in worst.items(): print(k, f"{v:.2e}")
print("LIB", frn.LIB_PATH)
