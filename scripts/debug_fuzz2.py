import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
os.environ["FRN_DP_CHAIN"] = "1"
import tf_fast_rnnt
from oracle import rnnt_oracle as orc
from tests.helpers import random_pxpy
from tests.test_gpu_dp import _boundaries
rng = np.random.default_rng(2024)
for case in range(22):
    modified = bool(case & 1)
    B = int(rng.integers(1, 5)); S = int(rng.integers(1, 70)); T = int(rng.integers(max(S // 8, 1), 750))
    px, py = random_pxpy(int(rng.integers(1 << 30)), B, S, T, modified)
    if case % 3 == 0:
        px[rng.random(px.shape) < 0.02] = -np.inf
        py[rng.random(py.shape) < 0.01] = -np.inf
    bd = _boundaries(rng, B, S, T, ["full", "ragged", "begin"][case % 3])
px, py, bd = px[:1], py[:1], bd[:1]
np.savez("gpurun_out/case21.npz", px=px, py=py, bd=bd)
o_ans, (o_gx, o_gy) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
a, (gx, gy) = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
print("ans", a, o_ans)
ex = np.abs(gx[0] - o_gx[0]) > 1e-4
ey = np.abs(gy[0] - o_gy[0]) > 1e-4
print("bad gx per row s:", ex.sum(axis=1).tolist())
print("bad gy per row s:", ey.sum(axis=1).tolist())
tb = np.where(ex.any(axis=0))[0]; print("bad gx t range", tb.min(), tb.max(), "count", len(tb))
tb = np.where(ey.any(axis=0))[0]; print("bad gy t range", tb.min(), tb.max(), "count", len(tb))
print("gy col sums (should be 1 for t<T): first bad", np.where(np.abs(gy[0].sum(axis=0) - 1) > 1e-3)[0][:10], "oracle", o_gy[0].sum(axis=0)[:3])
# shrink T from the right by boundary
for t_end in (342, 300, 200, 100, 50, 20):
    b2 = bd.copy(); b2[0, 3] = t_end
    o2, (ogx2, ogy2) = orc.mutual_information_recursion(px, py, b2, True, np.float64)
    a2, (gx2, gy2) = tf_fast_rnnt.mutual_information_recursion(px, py, b2, calc_gradients=True)
    print("t_end", t_end, "ans", float(a2[0]), float(o2[0]), "max gx err", np.abs(gx2 - ogx2).max(), "gy err", np.abs(gy2 - ogy2).max())
# which dead arcs matter: restore them one kind at a time
for name in ("px", "py"):
    p2, q2 = px.copy(), py.copy()
    arr = p2 if name == "px" else q2
    arr[~np.isfinite(arr)] = -5.0
    o2, (ogx2, ogy2) = orc.mutual_information_recursion(p2, q2, bd, True, np.float64)
    a2, (gx2, gy2) = tf_fast_rnnt.mutual_information_recursion(p2, q2, bd, calc_gradients=True)
    print("finite", name, "-> max gx err", np.abs(gx2 - ogx2).max())
