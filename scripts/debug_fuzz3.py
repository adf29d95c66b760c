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
def run(tag, px, py, bd):
    o_ans, (o_gx, o_gy) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    for rep in range(2):
        a, (gx, gy) = tf_fast_rnnt.mutual_information_recursion(px, py, bd, calc_gradients=True)
        print(tag, "rep", rep, "ans", a, "per-utt max gx err", np.abs(gx - o_gx).max(axis=(1, 2)), "gy err", np.abs(gy - o_gy).max(axis=(1, 2)))
run("orig ", px, py, bd)
run("swap ", px[::-1].copy(), py[::-1].copy(), bd[::-1].copy())
run("dup0 ", np.repeat(px[:1], 2, 0), np.repeat(py[:1], 2, 0), np.repeat(bd[:1], 2, 0))
run("dup1 ", np.repeat(px[1:], 2, 0), np.repeat(py[1:], 2, 0), np.repeat(bd[1:], 2, 0))
run("one0 ", px[:1], py[:1], bd[:1])
run("one1 ", px[1:], py[1:], bd[1:])
print("dead counts", [(~np.isfinite(px[i])).sum() for i in range(2)], [(~np.isfinite(py[i])).sum() for i in range(2)])
