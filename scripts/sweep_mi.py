"""Lattice recursion (frn_mi_fwd_bwd): row-scan kernels against wavefront chain kernels over a grid of
shapes; run once with FRN_DP_CHAIN=1 and once without."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
tag = "chain" if os.environ.get("FRN_DP_CHAIN") == "1" else "scan"
rng = np.random.default_rng(0)
for B, T, S in [(32, 500, 20), (32, 500, 50), (32, 500, 100), (32, 500, 200), (32, 500, 400), (32, 200, 100),
                (32, 1000, 100), (32, 1000, 250), (16, 1500, 100), (16, 1500, 400), (64, 500, 100), (128, 500, 100)]:
    px = torch.from_numpy((rng.standard_normal((B, S, T + 1)) - 6).astype(np.float32)).cuda()
    py = torch.from_numpy((rng.standard_normal((B, S + 1, T)) - 0.5).astype(np.float32)).cuda()
    bd = torch.tensor([[0, 0, S, T]] * B, dtype=torch.int32).cuda()
    for _ in range(3):
        frn.mutual_information_recursion(px, py, bd, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        frn.mutual_information_recursion(px, py, bd, True)
    e1.record()
    torch.cuda.synchronize()
    print(tag, "B %d T %d S %d  ms/call %.4f" % (B, T, S, e0.elapsed_time(e1) / 30), flush=True)
