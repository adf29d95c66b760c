"""Time frn_mi_fwd_bwd (lattice recursion, dense px/py) with CUDA events at a given shape.
FRN_DP_CHAIN=1 selects the wavefront chain kernels instead of the row-scan kernels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
B, T, S = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 500, 100)))
mod = len(sys.argv) > 4 and sys.argv[4] == "modified"
rng = np.random.default_rng(0)
T1 = T if mod else T + 1
px = torch.from_numpy((rng.standard_normal((B, S, T1)) - 6).astype(np.float32)).cuda()
py = torch.from_numpy((rng.standard_normal((B, S + 1, T)) - 0.5).astype(np.float32)).cuda()
bd = torch.tensor([[0, 0, S, T]] * B, dtype=torch.int32).cuda()
for _ in range(3):
    ans, (gx, gy) = frn.mutual_information_recursion(px, py, bd, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 50
e0.record()
for _ in range(n):
    ans, (gx, gy) = frn.mutual_information_recursion(px, py, bd, True)
e1.record()
torch.cuda.synchronize()
print("shape", B, T, S, "modified" if mod else "regular", "chain" if os.environ.get("FRN_DP_CHAIN") == "1" else "scan",
      "ms/call %.4f" % (e0.elapsed_time(e1) / n), "ans", ans[:2].tolist(),
      "sum gx %.6f sum gy %.6f" % (gx.double().sum().item() / B, gy.double().sum().item() / B))
