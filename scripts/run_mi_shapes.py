"""Run the lattice DP at a few (S, T) shapes (for ncu launch lists: cycles per step vs pipeline width)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
B = 32
for S, T in ((31, 500), (63, 500), (100, 500), (127, 500), (255, 500)):
    rng = np.random.default_rng(0)
    px = torch.from_numpy((rng.standard_normal((B, S, T + 1)) - 6).astype(np.float32)).cuda()
    py = torch.from_numpy((rng.standard_normal((B, S + 1, T)) - 0.5).astype(np.float32)).cuda()
    bd = torch.tensor([[0, 0, S, T]] * B, dtype=torch.int32).cuda()
    for _ in range(3):
        ans, _ = frn.mutual_information_recursion(px, py, bd, True)
    torch.cuda.synchronize()
    print(S, T, ans[:2].tolist())
