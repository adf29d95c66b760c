"""Run the lattice DP at a few (B, S, T) shapes (for ncu launch lists: cycles per step vs pipeline width).
usage: run_mi_shapes.py [B,S,T ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(32, 31, 500), (32, 63, 500), (32, 100, 500), (32, 127, 500), (32, 255, 500)]
for B, S, T in shapes:
    rng = np.random.default_rng(0)
    px = torch.from_numpy((rng.standard_normal((B, S, T + 1)) - 6).astype(np.float32)).cuda()
    py = torch.from_numpy((rng.standard_normal((B, S + 1, T)) - 0.5).astype(np.float32)).cuda()
    bd = torch.tensor([[0, 0, S, T]] * B, dtype=torch.int32).cuda()
    for _ in range(3):
        ans, _ = frn.mutual_information_recursion(px, py, bd, True)
    torch.cuda.synchronize()
    print(B, S, T, ans[:2].tolist())
