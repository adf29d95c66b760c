"""Run get_rnnt_logprobs a few times at the c2 shape (for ncu captures)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
B, T, S, C = 32, 500, 100, 500
rng = np.random.default_rng(0)
am = torch.from_numpy(rng.standard_normal((B, T, C), dtype=np.float32)).cuda()
lm = torch.from_numpy(rng.standard_normal((B, S + 1, C), dtype=np.float32)).cuda()
sym = torch.from_numpy(rng.integers(0, C - 1, (B, S)).astype(np.int32)).cuda()
bd = torch.tensor([[0, 0, S, T]] * B, dtype=torch.int32).cuda()
for _ in range(4):
    px, py = frn.get_rnnt_logprobs(lm, am, sym, C - 1, "regular", bd)
torch.cuda.synchronize()
print(float(py[0, 0, 0]))
