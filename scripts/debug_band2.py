import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from oracle import rnnt_oracle as orc
B, T, S, C, R = 1, 40, 75, 6, 5
rng = np.random.default_rng(0)
sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
bd = np.array([[0, 0, S, T]], np.int32)
r0 = np.minimum(np.arange(T) * 4, S - R + 1)
for C in (6, 8):
    ranges = (r0[None, :, None] + np.arange(R)[None, None, :]).astype(np.int32)
    logits = rng.standard_normal((B, T, R, C)).astype(np.float32)
    scores, grad = frn.pruned_loss_fwd_bwd(torch.from_numpy(logits).cuda(), sym % (C - 1), ranges, C - 1, bd, "regular", 0.0, -np.ones(B, np.float32))
    o_grad, o_scores = orc.pruned_logits_grad(logits, sym % (C - 1), ranges, C - 1, bd, "regular", 0.0, np.ones(B), np.float64, True)
    g = grad.cpu().numpy()
    err = np.abs(g - o_grad).max(axis=3)[0]
    print("C", C, "scores", scores.cpu().numpy(), o_scores, "bad (t,i):", np.argwhere(err > 1e-4).tolist()[:30])
