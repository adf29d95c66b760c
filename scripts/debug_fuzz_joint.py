"""Ad-hoc fuzz: rnnt_loss on the full joiner output (frn_joint_loss: identity band -> dense-lattice kernels,
1 / 2 / 4 lattice rows per lane) against the float64 oracle, loss and logits gradient."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from oracle import rnnt_oracle as orc
from tests.helpers import make_inputs
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
    rnnt_type = ["regular", "modified", "constrained"][case % 3]
    B = int(rng.integers(1, 3)); S = int(rng.integers(1, 360)); T = int(rng.integers(max(S, 2), 420)); C = int(rng.integers(2, 6))
    am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=bool(case % 4 == 0))
    dp = [0.0, 0.1][case % 2]
    full = (am[:, :, None, :] + lm[:, None, :, :]).astype(np.float32)
    lg = torch.from_numpy(full).cuda().requires_grad_(True)
    loss = frn.rnnt_loss(lg, sym, term, bd, rnnt_type, dp, "none")
    w = torch.from_numpy(rng.standard_normal(B).astype(np.float32)).cuda()
    (loss * w).sum().backward()
    o_loss = orc.rnnt_loss(full, sym, term, bd, rnnt_type, dp, "none", dtype=np.float64)
    ranges = np.broadcast_to(np.arange(S + 1, dtype=np.int32)[None, None, :], (B, T, S + 1)).copy()
    o_grad = orc.pruned_logits_grad(full, sym, ranges, term, bd, rnnt_type, dp, w.cpu().numpy(), np.float64)
    ok = np.isfinite(o_loss)
    el = np.abs(loss.detach().cpu().numpy()[ok] - o_loss[ok]).max(initial=0) / max(1.0, np.abs(o_loss[ok]).max(initial=1))
    eg = np.abs(lg.grad.cpu().numpy()[ok] - o_grad[ok]).max(initial=0)
    flag = "  <<<<" if (el > 1e-5 or eg > 1e-4 or not np.array_equal(np.isfinite(loss.detach().cpu().numpy()), ok)) else ""
    bad += bool(flag)
    print(f"case {case:2d} {rnnt_type:11s} B={B} S={S:3d} T={T:3d} C={C} dp={dp} rel loss err {el:.1e} max grad err {eg:.1e}{flag}")
print("bad", bad)
