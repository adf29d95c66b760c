import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
RL = sys.modules["tf_fast_rnnt.rnnt_loss"]
from oracle import rnnt_oracle as orc
from tests.helpers import make_inputs
want = [int(a) for a in sys.argv[1:]] or [13, 27, 33]
rng = np.random.default_rng(77)
orig_ws = RL._workspace
for case in range(36):
    rnnt_type = ["regular", "modified", "constrained"][case % 3]
    B = int(rng.integers(1, 4)); S = int(rng.integers(2, 40)); T = int(rng.integers(max(S, 4), 300))
    C = int(rng.integers(3, 20)); R = int(rng.integers(1, min(8, S + 1) + 1))
    am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=False)
    dp = [0.0, 0.3][case % 2]
    RL._workspace = orig_ws
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
    logits = am_p + lm_p
    if case % 4 == 0:
        logits[rng.random(logits.shape) < 0.01] = -np.inf
    w = rng.standard_normal(B).astype(np.float32)
    if case not in want:
        continue
    o_grad, o_scores = orc.pruned_logits_grad(logits, sym, ranges, term, bd, rnnt_type, dp, w, np.float64, return_scores=True)
    o_grad = -np.nan_to_num(o_grad)
    print("case", case, rnnt_type, "B S T C R", B, S, T, C, R, "bd", bd.tolist())
    for fill in (0x00, 0xFF, 0x7F, 0x3F):
        def ws(nbytes, dev, fill=fill):
            return torch.full((max(int(nbytes), 256),), fill, dtype=torch.uint8, device=dev)
        RL._workspace = ws
        sc, gr = frn.pruned_loss_fwd_bwd(torch.from_numpy(logits).cuda(), sym, ranges, term, bd, rnnt_type, dp, torch.from_numpy(w).cuda())
        gr = gr.cpu().numpy()
        print(f"   fill {fill:#04x}: scores {sc.cpu().numpy()} nan in grad {int(np.isnan(gr).sum())} max err {np.abs(np.nan_to_num(gr) - o_grad).max(axis=(1,2,3))}")
