import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from oracle import rnnt_oracle as orc
from tests.helpers import make_inputs
want = int(sys.argv[1]) if len(sys.argv) > 1 else 13
rng = np.random.default_rng(77)
def clear():
    for k in ("FRN_BAND_DENSE", "FRN_DP_CHAIN", "FRN_DP_SCAN"):
        os.environ.pop(k, None)
for case in range(36):
    rnnt_type = ["regular", "modified", "constrained"][case % 3]
    B = int(rng.integers(1, 4)); S = int(rng.integers(2, 40)); T = int(rng.integers(max(S, 4), 300))
    C = int(rng.integers(3, 20)); R = int(rng.integers(1, min(8, S + 1) + 1))
    am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=False)
    dp = [0.0, 0.3][case % 2]
    clear()
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
    logits = am_p + lm_p
    if case % 4 == 0:
        logits[rng.random(logits.shape) < 0.01] = -np.inf
    w = rng.standard_normal(B).astype(np.float32)
    if case != want:
        continue
    print("case", case, rnnt_type, "B S T C R dp", B, S, T, C, R, dp, "bd", bd.tolist())
    o_grad, o_scores = orc.pruned_logits_grad(logits, sym, ranges, term, bd, rnnt_type, dp, w, np.float64, return_scores=True)
    for name, env in (("band", {}), ("dense chain", {"FRN_BAND_DENSE": "1", "FRN_DP_CHAIN": "1"}), ("dense scan", {"FRN_BAND_DENSE": "1", "FRN_DP_SCAN": "1"})):
        clear(); os.environ.update(env)
        sc, gr = frn.pruned_loss_fwd_bwd(torch.from_numpy(logits).cuda(), sym, ranges, term, bd, rnnt_type, dp, torch.from_numpy(w).cuda())
        sc, gr = sc.cpu().numpy(), gr.cpu().numpy()
        print(f"{name:12s} scores {sc} oracle {o_scores}  per-utt max grad err", np.abs(gr - o_grad).max(axis=(1, 2, 3)))
