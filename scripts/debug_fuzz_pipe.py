import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from oracle import rnnt_oracle as orc
from tests.helpers import make_inputs
want = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.default_rng(4242)
for case in range(30):
    rnnt_type = ["regular", "modified", "constrained"][case % 3]
    B = int(rng.integers(1, 4)); S = int(rng.integers(1, 26)); T = int(rng.integers(max(S, 2), 130))
    C = int(rng.integers(2, 12)) * 4 if case % 2 else int(rng.integers(3, 30))
    R = int(rng.integers(2 if rnnt_type == "regular" else 1, 10))
    am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=bool(case % 5 == 0))
    dp = float([0.0, 0.25, 0.6][case % 3 if case % 2 else 0])
    for k in ("FRN_DP_CHAIN", "FRN_DP_SCAN", "FRN_BAND_DENSE"):
        os.environ.pop(k, None)
    os.environ["FRN_DP_SCAN" if case % 2 else "FRN_DP_CHAIN"] = "1"
    loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, dp, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
    logits = (am_p + lm_p).astype(np.float32)
    w = rng.standard_normal(B).astype(np.float32)
    if case != want:
        continue
    print("case", case, rnnt_type, B, S, T, C, R, dp, bd.tolist())
    o_grad, o_scores = orc.pruned_logits_grad(logits, sym, ranges, term, bd, rnnt_type, dp, w, np.float64, return_scores=True)
    # numerical check of the oracle itself on a few entries (finite differences of the oracle loss)
    def total(lg):
        l = orc.rnnt_loss_pruned(lg, sym, ranges, term, bd, rnnt_type, dp, "none", dtype=np.float64)
        return float((w * l).sum())
    lg64 = logits.astype(np.float64)
    idxs = [tuple(int(v) for v in rng.integers(0, n, 1)) for n in logits.shape]
    worst = np.unravel_index(np.argmax(np.abs(o_grad)), o_grad.shape)
    for idx in [worst, (0, bd[0,1], 0, 0), (0, bd[0,1]+1, 0, 1)]:
        e = 1e-6
        a = lg64.copy(); a[idx] += e; c = lg64.copy(); c[idx] -= e
        print("  FD at", idx, (total(a) - total(c)) / (2 * e), "oracle analytic", o_grad[idx])
    for name, env in (("band", {}), ("dense", {"FRN_BAND_DENSE": "1"})):
        os.environ.pop("FRN_BAND_DENSE", None); os.environ.update(env)
        sc, gr = frn.pruned_loss_fwd_bwd(torch.from_numpy(logits).cuda(), sym, ranges, term, bd, rnnt_type, dp, torch.from_numpy(-w).cuda())
        gr = gr.cpu().numpy()
        print(f"  {name}: scores {sc.cpu().numpy()} oracle {o_scores} per-utt max grad err {np.abs(gr - o_grad).max(axis=(1,2,3))}")
        bad = np.argwhere(np.abs(gr - o_grad) > 1e-3)
        print("    bad idx (b,t,i,c) first:", bad[:6].tolist(), "n", len(bad), " ours/oracle at first:", (gr[tuple(bad[0])], o_grad[tuple(bad[0])]) if len(bad) else None)
    print("ranges[0,:,0]", ranges[0, :, 0].tolist())
