import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from tests.helpers import make_inputs
rng = np.random.default_rng(77)
def setenv(env):
    for k in ("FRN_BAND_DENSE", "FRN_DP_CHAIN", "FRN_DP_SCAN"):
        os.environ.pop(k, None)
    os.environ.update(env)
ENVS = (("band", {}), ("chain", {"FRN_BAND_DENSE": "1", "FRN_DP_CHAIN": "1"}), ("scan", {"FRN_BAND_DENSE": "1", "FRN_DP_SCAN": "1"}))
for case in range(int(os.environ.get('FUZZ_N', '36'))):
    rnnt_type = ["regular", "modified", "constrained"][case % 3]
    if os.environ.get("FUZZ_BIG"):
        B = int(rng.integers(1, 4)); S = int(rng.integers(40, 400)); T = int(rng.integers(max(S, 300), 1500))
    else:
        B = int(rng.integers(1, 4)); S = int(rng.integers(2, 40)); T = int(rng.integers(max(S, 4), 300))
    C = int(rng.integers(3, 20)); R = int(rng.integers(1, min(8, S + 1) + 1))
    am, lm, sym, term, bd = make_inputs(int(rng.integers(1 << 30)), B, T, S, C, ragged=True, begin=False)
    dp = [0.0, 0.3][case % 2]
    setenv({})
    _, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, rnnt_type, 0.0, "none", True)
    ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
    am_p, lm_p = frn.do_rnnt_pruning(am, lm, ranges)
    logits = am_p + lm_p
    if case % 4 == 0:
        logits[rng.random(logits.shape) < 0.01] = -np.inf
    w = rng.standard_normal(B).astype(np.float32)
    out = {}
    for rep in range(2):
        for name, env in ENVS:
            setenv(env)
            sc, gr = frn.pruned_loss_fwd_bwd(torch.from_numpy(logits).cuda(), sym, ranges, term, bd, rnnt_type, dp, torch.from_numpy(w).cuda())
            out[(name, rep)] = (sc.cpu().numpy(), np.nan_to_num(gr.cpu().numpy()))
    d = lambda a, b: float(np.abs(out[a][1] - out[b][1]).max())
    line = f"case {case:2d} {rnnt_type:11s} B={B} S={S:2d} T={T:3d} R={R} | band-chain {d(('band',0),('chain',0)):.1e} band-scan {d(('band',0),('scan',0)):.1e} chain-scan {d(('chain',0),('scan',0)):.1e} | rep: band {d(('band',0),('band',1)):.1e} chain {d(('chain',0),('chain',1)):.1e} scan {d(('scan',0),('scan',1)):.1e}"
    bad = max(d(('band',0),('chain',0)), d(('band',0),('scan',0)), d(('chain',0),('chain',1))) > 1e-4
    print(line + ("  <<<<" if bad else ""))
    if bad:
        print("      scores band", out[("band", 0)][0], "chain", out[("chain", 0)][0], "dp", dp)
