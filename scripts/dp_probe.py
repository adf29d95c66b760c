"""Diagnostics on the GPU box: accuracy of the lattice recursion (ours, the
reference's kernels, the float32 oracle — all against the float64 oracle) and a
first timing of ours vs the reference's kernels.  Not part of the product."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from oracle import rnnt_oracle as orc
from tests.helpers import RefKernels, random_pxpy, make_inputs

ref = RefKernels()

def stats(name, g, g64):
    d = np.abs(g.astype(np.float64) - g64)
    m = g64 > 1e-3
    rel = d[m] / g64[m]
    i = np.unravel_index(np.argmax(d), d.shape)
    print(f"   {name:10s} max abs {d.max():.2e} at {i} (truth {g64[i]:.4f})  max rel(>1e-3) {rel.max():.2e}  median rel {np.median(rel):.2e}")

for (B, S, T, mod, scale) in [(2, 127, 129, False, 1.0), (4, 33, 70, True, 1.0), (2, 64, 64, False, 1.0),
                              (2, 130, 300, False, 1.0), (2, 100, 500, False, 4.0)]:
    px, py = random_pxpy(1, B, S, T, mod, scale)
    bd = np.tile(np.array([0, 0, S, T], np.int32), (B, 1))
    a64, (gx64, gy64) = orc.mutual_information_recursion(px, py, bd, True, np.float64)
    a32, (gx32, gy32) = orc.mutual_information_recursion(px, py, bd, True, np.float32)
    a, (gx, gy) = frn.mutual_information_recursion(px, py, bd, True)
    r_a, r_gx, r_gy, _ = ref.fast_rnnt_loss(px, py, bd)
    print(f"B{B} S{S} T{T} modified={mod} scale={scale}: score {a64[0]:.3f}; rel err ours {abs(a[0]-a64[0])/abs(a64[0]):.1e} ref {abs(r_a[0]-a64[0])/abs(a64[0]):.1e} f32 {abs(a32[0]-a64[0])/abs(a64[0]):.1e}")
    stats("ours gx", gx, gx64); stats("ref  gx", r_gx, gx64); stats("f32  gx", gx32, gx64)
    stats("ours gy", gy, gy64); stats("ref  gy", r_gy, gy64)

# timing at c2 (DP only, device-resident)
B, T, S, C = 32, 500, 100, 500
am, lm, sym, term, bd = make_inputs(1234, B, T, S, C, ragged=False)
px, py = orc.get_rnnt_logprobs(lm, am, sym, term, "regular", bd)
pxd, pyd, bdd = torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda(), torch.from_numpy(bd).cuda()
for name, fn in [("ours", lambda: frn.mutual_information_recursion(pxd, pyd, bdd, True)),
                 ("reference kernels", lambda: ref.fast_rnnt_loss(pxd, pyd, bdd))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"c2 DP fwd+bwd {name}: {e0.elapsed_time(e1) / 10 * 1000:.1f} us per call (incl. host-side wrapper)")
