"""Dump the DP workspace planes of a tiny lattice (debug aid)."""
import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
from tf_fast_rnnt import _lib
from oracle import rnnt_oracle as orc
lib = _lib.lib
B, S, T = 1, 2, 3
rng = np.random.default_rng(0)
px = (rng.standard_normal((B, S, T + 1)) - 2).astype(np.float32)
py = (rng.standard_normal((B, S + 1, T)) - 2).astype(np.float32)
bd = np.array([[0, 0, S, T]], np.int32)
d = lambda a: torch.from_numpy(a).cuda()
pxd, pyd, bdd = d(px), d(py), d(bd)
n = lib.frn_mi_workspace_bytes(B, S, T, T + 1)
ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
ans = torch.zeros(B, device="cuda"); gx = torch.zeros(B, S, T + 1, device="cuda"); gy = torch.zeros(B, S + 1, T, device="cuda")
rc = lib.frn_mi_fwd_bwd(pxd.data_ptr(), pyd.data_ptr(), bdd.data_ptr(), B, S, T, T + 1, 1, ans.data_ptr(), gx.data_ptr(),
                        gy.data_ptr(), ws.data_ptr(), n, None)
torch.cuda.synchronize()
print("rc", rc, "ans", ans.cpu().numpy(), "oracle", orc.mutual_information_recursion(px, py, bd, True, np.float64)[0])
P, Dn = 32, 16
cells = B * Dn * P
raw = ws.cpu().numpy()
XY = raw[:cells * 16].view(np.float32).reshape(Dn, P, 4)
XYi = raw[:cells * 16].view(np.int32).reshape(Dn, P, 4)
offA = (cells * 16 + 255) // 256 * 256
A = raw[offA:offA + cells * 8].view(np.float32).reshape(Dn, P, 2)
Ai = raw[offA:offA + cells * 8].view(np.int32).reshape(Dn, P, 2)
for dd in range(T + S + 1):
    print("d", dd, "X", [(float(XY[dd, s, 0]), int(XYi[dd, s, 1])) for s in range(S + 1)],
          "Y", [(float(XY[dd, s, 2]), int(XYi[dd, s, 3])) for s in range(S + 1)],
          "A", [(float(A[dd, s, 0]), int(Ai[dd, s, 1])) for s in range(S + 1)])
print("px*log2e", px * 1.4427, "py*log2e", py * 1.4427)
