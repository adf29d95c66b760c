"""Band path vs dense-lattice path (FRN_BAND_DENSE=1 in a child process) on the c2 test inputs."""
import os, sys, subprocess
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from tf_fast_rnnt import _lib
from tests.helpers import make_inputs
lib = _lib.lib
B, T, S, C, R = 32, 500, 100, 500, 5
am, lm, sym, term, bd = make_inputs(1234, B, T, S, C, ragged=True)
loss, (gx, gy) = frn.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "none", True)
ranges = frn.get_rnnt_prune_ranges(gx, gy, bd, R)
am_d, lm_d, rg_d = (torch.from_numpy(x).cuda() for x in (am, lm, ranges))
am_p, lm_p = frn.do_rnnt_pruning(am_d, lm_d, rg_d)
logits = (am_p + lm_p).contiguous()
n = lib.frn_pruned_loss_workspace_bytes(B, S, T, R)
ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
scores = torch.zeros(B, device="cuda"); grad = torch.zeros_like(logits)
sym_d = torch.from_numpy(sym).cuda(); bd_d = torch.from_numpy(bd).cuda()
rc = lib.frn_pruned_loss(logits.data_ptr(), 0, sym_d.data_ptr(), rg_d.data_ptr(), bd_d.data_ptr(), B, S, T, R, C, term, 0, 0.0,
                         None, scores.data_ptr(), grad.data_ptr(), ws.data_ptr(), n, None)
torch.cuda.synchronize()
assert rc == 0
seg = (B * T * R * 4 + 255) // 256 * 256
raw = ws.cpu().numpy()
arr = lambda k: raw[k * seg:k * seg + B * T * R * 4].view(np.float32).reshape(B, T, R)
pxc, pyc, lse, gxc, gyc = (arr(k).copy() for k in range(5))
tag = "dense" if os.environ.get("FRN_BAND_DENSE") == "1" else "band"
np.savez(f"/tmp/dbg_{tag}.npz", pxc=pxc, pyc=pyc, lse=lse, gxc=gxc, gyc=gyc, scores=scores.cpu().numpy())
if tag == "band":
    subprocess.check_call([sys.executable, __file__], env=dict(os.environ, FRN_BAND_DENSE="1"))
    d = np.load("/tmp/dbg_dense.npz")
    for name in ("pxc", "pyc", "lse"):
        print(name, "max diff", np.nanmax(np.abs(np.nan_to_num(d[name], neginf=-1e30) - np.nan_to_num(locals()[name], neginf=-1e30))))
    ex, ey = np.abs(gxc - d["gxc"]), np.abs(gyc - d["gyc"])
    print("gxc bad", (ex > 1e-4).sum(), "gyc bad", (ey > 1e-4).sum())
    for b, t, i in np.argwhere(ey > 1e-4)[:12]:
        print("gyc b", b, "t", t, "i", i, "band", gyc[b, t, i], "dense", d["gyc"][b, t, i], "r0", ranges[b, t, 0], ranges[b, t + 1, 0], "pyc", pyc[b, t, i])
    for b, t, i in np.argwhere(ex > 1e-4)[:12]:
        print("gxc b", b, "t", t, "i", i, "band", gxc[b, t, i], "dense", d["gxc"][b, t, i], "r0", ranges[b, t, 0], ranges[b, t + 1, 0], "pxc", pxc[b, t, i])
    print("b6 t0 gyc band", gyc[6, 0], "dense", d["gyc"][6, 0], "gxc band", gxc[6, 0], "dense", d["gxc"][6, 0])
    print("b6 t1 gyc band", gyc[6, 1], "dense", d["gyc"][6, 1], "gxc band", gxc[6, 1], "dense", d["gxc"][6, 1])
