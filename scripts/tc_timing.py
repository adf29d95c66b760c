"""Phase boundaries of the normaliser kernel (diagnostic build: make -C tf-fast-rnnt_b200/csrc OUT=<dir>
EXTRA_NVFLAGS="-DFRN_DEBUG_HOOKS -DFRN_TC_TIMING", FAST_RNNT_B200_LIB=<dir>/libfast_rnnt_b200.so): cycles of one CTA
from kernel entry to: operands' constants + dead fill done, contraction done, pass-0 arcs staged, pass-0 written,
pass-1 written, end."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from tf_fast_rnnt import _lib
B, T, S, C = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (32, 500, 100, 500)))
rng = np.random.default_rng(0)
am = torch.from_numpy(rng.standard_normal((B, T, C), dtype=np.float32)).cuda()
lm = torch.from_numpy(rng.standard_normal((B, S + 1, C), dtype=np.float32)).cuda()
sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
bd = np.tile(np.array([0, 0, S, T], np.int32), (B, 1))
for _ in range(4):
    loss, _ = frn.rnnt_loss_simple(lm, am, sym, C - 1, bd, "regular", 0.0, "none", True)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 8)()
assert lib.frn_debug_tc_timing(buf) == 0
t = np.array(list(buf), dtype=np.int64)
names = ["prologue (constants, dead fill, tmem)", "contraction (TMA, MMA, drains)", "pass 0 arcs computed + staged",
         "pass 0 written", "pass 1 computed + written", "end"]
for i, n in enumerate(names):
    print(f"{n:42s} {t[i + 1] - t[i]:8d} cycles   (at {t[i + 1] - t[0]})")
