"""Where the wavefront kernel's cycles go (diagnostic build: make -C tf-fast-rnnt_b200/csrc OUT=<dir>
EXTRA_NVFLAGS="-DFRN_DEBUG_HOOKS -DFRN_CHAIN_TIMING", FAST_RNNT_B200_LIB=<dir>/libfast_rnnt_b200.so): per warp of
block 0, cycles waiting for arc chunks, waiting for the feeding warp, inside the steps, total."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
os.environ["FRN_DP_CHAIN"] = "1"
import tf_fast_rnnt as frn
from tf_fast_rnnt import _lib
B, T, S = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 500, 100)))
rng = np.random.default_rng(0)
px = torch.from_numpy((rng.standard_normal((B, S, T + 1)) - 6).astype(np.float32)).cuda()
py = torch.from_numpy((rng.standard_normal((B, S + 1, T)) - 0.5).astype(np.float32)).cuda()
bd = torch.tensor([[0, 0, S, T]] * B, dtype=torch.int32).cuda()
for _ in range(5):
    ans, _ = frn.mutual_information_recursion(px, py, bd, True)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * 64)()
assert lib.frn_debug_chain_timing(buf) == 0
t = np.array(list(buf), dtype=np.int64).reshape(2, 8, 4)
W = (S + 1 + 31) // 32
for d in range(2):
    for w in range(min(W, 8)):
        xy, fin, steps, tot = t[d, w]
        print(f"dir {d} warp {w}: wait_xy {xy:7d}  wait_feeder {fin:7d}  steps {steps:7d}  total {tot:7d}  "
              f"(per diagonal: steps {steps / (S + T + 1):.1f}, total {tot / (S + T + 1):.1f} cycles)")
