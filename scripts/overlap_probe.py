"""Why "fill the idle SMs"  does not pay: frn_simple_loss (latency-bound: normaliser, wavefront
recursion on 64 of 148 SMs, read-out) alone, the am half of do_rnnt_pruning on G copy-engine CTAs alone
(frn_broadcast_am_pruned), and both at once on two streams - CUDA events, c2 shape.  ncu cannot show this (it
serialises kernels)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
import tf_fast_rnnt as frn
from tf_fast_rnnt import _lib
lib = _lib.lib
B, T, S, C, R = 32, 500, 100, 500, 5
rng = np.random.default_rng(1234)
am = torch.from_numpy(rng.standard_normal((B, T, C), dtype=np.float32)).cuda()
lm = torch.from_numpy(rng.standard_normal((B, S + 1, C), dtype=np.float32)).cuda()
sym = torch.from_numpy(rng.integers(0, C - 1, (B, S)).astype(np.int32)).cuda()
bd = torch.tensor([[0, 0, S, T]] * B, dtype=torch.int32).cuda()
am_p = torch.empty((B, T, R, C), dtype=torch.float32, device="cuda")
scores = torch.empty(B, device="cuda"); gx = torch.empty((B, S, T + 1), device="cuda"); gy = torch.empty((B, S + 1, T), device="cuda")
ws = torch.empty(int(lib.frn_simple_loss_workspace_bytes(B, S, T, C)), dtype=torch.uint8, device="cuda")
main, side = torch.cuda.current_stream(), torch.cuda.Stream()


def loss():
    _lib.check(lib.frn_simple_loss(lm.data_ptr(), am.data_ptr(), sym.data_ptr(), bd.data_ptr(), B, S, T, C, C - 1, 0, 0,
                                   0.0, 0.0, 0.0, 1, scores.data_ptr(), gx.data_ptr(), gy.data_ptr(), ws.data_ptr(),
                                   ws.numel(), main.cuda_stream), "simple_loss")


def copy(G):
    _lib.check(lib.frn_broadcast_am_pruned(am.data_ptr(), B, T, R, C, am_p.data_ptr(), G, side.cuda_stream), "bcast")


def timed(fn, n=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(n):
        fn()
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def both(G):
    side.wait_stream(main)
    copy(G)
    loss()
    main.wait_stream(side)


def copy_alone(G):
    side.wait_stream(main)
    copy(G)
    main.wait_stream(side)


print(f"frn_simple_loss alone: {timed(loss):.1f} us")
for G in (16, 40, 84):
    print(f"G={G:3d}: copy alone {timed(lambda: copy_alone(G)):.1f} us, both at once {timed(lambda: both(G)):.1f} us")
