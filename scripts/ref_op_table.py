"""the reference's own CUDA op (oracle/_ref/libref_mi.so = mutual_information_cuda.cu compiled
unmodified, driven like tf_fast_rnnt_op.cc:66-113 incl. memsets, H2D copy and stream sync) against frn_mi_fwd_bwd
on the same dense px/py, at the BASELINE config shapes.  Host wall clock per call (the reference op synchronises
the stream itself), many repetitions; prints a markdown table."""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tf-fast-rnnt_b200")]
from tf_fast_rnnt import _lib
lib = _lib.lib
ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_mi.so"))
P, I = ctypes.c_void_p, ctypes.c_int
ref.ref_fast_rnnt_loss.restype = I
ref.ref_fast_rnnt_loss.argtypes = [P, P, P, I, I, I, I, I, P, P, P, P, P, P, I, P]
ref.ref_cummin.restype = I
ref.ref_cummin.argtypes = [P, P, I, I, P]
dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


rows = []
for name, B, T, S, modified in [("c1 B2 T50 S10", 2, 50, 10, False), ("c2 B32 T500 S100", 32, 500, 100, False),
                                ("c3 modified B32 T500 S100", 32, 500, 100, True),
                                ("c4 B16 T1500 S400", 16, 1500, 400, False)]:
    T1 = T if modified else T + 1
    rng = np.random.default_rng(0)
    px = torch.from_numpy((rng.standard_normal((B, S, T1)) - 6).astype(np.float32)).to(dev)
    py = torch.from_numpy((rng.standard_normal((B, S + 1, T)) - 0.5).astype(np.float32)).to(dev)
    bd = torch.tensor([[0, 0, S, T]] * B, dtype=torch.int32, device=dev)
    e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    p_, pg, ans, ag, gx, gy = e(B, S + 1, T + 1), e(B, S + 1, T + 1), e(B), e(B), e(B, S, T1), e(B, S + 1, T)
    ans2, gx2, gy2 = e(B), e(B, S, T1), e(B, S + 1, T)
    ws = torch.empty(max(int(lib.frn_mi_workspace_bytes(B, S, T, T1)), 256), dtype=torch.uint8, device=dev)
    ptr = lambda t: t.data_ptr()

    def ref_op():
        assert ref.ref_fast_rnnt_loss(ptr(px), ptr(py), ptr(bd), B, S, T, T1, 1, ptr(p_), ptr(ans), ptr(pg), ptr(gx),
                                      ptr(gy), ptr(ag), T + 1, st) == 1

    def our_op():
        assert lib.frn_mi_fwd_bwd(ptr(px), ptr(py), ptr(bd), B, S, T, T1, 1, ptr(ans2), ptr(gx2), ptr(gy2), ptr(ws),
                                  ws.numel(), st) == 0

    r, o = timed(ref_op, 30), timed(our_op, 200)
    torch.cuda.synchronize()
    err = (ans - ans2).abs().max().item() / ans.abs().max().item()
    rows.append((f"DP fwd+bwd, dense px/py", name, r, o, f"score rel diff {err:.1e}"))
    if name.startswith("c2") or name.startswith("c4"):
        x = torch.randint(0, S, (B, T), dtype=torch.int32, device=dev)
        y = torch.empty_like(x)
        r = 2 * timed(lambda: ref.ref_cummin(ptr(x), ptr(y), B, T, st), 30)
        o = 2 * timed(lambda: lib.frn_cummin(ptr(x), ptr(y), B, T, st), 200)
        rows.append(("2x cummin [B,T] int32", name, r, o, ""))
print("| sub-step timed | config | reference kernels (us) | new kernels (us) | speed-up | note |")
print("|---|---|---|---|---|---|")
for what, name, r, o, note in rows:
    print(f"| {what} | {name} | {r:.1f} | {o:.1f} | {r / o:.1f}x | {note} |")
