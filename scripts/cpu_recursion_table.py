"""the lattice recursion (fwd + bwd) of the CPU oracle on the host cores of the box it runs on
(oracle/mi_recursion.c: plain C restatement of mutual_information_cuda.cu:174-874, OpenMP over utterances), plus a
NumPy restatement vectorised over the batch and the anti-diagonal.  Test infrastructure timed as a baseline only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from oracle import rnnt_oracle as orc


def numpy_recursion(px, py):
    """p[b,s,t] = logaddexp(p[b,s-1,t] + px[b,s-1,t], p[b,s,t-1] + py[b,s,t-1]) by anti-diagonals (regular
    recursion, full boundaries), then the backward recursion for the occupation counts (cu:472-481)."""
    B, S, T1 = px.shape
    T = py.shape[2]
    p = np.full((B, S + 1, T + 1), -np.inf, np.float32)
    p[:, 0, 0] = 0
    for d in range(1, S + T + 1):
        s = np.arange(max(0, d - T), min(S, d) + 1)
        t = d - s
        a = np.full((B, len(s)), -np.inf, np.float32)
        m = s > 0
        a[:, m] = p[:, s[m] - 1, t[m]] + px[:, s[m] - 1, t[m]]
        c = np.full((B, len(s)), -np.inf, np.float32)
        m = t > 0
        c[:, m] = p[:, s[m], t[m] - 1] + py[:, s[m], t[m] - 1]
        p[:, s, t] = np.logaddexp(a, c)
    g = np.zeros_like(p)
    g[:, S, T] = 1
    gx, gy = np.zeros_like(px), np.zeros_like(py)
    for d in range(S + T - 1, -1, -1):
        s = np.arange(max(0, d - T), min(S, d) + 1)
        t = d - s
        acc = np.zeros((B, len(s)), np.float32)
        m = s < S
        x = np.exp(p[:, s[m], t[m]] + px[:, s[m], t[m]] - p[:, s[m] + 1, t[m]]) * g[:, s[m] + 1, t[m]]
        gx[:, s[m], t[m]] = x
        acc[:, m] += x
        m = t < T
        y = np.exp(p[:, s[m], t[m]] + py[:, s[m], t[m]] - p[:, s[m], t[m] + 1]) * g[:, s[m], t[m] + 1]
        gy[:, s[m], t[m]] = y
        acc[:, m] += y
        g[:, s, t] = acc
    return p[:, S, T], gx, gy


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"| config | NumPy recursion fwd+bwd (ms) | C/OpenMP (ms, threads) | host cores |")
print("|---|---|---|---|")
for name, B, T, S in [("c1 B2 T50 S10", 2, 50, 10), ("c2 B32 T500 S100", 32, 500, 100)]:
    rng = np.random.default_rng(0)
    px = (rng.standard_normal((B, S, T + 1)) - 6).astype(np.float32)
    py = (rng.standard_normal((B, S + 1, T)) - 0.5).astype(np.float32)
    px[:, :, T] = -np.inf
    bd = np.tile(np.array([0, 0, S, T], np.int32), (B, 1))
    a1, (g1, _) = orc.mutual_information_recursion(px, py, bd, True)
    a2, g2, _ = numpy_recursion(px, py)
    assert np.allclose(a1, a2, rtol=1e-4) and np.allclose(g1, g2, rtol=2e-2, atol=1e-5)
    c_ms = timed(lambda: orc.mutual_information_recursion(px, py, bd, True), 20 if B > 2 else 200)
    n_ms = timed(lambda: numpy_recursion(px, py), 2 if B > 2 else 20)
    thr = os.environ.get("OMP_NUM_THREADS", str(os.cpu_count()))
    print(f"| {name} | {n_ms:.2f} | {c_ms:.3f} ({thr}) | {os.cpu_count()} |")
