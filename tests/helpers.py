"""Shared helpers for the parity tests (inputs, oracle/_ref binding, comparisons)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# Tolerances of BASELINE.json north_star: losses 1e-5 relative, px/py gradients
# 1e-4 relative (+1e-6 absolute: occupation counts span e^-80..1), both against
# the float64 oracle.
LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4
GRAD_ATOL = 1e-6


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def make_inputs(seed, B, T, S, C, ragged=True, begin=False):
    """Synthetic inputs of SURVEY.md §8(d): N(0,1) am/lm, uniform symbols,
    termination_symbol = C-1, boundary [s_begin, t_begin, S_b, T_b]."""
    rng = np.random.default_rng(seed)
    am = rng.standard_normal((B, T, C), dtype=np.float32)
    lm = rng.standard_normal((B, S + 1, C), dtype=np.float32)
    symbols = rng.integers(0, C - 1, (B, S)).astype(np.int32)
    boundary = np.zeros((B, 4), dtype=np.int32)
    boundary[:, 2] = S
    boundary[:, 3] = T
    if ragged:
        tb = rng.integers(max(1, int(0.6 * T)), T + 1, (B,))
        sb = rng.integers(max(1, int(0.2 * S)), S + 1, (B,))
        sb = np.minimum(sb, tb)
        boundary[:, 2] = sb
        boundary[:, 3] = tb
        boundary[0, 2], boundary[0, 3] = S, T          # keep one full-size utterance
    if begin:
        boundary[:, 0] = rng.integers(0, 2, (B,))
        boundary[:, 1] = rng.integers(0, 3, (B,))
        boundary[:, 0] = np.minimum(boundary[:, 0], boundary[:, 2])
        boundary[:, 1] = np.minimum(boundary[:, 1], boundary[:, 3])
    return am, lm, symbols, C - 1, boundary


def random_pxpy(seed, B, S, T, modified, scale=1.0):
    rng = np.random.default_rng(seed)
    T1 = T if modified else T + 1
    px = (rng.standard_normal((B, S, T1)) * scale - 1.0).astype(np.float32)
    py = (rng.standard_normal((B, S + 1, T)) * scale - 1.0).astype(np.float32)
    return px, py


def assert_close(actual, desired, rtol, atol=0.0, what=""):
    actual = np.asarray(actual, dtype=np.float64)
    desired = np.asarray(desired, dtype=np.float64)
    assert actual.shape == desired.shape, (what, actual.shape, desired.shape)
    fin = np.isfinite(desired)
    assert np.array_equal(np.isfinite(actual), fin), f"{what}: finite pattern differs"
    assert np.array_equal(actual[~fin], desired[~fin]), f"{what}: inf pattern differs"
    err = np.abs(actual[fin] - desired[fin])
    tol = atol + rtol * np.abs(desired[fin])
    bad = err > tol
    assert not bad.any(), (
        f"{what}: {bad.sum()} / {bad.size} outside rtol={rtol} atol={atol}; "
        f"max abs err {err.max():.3e}, max err/tol {(err / np.maximum(tol, 1e-300)).max():.3f}")


class RefKernels:
    """ctypes binding of oracle/_ref/libref_mi.so — the reference's own CUDA
    kernels (built by oracle/Makefile from /root/reference, travels prebuilt)."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "_ref", "libref_mi.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        P = ctypes.c_void_p
        I = ctypes.c_int
        self.lib.ref_fast_rnnt_loss.restype = I
        self.lib.ref_fast_rnnt_loss.argtypes = [P, P, P, I, I, I, I, I, P, P, P, P, P, P, I, P]
        self.lib.ref_cummin.restype = I
        self.lib.ref_cummin.argtypes = [P, P, I, I, P]

    def fast_rnnt_loss(self, px, py, boundary, calc_gradients=True):
        import torch
        dev = torch.device("cuda")
        px_d = torch.as_tensor(px, dtype=torch.float32, device=dev).contiguous()
        py_d = torch.as_tensor(py, dtype=torch.float32, device=dev).contiguous()
        bd = torch.as_tensor(boundary, dtype=torch.int32, device=dev).contiguous()
        B, S, T1 = px_d.shape
        T = py_d.shape[2]
        p = torch.full((B, S + 1, T + 1), float("nan"), device=dev)
        ans = torch.zeros(B, device=dev)
        pg = torch.zeros((B, S + 1, T + 1), device=dev)
        gx = torch.zeros((B, S, T1), device=dev)
        gy = torch.zeros((B, S + 1, T), device=dev)
        ag = torch.zeros(B, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        rc = self.lib.ref_fast_rnnt_loss(px_d.data_ptr(), py_d.data_ptr(), bd.data_ptr(), B, S, T, T1,
                                         int(calc_gradients), p.data_ptr(), ans.data_ptr(), pg.data_ptr(),
                                         gx.data_ptr(), gy.data_ptr(), ag.data_ptr(), T1, stream)
        assert rc == 1
        return ans.cpu().numpy(), gx.cpu().numpy(), gy.cpu().numpy(), ag.cpu().numpy()

    def cummin(self, x):
        import torch
        x_d = torch.as_tensor(x, dtype=torch.int32, device="cuda").contiguous()
        out = torch.empty_like(x_d)
        rc = self.lib.ref_cummin(x_d.data_ptr(), out.data_ptr(), x_d.shape[0], x_d.shape[1],
                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 1
        return out.cpu().numpy()
