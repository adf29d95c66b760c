#!/usr/bin/env python
"""Benchmark of the pruned RNN-T loss hot path (BASELINE.json metric).

One "step" = one pass of the full pruned pipeline over one batch of synthetic
utterances, through the C ABI (include/fast_rnnt_b200.h):

    frn_simple_loss (fwd + occupation counts)  ->  frn_prune_ranges
    -> frn_do_pruning -> frn_add_joiner (additive joiner of the reference's
    tests, standing in for the user's joiner network) -> frn_pruned_loss
    (fwd + logits gradient) -> frn_reduce x2 (+ one 2-float NCCL all-reduce
    when N > 1)

Workload: BASELINE.json configs[1] — B=32 T=500 S=100 C=500 s_range=5 fp32,
rnnt_type=regular, reduction=sum.  `value` is utterances/s with inputs resident
in HBM; `e2e` is the same step driven from pinned HOST buffers (H2D of
am/lm/symbols/boundary and D2H of the two losses inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the CPU restatement of the reference's path (the
oracle; the reference has no CPU implementation and its TensorFlow half cannot
run here) on the host cores.  Under torchrun only rank 0 runs it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "tf-fast-rnnt_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: (B, T, S, C, R)
    "c1": (2, 50, 10, 16, 5),
    "c2": (32, 500, 100, 500, 5),
    "c4": (16, 1500, 400, 5000, 5),
    "c5": (256, 1500, 400, 500, 5),     # ragged: run_c5(); the CPU arm times 2 full-length utterances of it
}
METRIC = "utterances/sec (pruned loss fwd+bwd) at B32 T500 S100 C500; lattice cells/s; HBM %"
UNIT = "utterances/s"


# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, for one) write to
# file descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the
# JSON line goes to a private duplicate of the original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth(B, T, S, C, seed):
    """N(0,1) am/lm, uniform symbols, blank = C-1, full-length boundary."""
    rng = np.random.default_rng(seed)
    am = rng.standard_normal((B, T, C), dtype=np.float32)
    lm = rng.standard_normal((B, S + 1, C), dtype=np.float32)
    sym = rng.integers(0, C - 1, (B, S)).astype(np.int32)
    bd = np.tile(np.array([0, 0, S, T], np.int32), (B, 1))
    return am, lm, sym, bd


# ----------------------------------------------------------------------------
# CPU arm (oracle port) — also the cpu_baseline leg of the GPU arm
# ----------------------------------------------------------------------------
def cpu_step(am, lm, sym, bd, C, R):
    from oracle import rnnt_oracle as orc
    term = C - 1
    loss, (gx, gy) = orc.rnnt_loss_simple(lm, am, sym, term, bd, "regular", 0.0, "sum", True)
    ranges = orc.get_rnnt_prune_ranges(gx, gy, bd, R)
    am_p, lm_p = orc.do_rnnt_pruning(am, lm, ranges)
    logits = am_p + lm_p
    grad, scores = orc.pruned_logits_grad(logits, sym, ranges, term, bd, "regular", 0.0, None, np.float32,
                                          return_scores=True)      # pruned loss fwd + bwd in one pass
    return float(loss), float(-scores.sum()), grad


def cpu_bench(workload, steps, warmup, sample_B=None):
    B, T, S, C, R = WORKLOADS[workload]
    Bs = sample_B or B
    am, lm, sym, bd = synth(Bs, T, S, C, 1234)
    for _ in range(warmup):
        cpu_step(am, lm, sym, bd, C, R)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(am, lm, sym, bd, C, R)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    cores = os.cpu_count() or 1
    return {"value": Bs / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} full steps of {workload} with B={Bs} (numpy/BLAS + OpenMP C recursion, "
                      f"all {cores} host threads), {dt * 1e3:.0f} ms per step"}, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    B = WORKLOADS[workload][0]
    steps = max(1, min(args.steps, 5))
    base, dt = cpu_bench(workload, steps, min(args.warmup, 1), 2 if workload == "c5" else None)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{workload}: full pruned pipeline B={B} T/S/C/R={WORKLOADS[workload][1:]} fp32 "
                               "regular sum; CPU restatement of the reference path (no CPU kernel exists in "
                               "the reference; TensorFlow absent)"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
class Pipeline:
    """Pre-allocated device buffers + the C-ABI call sequence of one step.

    `nsplit` > 1 runs the step as that many independent sub-batches, each on its own CUDA stream
    (utterances are independent in every kernel): the dependency-chain-bound kernels of one
    sub-batch (lattice recursions, normaliser) overlap the bandwidth-bound kernels of another.
    Same calls, same results; the two loss sums are reduced over the whole batch at the end."""

    def __init__(self, B, T, S, C, R, dev, nsplit=1, overlap=False, fuse_add=True, am_side=0, am_side_at="start", logits_bf16=False,
                 training=False):
        import torch
        from tf_fast_rnnt import _lib
        self.torch, self.lib, self._lib = torch, _lib.lib, _lib
        self.B, self.T, self.S, self.C, self.R, self.dev = B, T, S, C, R, dev
        f32, i32 = torch.float32, torch.int32
        e = lambda *shape, dtype=f32: torch.empty(shape, dtype=dtype, device=dev)
        self.scores, self.pscores = e(B), e(B)
        self.gx, self.gy = e(B, S, T + 1), e(B, S + 1, T)
        self.ranges = e(B, T, R, dtype=i32)
        self.am_p, self.lm_p = e(B, T, R, C), e(B, T, R, C)
        # bf16 joiner logits into the float32 loss (BASELINE.json configs[3]): the joiner writes bf16
        # (frn_pruned_add_joiner), frn_pruned_loss reads them and returns a bf16 logits gradient
        self.logits_bf16 = logits_bf16
        ldt = torch.bfloat16 if logits_bf16 else f32
        self.logits, self.dlogits = e(B, T, R, C, dtype=ldt), e(B, T, R, C, dtype=ldt)
        # training step (A9 + the gradient of do_rnnt_pruning): am / lm gradients of both losses
        self.training = training
        if training:
            self.am_g_simple, self.lm_g_simple = e(B, T, C), e(B, S + 1, C)
            self.am_g_pruned, self.lm_g_pruned = e(B, T, C), e(B, S + 1, C)
            self.ws_bwd = torch.empty(max(int(self.lib.frn_simple_loss_bwd_workspace_bytes(B, S, T, C)), 256),
                                      dtype=torch.uint8, device=dev)
            assert not logits_bf16, "the training-step variant runs float32 logits"
        self.losses = e(2)
        self.sgrad = torch.full((B,), -1.0, dtype=f32, device=dev)  # d(sum loss)/d scores
        # am_pruned[b,t,i,:] = am[b,t,:] does not depend on the ranges (rnnt_loss.py:802-806 broadcasts am
        # before it gathers lm): with `overlap` that half of do_rnnt_pruning runs on a second stream
        # beside the dependency-chain-bound kernels of the simple loss, the lm half after the ranges.
        self.overlap = overlap and not fuse_add
        self.fuse_add = fuse_add and not logits_bf16
        # `am_side` = G > 0: the same idea on the copy engine - frn_broadcast_am_pruned, a persistent grid of G
        # single-warp CTAs doing nothing but bulk copies, with a shared-memory footprint that keeps the
        # normaliser's / recursion's big CTAs on the other SMs; frn_do_pruning_add_joiner then writes
        # lm_pruned and the logits only.
        self.am_side = am_side if self.fuse_add else 0     # (bf16 logits take the two-pass form: no am-side copy)
        self.am_side_at = am_side_at          # "start": beside the whole simple loss; "chain": forked behind the normaliser
        self.side = torch.cuda.Stream(dev) if (overlap or self.am_side) else None
        self.fork_ev = self.join_ev = None
        if self.am_side and am_side_at == "chain":
            # two caller-owned events for frn_simple_loss_bcast (created by their first record)
            self.fork_ev, self.join_ev = torch.cuda.Event(), torch.cuda.Event()
            self.fork_ev.record(); self.join_ev.record()
            torch.cuda.synchronize(dev)
        self.full = self._part(0, B)
        self.ws_pruned = self.full["ws_pruned"]
        nsplit = max(1, min(nsplit, B))
        self.parts, self.streams = [], []
        if nsplit > 1:
            edges = [round(i * B / nsplit) for i in range(nsplit + 1)]
            self.parts = [self._part(edges[i], edges[i + 1] - edges[i]) for i in range(nsplit)]
            self.streams = [torch.cuda.Stream(dev) for _ in range(nsplit)]

    def _part(self, b0, Bh):
        lib, T, S, C, R, dev = self.lib, self.T, self.S, self.C, self.R, self.dev
        u8 = lambda n: self.torch.empty(max(int(n), 256), dtype=self.torch.uint8, device=dev)
        return {"b0": b0, "B": Bh,
                "ws_simple": u8(lib.frn_simple_loss_workspace_bytes(Bh, S, T, C)),
                "ws_prune": u8(lib.frn_prune_ranges_workspace_bytes(Bh, T)),
                "ws_pruned": u8(lib.frn_pruned_loss_workspace_bytes(Bh, S, T, R))}

    def stages(self, am, lm, sym, bd, part=None, with_reduce=True):
        """List of (name, algorithmic bytes, callable) — SURVEY.md §8(d) byte counts."""
        part = part or self.full
        lib, T, S, C, R = self.lib, self.T, self.S, self.C, self.R
        b0, B = part["b0"], part["B"]
        p = lambda t: t.data_ptr() + b0 * t.stride(0) * t.element_size()     # sub-batch view of a [B,...] tensor
        term = C - 1
        st = lambda: self.torch.cuda.current_stream(self.dev).cuda_stream
        chk = self._lib.check
        n_logits = B * T * R * C
        ws_s, ws_p, ws_q = part["ws_simple"], part["ws_prune"], part["ws_pruned"]
        out = [
            ("simple_loss", 4 * B * ((T + S + 1) * C + 2 * (S * (T + 1) + (S + 1) * T)),
             lambda: chk(lib.frn_simple_loss(p(lm), p(am), p(sym), p(bd), B, S, T, C, term, 0, 0, 0.0, 0.0, 0.0, 1,
                                             p(self.scores), p(self.gx), p(self.gy), ws_s.data_ptr(), ws_s.numel(),
                                             st()), "simple_loss")),
            ("prune_ranges", 4 * B * (S * (T + 1) + (S + 1) * T) + 4 * B * T * R,
             lambda: chk(lib.frn_prune_ranges(p(self.gx), p(self.gy), p(bd), B, S, T, T + 1, R, p(self.ranges),
                                              ws_p.data_ptr(), ws_p.numel(), st()), "prune_ranges")),
            ("do_pruning", 4 * B * (T * C + (S + 1) * C + T * R) + 8 * n_logits,
             lambda: chk(lib.frn_do_pruning(p(am), p(lm), p(self.ranges), B, S, T, R, C, p(self.am_p),
                                            p(self.lm_p), st()), "do_pruning")),
            ("add_joiner", 12 * n_logits,
             lambda: chk(lib.frn_add_joiner(p(self.am_p), p(self.lm_p), p(self.logits), n_logits, st()),
                         "add_joiner")),
            ("pruned_loss", 4 * n_logits + 8 * B * T * R + 8 * n_logits + 16 * B * T * R,
             lambda: chk(lib.frn_pruned_loss(p(self.logits), 0, p(sym), p(self.ranges), p(bd), B, S, T, R, C, term,
                                             0, 0.0, p(self.sgrad), p(self.pscores), p(self.dlogits),
                                             ws_q.data_ptr(), ws_q.numel(), st()), "pruned_loss")),
        ]
        if self.logits_bf16:
            ldt = 1
            out[3] = ("add_joiner(bf16)", 4 * B * (T * C + (S + 1) * C + T * R) + 2 * n_logits,
                      lambda: chk(lib.frn_pruned_add_joiner(p(am), p(lm), p(self.ranges), B, S, T, R, C, ldt,
                                                            p(self.logits), st()), "pruned_add_joiner"))
            out[4] = ("pruned_loss", 2 * n_logits + 8 * B * T * R + 4 * n_logits + 16 * B * T * R,
                      lambda: chk(lib.frn_pruned_loss(p(self.logits), ldt, p(sym), p(self.ranges), p(bd), B, S, T, R, C,
                                                      term, 0, 0.0, p(self.sgrad), p(self.pscores), p(self.dlogits),
                                                      ws_q.data_ptr(), ws_q.numel(), st()), "pruned_loss"))
        if self.training:
            # backward of the step: A9 (am / lm gradients of the simple loss from its occupation counts) and the
            # gradient of do_rnnt_pruning + additive joiner (d am_pruned = d lm_pruned = d logits)
            ws_b = self.ws_bwd
            out.append(("simple_loss_bwd", 4 * B * (2 * (T + S + 1) * C + 2 * (S * (T + 1) + (S + 1) * T)),
                        lambda: chk(lib.frn_simple_loss_bwd(p(lm), p(am), p(sym), p(bd), p(self.gx), p(self.gy),
                                                            p(self.sgrad), B, S, T, C, term, 0, p(self.am_g_simple),
                                                            p(self.lm_g_simple), ws_b.data_ptr(), ws_b.numel(), st()),
                                    "simple_loss_bwd")))
            out.append(("do_pruning_bwd", 4 * n_logits * 2 + 4 * B * (T + S + 1) * C,
                        lambda: chk(lib.frn_do_pruning_bwd(p(self.dlogits), p(self.dlogits), p(self.ranges), B, S, T, R, C,
                                                           p(self.am_g_pruned), p(self.lm_g_pruned), st()),
                                    "do_pruning_bwd")))
        if with_reduce:
            out.append(("reduce", 8 * self.B, self._reduce))
        if self.fuse_add:
            # do_rnnt_pruning and the additive joiner in ONE pass (frn_do_pruning_add_joiner): am_pruned,
            # lm_pruned and logits are all written, the two pruned tensors are not read back for the sum
            if self.am_side and part is self.full:
                # am_pruned comes from frn_broadcast_am_pruned (beside the recursion): lm_pruned and the logits here
                out[2:4] = [("do_pruning+add_joiner(lm,logits)", 4 * B * (T * C + (S + 1) * C + T * R) + 8 * n_logits,
                             lambda: chk(lib.frn_do_pruning_add_joiner(p(am), p(lm), p(self.ranges), B, S, T, R, C, 0,
                                                                       p(self.lm_p), p(self.logits), st()),
                                         "do_pruning_add_joiner(lm, logits)"))]
            else:
                out[2:4] = [("do_pruning+add_joiner", 4 * B * (T * C + (S + 1) * C + T * R) + 12 * n_logits,
                             lambda: chk(lib.frn_do_pruning_add_joiner(p(am), p(lm), p(self.ranges), B, S, T, R, C,
                                                                       p(self.am_p), p(self.lm_p), p(self.logits), st()),
                                         "do_pruning_add_joiner"))]
        return out

    def _reduce(self):
        lib, chk, B = self.lib, self._lib.check, self.B
        st = self.torch.cuda.current_stream(self.dev).cuda_stream
        chk(lib.frn_reduce_pair(self.scores.data_ptr(), self.pscores.data_ptr(), B, 2, 0.0, self.losses.data_ptr(),
                                self.losses.data_ptr() + 4, st), "reduce_pair")

    def step(self, am, lm, sym, bd):
        torch = self.torch
        if not self.parts and self.am_side:
            lib, chk, B, S, T, R, C = self.lib, self._lib.check, self.B, self.S, self.T, self.R, self.C
            st = self.stages(am, lm, sym, bd)
            main = torch.cuda.current_stream(self.dev)
            if self.am_side_at == "chain":
                # one call: the broadcast is forked behind the normaliser and runs beside the recursion only
                ws_s = self.full["ws_simple"]
                chk(lib.frn_simple_loss_bcast(lm.data_ptr(), am.data_ptr(), sym.data_ptr(), bd.data_ptr(), B, S, T, C,
                                              C - 1, 0, 0, 0.0, 0.0, 0.0, 1, self.scores.data_ptr(), self.gx.data_ptr(),
                                              self.gy.data_ptr(), R, self.am_p.data_ptr(), self.am_side,
                                              self.side.cuda_stream, self.fork_ev.cuda_event, self.join_ev.cuda_event,
                                              ws_s.data_ptr(), ws_s.numel(), main.cuda_stream), "simple_loss_bcast")
                st[1][2]()                                           # prune ranges
            else:
                self.side.wait_stream(main)
                with torch.cuda.stream(self.side):
                    chk(lib.frn_broadcast_am_pruned(am.data_ptr(), B, T, R, C, self.am_p.data_ptr(), self.am_side,
                                                    self.side.cuda_stream), "broadcast_am_pruned")
                st[0][2](); st[1][2]()                               # simple loss, prune ranges
                main.wait_stream(self.side)
            for _, _, fn in st[2:]:                                  # lm_pruned + logits, pruned loss, reductions
                fn()
            return
        if not self.parts and self.overlap:
            lib, chk, B, S, T, R, C = self.lib, self._lib.check, self.B, self.S, self.T, self.R, self.C
            st = self.stages(am, lm, sym, bd)
            main = torch.cuda.current_stream(self.dev)
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                chk(lib.frn_do_pruning(am.data_ptr(), 0, 0, B, S, T, R, C, self.am_p.data_ptr(), 0,
                                       self.side.cuda_stream), "do_pruning(am)")
            st[0][2](); st[1][2]()                                   # simple loss, prune ranges
            chk(lib.frn_do_pruning(0, lm.data_ptr(), self.ranges.data_ptr(), B, S, T, R, C, 0, self.lm_p.data_ptr(),
                                   main.cuda_stream), "do_pruning(lm)")
            main.wait_stream(self.side)
            for _, _, fn in st[3:]:                                  # joiner, pruned loss, reductions (fuse_add is off here)
                fn()
            return
        if not self.parts:
            for _, _, fn in self.stages(am, lm, sym, bd):
                fn()
            return
        main = torch.cuda.current_stream(self.dev)
        for part, s in zip(self.parts, self.streams):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                for _, _, fn in self.stages(am, lm, sym, bd, part, with_reduce=False):
                    fn()
        for s in self.streams:
            main.wait_stream(s)
        self._reduce()


def reference_gpu_op_leg(pipe, am, lm, sym, bd, reps=50):
    """The reference's OWN CUDA op (oracle/_ref/libref_mi.so = mutual_information_cuda.cu compiled
    unmodified for sm_100, driven like tf_fast_rnnt_op.cc:66-113 incl. its memsets, H2D copy and
    stream sync) timed on this GPU beside frn_mi_fwd_bwd on the SAME px/py — a reported baseline
    (BASELINE.json north_star), not part of the product path.  Returns None if the .so is absent."""
    import ctypes
    import torch
    path = os.path.join(ROOT, "oracle", "_ref", "libref_mi.so")
    if not os.path.exists(path):
        return None
    ref = ctypes.CDLL(path)
    P, I = ctypes.c_void_p, ctypes.c_int
    ref.ref_fast_rnnt_loss.restype = I
    ref.ref_fast_rnnt_loss.argtypes = [P, P, P, I, I, I, I, I, P, P, P, P, P, P, I, P]
    ref.ref_cummin.restype = I
    ref.ref_cummin.argtypes = [P, P, I, I, P]
    lib, chk = pipe.lib, pipe._lib.check
    B, T, S, C, R, dev = pipe.B, pipe.T, pipe.S, pipe.C, pipe.R, pipe.dev
    f32 = torch.float32
    e = lambda *shape, dtype=f32: torch.empty(shape, dtype=dtype, device=dev)
    ptr = lambda t: t.data_ptr()
    st = torch.cuda.current_stream(dev).cuda_stream
    px, py = e(B, S, T + 1), e(B, S + 1, T)
    ws = torch.empty(max(lib.frn_simple_logprobs_workspace_bytes(B, S, T, C),
                         lib.frn_pruned_logprobs_workspace_bytes(B, S, T, R),
                         lib.frn_mi_workspace_bytes(B, S, T, T + 1)), dtype=torch.uint8, device=dev)
    p_, pg = e(B, S + 1, T + 1), e(B, S + 1, T + 1)
    ans, ag, gx, gy = e(B), e(B), e(B, S, T + 1), e(B, S + 1, T)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    def ref_op():
        assert ref.ref_fast_rnnt_loss(ptr(px), ptr(py), ptr(bd), B, S, T, T + 1, 1, ptr(p_), ptr(ans), ptr(pg),
                                      ptr(gx), ptr(gy), ptr(ag), T + 1, st) == 1

    def our_op():
        chk(lib.frn_mi_fwd_bwd(ptr(px), ptr(py), ptr(bd), B, S, T, T + 1, 1, ptr(ans), ptr(gx), ptr(gy), ptr(ws),
                               ws.numel(), st), "mi_fwd_bwd")

    out = {"what": "reference FastRNNTLoss op (its own kernels, sm_100 build, op.cc launch sequence incl. sync) vs "
                   "frn_mi_fwd_bwd on the same dense px/py; host wall clock per call over %d back-to-back calls, ms" % reps}
    # lattice of the simple loss
    chk(lib.frn_simple_logprobs(ptr(lm), ptr(am), ptr(sym), ptr(bd), B, S, T, C, C - 1, 0, 0, 0.0, 0.0, ptr(px),
                                ptr(py), ptr(ws), ws.numel(), st), "simple_logprobs")
    out["simple_lattice"] = {"reference_ms": timed(ref_op), "ours_ms": timed(our_op)}
    # dense lattice of the pruned loss, as the reference runs it (rnnt_loss.py:968-1018, 1116-1119)
    chk(lib.frn_pruned_logprobs(ptr(pipe.logits), 1 if pipe.logits_bf16 else 0, ptr(sym), ptr(pipe.ranges), ptr(bd), B, S, T, R, C, C - 1, 0,
                                ptr(px), ptr(py), ptr(ws), ws.numel(), st), "pruned_logprobs")
    # ... against the product path for that lattice: the recursion on the band itself.  pxc/pyc are the
    # first two [B][T][R] float32 segments of frn_pruned_loss's workspace (left there by pipe.step)
    seg = (B * T * R * 4 + 255) // 256 * 256
    pxc, pyc = pipe.ws_pruned.data_ptr(), pipe.ws_pruned.data_ptr() + seg
    gxc, gyc = e(B, T, R), e(B, T, R)
    bws = torch.empty(max(lib.frn_band_mi_workspace_bytes(B, S, T, R), 256), dtype=torch.uint8, device=dev)

    def our_band():
        chk(lib.frn_band_mi_fwd_bwd(pxc, pyc, ptr(pipe.ranges), ptr(bd), B, S, T, R, 0, 0.0, 1, ptr(ans), ptr(gxc),
                                    ptr(gyc), ptr(bws), bws.numel(), st), "band_mi_fwd_bwd")

    out["pruned_lattice"] = {"reference_ms": timed(ref_op), "ours_ms": timed(our_band),
                             "ours_dense_kernels_ms": timed(our_op),
                             "note": "reference: its dense kernels on the [B,S,T+1] lattice it builds from the band; "
                                     "ours: frn_band_mi_fwd_bwd on the band (what frn_pruned_loss runs)"}
    # the two Cummin op calls of _adjust_pruning_lower_bound (rnnt_loss.py:628,634)
    x = torch.randint(0, S, (B, T), dtype=torch.int32, device=dev)
    y = torch.empty_like(x)
    out["cummin_x2"] = {"reference_ms": 2 * timed(lambda: ref.ref_cummin(ptr(x), ptr(y), B, T, st)),
                        "ours_ms": 2 * timed(lambda: chk(lib.frn_cummin(ptr(x), ptr(y), B, T, st), "cummin"))}
    r = sum(v["reference_ms"] for v in out.values() if isinstance(v, dict))
    o = sum(v["ours_ms"] for v in out.values() if isinstance(v, dict))
    out["reference_ops_per_step_ms"] = r
    out["ours_same_ops_ms"] = o
    out["ratio"] = r / o
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML (same data as the
    nvidia-smi clocks line of B200_PROFILING.md, but fast enough for a short
    timed region)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.err = None
        self.ready = threading.Event()      # set once NVML is initialised (or has failed)

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may remap
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                tok = vis.split(",")[self.index]
                h = nv.nvmlDeviceGetHandleByUUID(tok) if tok.startswith("GPU-") else nv.nvmlDeviceGetHandleByIndex(int(tok))
            else:
                h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons",
                                  getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
            self.ready.set()
            while not self.stop_flag:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    bits = int(get_reasons(h))
                    for bit, n in names.items():
                        if bits & bit:
                            self.reasons.add(n)
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
        finally:
            self.ready.set()

    def summary(self):
        s = sorted(self.samples)
        out = {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(s)}
        if self.err:
            out["error"] = self.err
        return out


def bind_to_gpu_numa_node(local):
    """Pin this process (and, by first touch, the pinned host buffers it allocates afterwards) to the
    CPUs of the NUMA node its GPU hangs off: with one rank per GPU the H2D copies of the end-to-end leg
    then read local memory instead of all ranks pulling from one socket.  Best effort; returns the node
    or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:  # noqa: BLE001
        pass
    return None


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, T, S, C, R = WORKLOADS[args.workload]
    logits_bf16 = args.workload == "c4" and not args.fp32_logits     # BASELINE.json configs[3]: bf16 joiner logits
    pipe = Pipeline(B, T, S, C, R, dev, nsplit=args.streams, overlap=args.overlap and args.streams <= 1, fuse_add=not args.no_fuse_add and not args.overlap,
                    am_side=args.am_side, am_side_at=args.am_side_at, logits_bf16=logits_bf16, training=args.profile_training)

    # rotating input sets: 4 x (am+lm) = 154 MB > 126 MB L2, and the step itself
    # streams ~1.1 GB of intermediates, so no iteration finds its inputs in L2
    # Every input set is ONE contiguous block (am | lm | symbols | boundary, 256-byte aligned parts): the
    # end-to-end leg moves a step's inputs with one cudaMemcpyAsync from one pinned block.
    NSETS = 4
    shapes = [((B, T, C), np.float32), ((B, S + 1, C), np.float32), ((B, S), np.int32), ((B, 4), np.int32)]
    offs, total = [], 0
    for shp, dt in shapes:
        offs.append(total)
        total += (int(np.prod(shp)) * np.dtype(dt).itemsize + 255) // 256 * 256

    def views(block):
        tdt = {np.float32: torch.float32, np.int32: torch.int32}
        return [block[o:o + int(np.prod(shp)) * 4].view(tdt[dt]).view(shp) for o, (shp, dt) in zip(offs, shapes)]

    host_blocks, host_sets, dev_sets = [], [], []
    for i in range(NSETS):
        hb = torch.empty(total, dtype=torch.uint8).pin_memory()
        for v, a in zip(views(hb), synth(B, T, S, C, 1234 + 17 * rank + i)):
            v.copy_(torch.from_numpy(a))
        host_blocks.append(hb)
        host_sets.append(views(hb))
        dev_sets.append(views(hb.to(dev)))
    host_out = torch.empty(2, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    # reduction='sum' across ranks: one 2-float NCCL all-reduce per step.  It runs on its own stream on
    # a copy of the two sums, so the next step's kernels never wait for the collective (only for the
    # 8-byte copy); the timed regions end with a barrier + synchronize over all streams.
    red_stream = torch.cuda.Stream(dev) if world > 1 else None
    red_bufs = [torch.zeros(2, dtype=torch.float32, device=dev) for _ in range(2)]
    red_copied = torch.cuda.Event()
    red_n = [0]

    def allreduce_losses():
        """-> (tensor holding the completed sums, stream it becomes valid on)"""
        main = torch.cuda.current_stream(dev)
        if world == 1:
            return pipe.losses, main
        buf = red_bufs[red_n[0] % 2]
        red_n[0] += 1
        red_stream.wait_stream(main)
        with torch.cuda.stream(red_stream):
            buf.copy_(pipe.losses)
            red_copied.record(red_stream)
            dist.all_reduce(buf)
        main.wait_event(red_copied)
        return buf, red_stream

    # kernels per step, counted by the library itself while one step is enqueued
    # (the counter lives in the -DFRN_DEBUG_HOOKS build only; everything timed below runs the product library)
    pipe._lib.use_debug_hooks(True)
    n0 = pipe.lib.frn_kernel_launches()
    pipe.step(*dev_sets[0])
    kernels_per_step = int(pipe.lib.frn_kernel_launches() - n0)
    torch.cuda.synchronize()
    pipe._lib.use_debug_hooks(False)

    # ---- CUDA graphs of one step per input set (launch-bound inner loop) ----
    use_graph = not args.no_graph
    graphs = []
    side = torch.cuda.Stream(dev)
    if use_graph:
        try:
            with torch.cuda.stream(side):
                pipe.step(*dev_sets[0])      # warm the lazily-set function attributes outside capture
                side.synchronize()
                for i in range(NSETS):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        pipe.step(*dev_sets[i])
                    graphs.append(g)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed ({e}); direct launches", file=sys.stderr)
            graphs, use_graph = [], False

    def run_step(i):
        if use_graph:
            graphs[i % NSETS].replay()
        else:
            pipe.step(*dev_sets[i % NSETS])
        allreduce_losses()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    for i in range(args.warmup):
        run_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(30.0)  # NVML initialised before the timed region starts
    sampler.samples.clear()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        run_step(i)
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * B / (ms_per_step * 1e-3)
    loss_check = (red_bufs[(red_n[0] - 1) % 2] if world > 1 else pipe.losses).cpu().tolist()

    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of the losses ----
    # Direct C-ABI launches (no graph).  Two landing buffers: the H2D copy of step i+1 (copy stream)
    # overlaps the kernels of step i; the host reads the loss of step i-1 while step i is in flight,
    # so every step's inputs cross PCIe and every step's result reaches the host inside the timed region.
    NBUF = 2
    copy_stream = torch.cuda.Stream(dev)
    stage_blocks = [torch.empty(total, dtype=torch.uint8, device=dev) for _ in range(NBUF)]
    stage_bufs = [views(blk) for blk in stage_blocks]
    host_outs = [host_out] + [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(NBUF - 1)]
    h2d_done = [torch.cuda.Event() for _ in range(NBUF)]
    step_done = [torch.cuda.Event() for _ in range(NBUF)]
    read_back = []

    def e2e_step(i):
        k = i % NBUF
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(step_done[k])           # landing buffer k is free again
            stage_blocks[k].copy_(host_blocks[i % NSETS], non_blocking=True)     # one cudaMemcpyAsync per step
            h2d_done[k].record(copy_stream)
        cur.wait_event(h2d_done[k])
        pipe.step(*stage_bufs[k])
        res, res_stream = allreduce_losses()
        with torch.cuda.stream(res_stream):
            host_outs[k].copy_(res, non_blocking=True)
            step_done[k].record(res_stream)
        if i > 0:                                           # the caller reads the previous step's loss
            step_done[(i - 1) % NBUF].synchronize()
            read_back.append(float(host_outs[(i - 1) % NBUF][0]))

    def e2e_drain(i_last):
        step_done[i_last % NBUF].synchronize()
        read_back.append(float(host_outs[i_last % NBUF][0]))

    for ev in step_done:
        ev.record(torch.cuda.current_stream(dev))
    for i in range(args.warmup):
        e2e_step(i)
    e2e_drain(args.warmup - 1)
    barrier()
    read_back.clear()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e2e_drain(args.steps - 1)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    assert len(read_back) == args.steps and all(np.isfinite(read_back))
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t.item()) / args.steps
    h2d = total

    # ---- H2D ceiling: the bare copies of the end-to-end leg (same pinned blocks, same bytes, all ranks at
    #      once, nothing else running) - what `e2e` can reach at best on this host ----
    barrier()
    nrep_c = max(10, min(args.steps, 100))
    with torch.cuda.stream(copy_stream):
        for i in range(3):
            stage_blocks[i % NBUF].copy_(host_blocks[i % NSETS], non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        c0.record(copy_stream)
        for i in range(nrep_c):
            stage_blocks[i % NBUF].copy_(host_blocks[i % NSETS], non_blocking=True)
        c1.record(copy_stream)
    barrier()
    t = torch.tensor([c0.elapsed_time(c1) / nrep_c], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d_ceiling_ms = float(t.item())

    def time_variant(step_fn):
        """ms per step of another arrangement of the step (CUDA graph per input set when graphs are on)."""
        vgraphs = []
        with torch.cuda.stream(side):
            step_fn(*dev_sets[0])
            side.synchronize()
            if use_graph:
                for i in range(NSETS):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        step_fn(*dev_sets[i])
                    vgraphs.append(g)
        torch.cuda.synchronize()
        vrun = (lambda i: vgraphs[i % NSETS].replay()) if vgraphs else (lambda i: step_fn(*dev_sets[i % NSETS]))
        for i in range(3):
            vrun(i)
        torch.cuda.synchronize()
        fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = max(20, min(args.steps, 200))
        fa.record()
        for i in range(nrep):
            vrun(i)
        fb.record()
        torch.cuda.synchronize()
        return fa.elapsed_time(fb) / nrep

    # ---- extra (not the headline): the same step with the fused additive joiner of SURVEY.md §8(f)-2,
    #      frn_pruned_add_joiner, in place of do_rnnt_pruning + add: am_pruned / lm_pruned never exist ----
    fused_ms = None
    if world == 1 and not pipe.logits_bf16:
        try:
            lib, chk = pipe.lib, pipe._lib.check
            ptr = lambda t: t.data_ptr()
            sfn = lambda: torch.cuda.current_stream(dev).cuda_stream

            def fused_step(am, lm, sym, bd):
                st = {name: fn for name, _, fn in pipe.stages(am, lm, sym, bd)}
                st["simple_loss"](); st["prune_ranges"]()
                chk(lib.frn_pruned_add_joiner(ptr(am), ptr(lm), ptr(pipe.ranges), B, S, T, R, C, 0, ptr(pipe.logits),
                                              sfn()), "pruned_add_joiner")
                st["pruned_loss"](); st["reduce"]()

            fused_ms = time_variant(fused_step)
        except Exception as e:  # noqa: BLE001
            fused_ms = repr(e)

    # ---- extra: the TRAINING step = the step above + its backward to am / lm: A9 (frn_simple_loss_bwd, the two
    #      gradient contractions on tcgen05) and the gradient of do_rnnt_pruning + additive joiner
    #      (frn_do_pruning_bwd fed with the logits gradient) ----
    training = None
    if world == 1 and not pipe.logits_bf16 and not args.no_training and not pipe.training:
        try:
            tpipe = Pipeline(B, T, S, C, R, dev, fuse_add=pipe.fuse_add, training=True)
            t_ms = time_variant(tpipe.step)
            t_stage = {}
            for name, nbytes, fn in tpipe.stages(*dev_sets[0]):
                if name not in ("simple_loss_bwd", "do_pruning_bwd"):
                    fn()
                    continue
                fn(); fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    fn()
                b.record()
                torch.cuda.synchronize()
                t_stage[name] = (a.elapsed_time(b) / 10, nbytes)
            flops = 2 * 2.0 * B * (S + 1) * T * C                   # the two contractions, useful flops
            training = {
                "ms_per_step": t_ms, "utterances_per_s": B / (t_ms * 1e-3), "launch": "cuda_graph" if use_graph else "direct",
                "what": "step + frn_simple_loss_bwd (A9: am / lm gradients, contractions on tcgen05) + frn_do_pruning_bwd "
                        "(gradient of do_rnnt_pruning and the additive joiner from the logits gradient)",
                "stages_ms": {k: round(v[0], 4) for k, v in t_stage.items()},
                "stages_gbs": {k: round(v[1] / (v[0] * 1e-3) / 1e9, 1) for k, v in t_stage.items()},
                "a9_useful_tflops": flops / (t_stage["simple_loss_bwd"][0] * 1e-3) / 1e12,
            }
            del tpipe
        except Exception as e:  # noqa: BLE001
            training = {"error": repr(e)}

    # ---- the dependency-chain kernel alone (t >= (S+T+1) t_step whatever the bytes): the wavefront
    #      recursion of frn_mi_fwd_bwd on this lattice, isolated with the debug-hooks build's FRN_MI_PHASE ----
    chain_ms = None
    if world == 1:
        try:
            lib, chk = pipe.lib, pipe._lib.check
            px = torch.randn((B, S, T + 1), device=dev) - 3.0
            py = torch.randn((B, S + 1, T), device=dev) - 3.0
            pipe._lib.use_debug_hooks(True)
            os.environ["FRN_DP_CHAIN"] = "1"
            wsz = int(lib.frn_mi_workspace_bytes(B, S, T, T + 1))
            ws = torch.empty(max(wsz, 256), dtype=torch.uint8, device=dev)
            ans = torch.empty(B, device=dev)
            call = lambda: chk(lib.frn_mi_fwd_bwd(px.data_ptr(), py.data_ptr(), dev_sets[0][3].data_ptr(), B, S, T, T + 1, 1,
                                                  ans.data_ptr(), pipe.gx.data_ptr(), pipe.gy.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), torch.cuda.current_stream(dev).cuda_stream), "mi")
            call()
            os.environ["FRN_MI_PHASE"] = "2"
            call(); call()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                call()
            b.record()
            torch.cuda.synchronize()
            chain_ms = a.elapsed_time(b) / 20
        except Exception as e:  # noqa: BLE001
            chain_ms = repr(e)
        finally:
            os.environ.pop("FRN_MI_PHASE", None)
            os.environ.pop("FRN_DP_CHAIN", None)
            pipe._lib.use_debug_hooks(False)

    # ---- per-stage device times (outside the timed region) -> roofline of the dominant kernel ----
    stage_ms = {}
    for name, nbytes, fn in pipe.stages(*dev_sets[0]):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        stage_ms[name] = (a.elapsed_time(b) / reps, nbytes)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # Roofline of the dominant kernel: the stages that are ONE HBM-bound kernel each, timed live with
    # CUDA events above (DESIGN.md lists every kernel with its bound; the lattice recursion kernels are
    # dependency-chain bound and are reported in `stages_ms` / `lattice_cells_per_s` instead).
    kernel_of_stage = {"do_pruning": "do_pruning_vec_kernel<8, 1, 1, 0>", "add_joiner": "add_kernel",
                       "do_pruning+add_joiner": "do_pruning_vec_kernel<8, 1, 1, 1>",
                       "do_pruning+add_joiner(lm,logits)": "do_pruning_vec_kernel<8, 0, 1, 1>",
                       "add_joiner(bf16)": "pruned_add_joiner_vec_bf16_kernel<8>"}
    hbm_stages = {k: v for k, v in stage_ms.items() if k in kernel_of_stage}
    dom = max(hbm_stages, key=lambda k: hbm_stages[k][0])
    dom_ms, dom_bytes = hbm_stages[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # dram__bytes_{read,write}.sum per launch (ncu --set full)
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get(kernel_of_stage[dom], {}).get("dram_bytes_per_launch")
    cells = B * (S + 1) * (T + 1)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: " + ("TRAINING STEP (--profile-training): " if pipe.training else "")
                        + "full pruned pipeline (simple fwd+bwd -> prune ranges -> pruning"
                        + (" + additive joiner in one pass, am_pruned / lm_pruned / logits all written"
                           if pipe.fuse_add else (" -> additive joiner writing bf16 logits" if pipe.logits_bf16
                                                  else " -> additive joiner"))
                        + f" -> pruned loss fwd+bwd) B={B}/GPU T={T} S={S} C={C} s_range={R} "
                        + ("fp32 am/lm, bf16 joiner logits and logits gradient, fp32 losses" if pipe.logits_bf16 else "fp32")
                        + " regular sum",
            "launch": "cuda_graph" if use_graph else "direct",
            "streams": f"{max(1, len(pipe.parts))} sub-batch stream(s) per step"
                       + ("; am half of do_rnnt_pruning on a second stream beside the simple loss" if pipe.overlap and not pipe.parts else "")
                       + ((f"; am_pruned (the half of do_rnnt_pruning that does not depend on the ranges) broadcast by "
                           f"{pipe.am_side} copy-engine CTAs on a second stream, "
                           + ("forked behind the normaliser: beside the lattice recursion (frn_simple_loss_bcast)"
                              if pipe.am_side_at == "chain" else "beside the whole simple loss"))
                          if pipe.am_side and not pipe.parts else ""),
            "l2": f"{NSETS} rotating input sets ({NSETS * h2d / 1e6:.0f} MB) + ~1.1 GB of streamed intermediates per step (> 126 MB L2)",
            "sharding": ("utterances sharded across ranks, one 2-float NCCL all-reduce per step; rank 0 bound to NUMA "
                         f"node {numa_node}") if world > 1 else "single GPU",
            "loss_check": loss_check,
        },
        "clocks": sampler.summary(),
        "e2e": {"value": world * B / (e2e_ms_per_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms_per_step, "wall_ms_per_step": wall / args.steps * 1e3,
                "pipelining": "H2D of step i+1 (ONE cudaMemcpyAsync of one pinned block) overlaps the kernels of step i "
                              "(2 landing blocks); direct C-ABI launches, loss of every step read on the host",
                "h2d_ceiling": {"ms_per_step": h2d_ceiling_ms, "gbs_per_gpu": h2d / (h2d_ceiling_ms * 1e-3) / 1e9,
                                "value": world * B / (h2d_ceiling_ms * 1e-3), "unit": UNIT,
                                "what": "the same per-step copies alone, all ranks at once (max over ranks): "
                                        "the most any end-to-end pipeline can reach on this host"},
                "frac_of_h2d_ceiling": h2d_ceiling_ms / e2e_ms_per_step},
        "gpu_launches": kernels_per_step * args.steps,
        "kernels_per_step": kernels_per_step,
        "roofline": {"bound": "hbm", "kernel": kernel_of_stage[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes": dom_bytes, "kernel_ms": dom_ms},
        "stages_ms": {k: round(v[0], 4) for k, v in stage_ms.items()},
        "stages_gbs": {k: round(v[1] / (v[0] * 1e-3) / 1e9, 1) for k, v in stage_ms.items()},
        "lattice_cells_per_s": cells / (stage_ms["simple_loss"][0] * 1e-3),
    }
    if training is not None:
        line["training_step"] = training
    if isinstance(chain_ms, float):
        steps_dp = S + T + 1                       # dependent steps of one direction (both run concurrently)
        mhz = line["clocks"].get("sm_mhz") or 1965.0
        cps = chain_ms * 1e-3 * mhz * 1e6 / steps_dp
        line["latency_roofline"] = {
            "kernel": "dp_chain_kernel<1> (wavefront lattice recursion, forward and backward chains concurrently)",
            "steps": steps_dp, "kernel_ms": chain_ms, "cycles_per_step": cps, "floor_cycles": 41.0,
            "frac": 41.0 / cps, "sm_mhz": mhz,
            "floor_source": "scripts/ubench/chain_latency: one step's dependent (shuffle, multiply-add) chain on this part",
            "what": "t >= (S+T+1) t_step per direction whatever the bytes (SURVEY.md 8d); frac = floor / measured"}
    elif chain_ms is not None:
        line["latency_roofline"] = {"error": chain_ms}
    if fused_ms is not None:
        line["fused_joiner_variant"] = {
            "ms_per_step": fused_ms, "launch": "cuda_graph" if use_graph else "direct",
            "utterances_per_s": (B / (fused_ms * 1e-3)) if isinstance(fused_ms, float) else None,
            "what": "same step with frn_pruned_add_joiner instead of do_rnnt_pruning + add (SURVEY.md 8f-2); "
                    "not the headline: the reference API materialises am_pruned / lm_pruned"}
    if world == 1 and not args.no_cpu:
        sample_B = 8 if args.workload == "c2" else None
        base, _ = cpu_bench(args.workload, 2, 1, sample_B)
        line["cpu_baseline"] = base
    if world == 1 and not args.no_ref_gpu:
        try:
            for _, _, fn in pipe.stages(*dev_sets[0]):   # whole batch, one stream: leaves ranges / logits /
                fn()                                     # band log-probs of this input set in the buffers
            ref_leg = reference_gpu_op_leg(pipe, *dev_sets[0])
            if ref_leg is not None:
                line["reference_gpu_op"] = ref_leg
        except Exception as e:  # noqa: BLE001  (a baseline leg must never take the bench line down)
            line["reference_gpu_op"] = {"error": repr(e)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def c5_boundaries(B=256, T=1500, S=400, seed=5):
    """BASELINE.json configs[4] / SURVEY.md 8(d): T_b ~ U[200,1500], S_b ~ U[20, min(400, T_b)]."""
    rng = np.random.default_rng(seed)
    bd = np.zeros((B, 4), np.int32)
    bd[:, 3] = rng.integers(200, T + 1, B)
    bd[:, 2] = np.minimum(rng.integers(20, S + 1, B), bd[:, 3])
    return bd


def run_c5(args):
    """configs[4]: ONE ragged batch of 256 utterances (C=500, s_range=5, reduction=sum), sharded by utterance
    over the ranks (greedy LPT on lattice cells, tf_fast_rnnt.sharding.partition_batch) - strong scaling.
    Each rank cuts its shard into length buckets (sharding.plan_buckets) so that the bandwidth-bound kernels
    do not stream padding and every bucket gets the kernel variants of its own shape; the buckets are
    independent sub-batches, so they run on their own CUDA streams (the dependency-chain-bound kernels of one
    bucket beside the bandwidth-bound kernels of another), the whole step is one CUDA graph, and the two sums
    are completed with one all-reduce."""
    import torch
    import torch.distributed as dist
    from tf_fast_rnnt.sharding import partition_batch, plan_buckets

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, C, R = 256, 500, 5
    bd_all = c5_boundaries(B)
    mine = partition_batch(bd_all, world)[rank]
    buckets = plan_buckets(bd_all[mine], R, C, max_buckets=args.buckets, min_bucket=args.min_bucket)
    nb = len(buckets)

    # one contiguous block holds every bucket's (am, lm, symbols, boundary): the end-to-end leg moves a step's
    # inputs with one cudaMemcpyAsync; two landing blocks on the device + the resident set the timed loop uses
    f32, i32 = torch.float32, torch.int32
    layout, total_bytes = [], 0
    for bk in buckets:
        Bk, Sk, Tk = len(bk["idx"]), bk["S_max"], bk["T_max"]
        parts = []
        for shp, dt in (((Bk, Tk, C), f32), ((Bk, Sk + 1, C), f32), ((Bk, Sk), i32), ((Bk, 4), i32)):
            parts.append((total_bytes, shp, dt))
            total_bytes += (int(np.prod(shp)) * 4 + 255) // 256 * 256
        layout.append(parts)

    def views(block):
        return [[block[o:o + int(np.prod(shp)) * 4].view(dt).view(shp) for o, shp, dt in parts] for parts in layout]

    host_block = torch.empty(total_bytes, dtype=torch.uint8).pin_memory()
    gen = torch.Generator()
    gen.manual_seed(1234 + rank)
    for bk, (am, lm, sym, bd) in zip(buckets, views(host_block)):
        idx = mine[bk["idx"]]
        am.copy_(torch.randn(am.shape, generator=gen))
        lm.copy_(torch.randn(lm.shape, generator=gen))
        sym.copy_(torch.randint(0, C - 1, sym.shape, generator=gen, dtype=i32))
        bd.copy_(torch.from_numpy(bd_all[idx].copy()))
    dev_blocks = [host_block.to(dev) for _ in range(3)]         # [0] resident set, [1], [2] landing blocks
    inputs = [views(blk) for blk in dev_blocks]
    pipes = [Pipeline(len(bk["idx"]), bk["T_max"], bk["S_max"], C, R, dev, fuse_add=not args.no_fuse_add) for bk in buckets]
    streams = [torch.cuda.Stream(dev) for _ in range(nb)] if args.c5_streams else []
    totals = [torch.zeros(2, dtype=torch.float32, device=dev) for _ in range(3)]
    torch.cuda.synchronize()

    def compute(k):
        """the shard's pipelines on input set k -> totals[k] (two sums over the shard)"""
        main = torch.cuda.current_stream(dev)
        if streams:
            for st_, pipe, inp in zip(streams, pipes, inputs[k]):
                st_.wait_stream(main)
                with torch.cuda.stream(st_):
                    pipe.step(*inp)
            for st_ in streams:
                main.wait_stream(st_)
        else:
            for pipe, inp in zip(pipes, inputs[k]):
                pipe.step(*inp)
        totals[k].zero_()
        for pipe in pipes:
            totals[k].add_(pipe.losses)

    pipes[0]._lib.use_debug_hooks(True)
    n0 = pipes[0].lib.frn_kernel_launches()
    compute(0)
    kernels_per_step = int(pipes[0].lib.frn_kernel_launches() - n0)
    torch.cuda.synchronize()
    pipes[0]._lib.use_debug_hooks(False)

    use_graph = not args.no_graph
    graphs = [None] * 3
    if use_graph:
        try:
            side = torch.cuda.Stream(dev)
            with torch.cuda.stream(side):
                compute(0)
                side.synchronize()
                for k in range(3):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        compute(k)
                    graphs[k] = g
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed ({e}); direct launches", file=sys.stderr)
            graphs, use_graph = [None] * 3, False

    def step(k=0):
        if use_graph:
            graphs[k].replay()
        else:
            compute(k)
        if world > 1:
            dist.all_reduce(totals[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(30.0)
    sampler.samples.clear()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.stop_flag = True

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    loss_check = totals[0].cpu().tolist()

    # ---- end to end: the shard's inputs from one pinned block every step (one cudaMemcpyAsync into one of two
    #      landing blocks on a copy stream, overlapping the previous step's kernels), the two sums read back ----
    copy_stream = torch.cuda.Stream(dev)
    host_outs = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    step_done = [torch.cuda.Event() for _ in range(2)]
    read_back = []
    n_e2e = max(3, min(args.steps, 20))

    def e2e_step(i):
        k = i % 2
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(step_done[k])
            dev_blocks[1 + k].copy_(host_block, non_blocking=True)
            h2d_done[k].record(copy_stream)
        cur.wait_event(h2d_done[k])
        step(1 + k)
        host_outs[k].copy_(totals[1 + k], non_blocking=True)
        step_done[k].record(cur)
        if i > 0:
            step_done[(i - 1) % 2].synchronize()
            read_back.append(float(host_outs[(i - 1) % 2][0]))

    for ev in step_done:
        ev.record(torch.cuda.current_stream(dev))
    for i in range(2):
        e2e_step(i)
    step_done[1].synchronize()
    barrier()
    read_back.clear()
    e0.record()
    for i in range(n_e2e):
        e2e_step(i)
    step_done[(n_e2e - 1) % 2].synchronize()
    read_back.append(float(host_outs[(n_e2e - 1) % 2][0]))
    e1.record()
    barrier()
    assert len(read_back) == n_e2e and all(np.isfinite(read_back))
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
    # the bare copies alone (ceiling of any end-to-end pipeline on this host)
    with torch.cuda.stream(copy_stream):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev_blocks[1].copy_(host_block, non_blocking=True)
        c0.record(copy_stream)
        for i in range(10):
            dev_blocks[1 + i % 2].copy_(host_block, non_blocking=True)
        c1.record(copy_stream)
    barrier()
    h2d_ceiling_ms = max_over_ranks(c0.elapsed_time(c1) / 10)

    # ---- per-stage device times of the largest bucket -> roofline of the dominant HBM-bound kernel ----
    big = max(range(nb), key=lambda j: pipes[j].B * pipes[j].T)
    stage_ms = {}
    for name, nbytes, fn in pipes[big].stages(*inputs[0][big]):
        fn(); fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        stage_ms[name] = (a.elapsed_time(b) / 10, nbytes)
    frames = int(bd_all[mine, 3].sum())
    padded = int(sum(bk["padded_frames"] for bk in buckets))
    if rank == 0:
        peak, peak_src = measured_peaks()
        kernel_of_stage = {"do_pruning": "do_pruning_vec_kernel<8, 1, 1, 0>", "add_joiner": "add_kernel",
                           "do_pruning+add_joiner": "do_pruning_vec_kernel<8, 1, 1, 1>"}
        hbm = {k: v for k, v in stage_ms.items() if k in kernel_of_stage}
        dom = max(hbm, key=lambda k: hbm[k][0])
        dom_ms, dom_bytes = hbm[dom]
        line = {
            "metric": METRIC, "value": B / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c5: ragged batch B=256 (T_b 200-1500, S_b 20-400) C=500 s_range=5 fp32 regular sum, "
                                   "full pruned pipeline per length bucket, batch sharded by utterance over the ranks",
                       "launch": "cuda_graph" if use_graph else "direct",
                       "streams": f"{nb} bucket stream(s) per rank" if streams else "1 stream per rank",
                       "sharding": f"greedy LPT on lattice cells, {len(mine)} utterances on rank 0",
                       "buckets_rank0": [{"utterances": int(len(bk["idx"])), "T_max": bk["T_max"], "S_max": bk["S_max"]}
                                         for bk in buckets],
                       "frames_rank0": frames, "padded_frames_rank0": padded,
                       "l2": "every step streams > 1 GB per rank (> 126 MB L2)", "loss_check": loss_check},
            "clocks": sampler.summary(),
            "e2e": {"value": B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": total_bytes, "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_ms,
                    "pipelining": "H2D of step i+1 (one cudaMemcpyAsync of the rank's pinned block) overlaps the kernels of "
                                  "step i (2 landing blocks); loss of every step read on the host",
                    "h2d_ceiling": {"ms_per_step": h2d_ceiling_ms, "gbs_per_gpu": total_bytes / (h2d_ceiling_ms * 1e-3) / 1e9,
                                    "what": "the same per-step copies alone, all ranks at once (max over ranks)"},
                    "frac_of_h2d_ceiling": h2d_ceiling_ms / e2e_ms},
            "gpu_launches": kernels_per_step * args.steps, "kernels_per_step": kernels_per_step,
            "roofline": {"bound": "hbm", "kernel": kernel_of_stage[dom] + f" (largest bucket of rank 0: B={pipes[big].B} "
                                                                          f"T={pipes[big].T} S={pipes[big].S})",
                         "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes": dom_bytes, "kernel_ms": dom_ms},
            "stages_ms_largest_bucket": {k: round(v[0], 4) for k, v in stage_ms.items()},
        }
        if world == 1 and not args.no_cpu:
            base, _ = cpu_bench("c5", 2, 1, 2)
            line["cpu_baseline"] = base
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--buckets", type=int, default=16, help="c5: length buckets per rank (1 = pad the shard to its maxima; 4+: the planner may also split by label length)")
    ap.add_argument("--min-bucket", type=int, default=2, help="c5: fewest utterances a length bucket may hold")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--c5-streams", type=int, default=1, help="c5: 1 = every length bucket on its own CUDA stream (default), 0 = one stream")
    ap.add_argument("--streams", type=int, default=1,
                    help="run a step as this many independent sub-batches on separate CUDA streams")
    ap.add_argument("--overlap", action="store_true",
                    help="run the am half of do_rnnt_pruning on a second stream beside the simple loss (measured: "
                         "0.380 ms/step against 0.372 without - the copy slows the latency-bound kernels it overlaps)")
    ap.add_argument("--am-side", type=int, default=64,
                    help="G > 0: am half of do_rnnt_pruning by frn_broadcast_am_pruned (G persistent copy-engine CTAs) on a "
                         "second stream beside the simple loss")
    ap.add_argument("--am-side-at", default="chain", choices=["start", "chain"],
                    help="--am-side: start the broadcast with the step (beside the whole simple loss) or fork it behind "
                         "the normaliser so that it runs beside the lattice recursion only (frn_simple_loss_bcast)")
    ap.add_argument("--no-fuse-add", action="store_true",
                    help="do_rnnt_pruning and the additive joiner as two passes (frn_do_pruning, frn_add_joiner) "
                         "instead of the one-pass frn_do_pruning_add_joiner")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-training", action="store_true", help="skip the training-step leg")
    ap.add_argument("--profile-training", action="store_true",
                    help="the timed step IS the training step (step + A9 + pruning gradient): for ncu launch lists; not a headline run")
    ap.add_argument("--fp32-logits", action="store_true", help="c4 with float32 joiner logits (default there: bf16)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the reference's own CUDA op")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
