#!/usr/bin/env python
"""frn_broadcast_am_pruned alone at the c2 shape for several grid sizes: what G copy-engine CTAs move per second."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tf-fast-rnnt_b200"))
import torch
from tf_fast_rnnt import _lib
lib, chk = _lib.lib, _lib.check
B, T, R, C = 32, 500, 5, 500
dev = torch.device("cuda", 0)
am = torch.randn(B, T, C, device=dev)
out = torch.empty(B, T, R, C, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
for G in (8, 16, 20, 32, 48, 64, 84, 148, 296):
    chk(lib.frn_broadcast_am_pruned(am.data_ptr(), B, T, R, C, out.data_ptr(), G, st), "bc")
    torch.cuda.synchronize()
    ok = bool((out == am[:, :, None, :]).all())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        chk(lib.frn_broadcast_am_pruned(am.data_ptr(), B, T, R, C, out.data_ptr(), G, st), "bc")
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    nbytes = am.numel() * 4 * (1 + R)
    print(f"G={G:4d}  {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s  {nbytes/ms/1e6/G:6.1f} GB/s/CTA  exact={ok}")
    out.zero_()
