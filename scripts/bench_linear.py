#!/usr/bin/env python
"""Micro-benchmark of the token GEMMs (tcgen05) at the TBSRN shapes (T = 256*1024)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200 import _lib as L
T = 256 * 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for K, N, use_res in ((128, 384, False), (128, 128, True), (128, 128, False), (384, 128, True)):
    x = torch.randn(T, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    res = torch.randn(T, N, device="cuda").to(torch.bfloat16) if use_res else None
    y = torch.empty(T, N, dtype=torch.bfloat16, device="cuda")
    ws = torch.empty(L.lib.focr_linear_workspace_bytes(K, N), dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.focr_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), L.ptr(res), T, K, N, 0, ws.data_ptr(), ws.numel(), L.cur_stream()))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[2:])
    by = T * (K + N + (N if use_res else 0)) * 2
    print(f"linear K={K} N={N} res={use_res}: {t*1e3:.1f} us  {by/t/1e6:.0f} GB/s ({by/1e6:.0f} MB; incl. weight prep launch)")
