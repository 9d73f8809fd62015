#!/usr/bin/env python
"""Micro-benchmark of the fused linear weight+bias gradient (csrc/wgrad_tc.cu) at the TBSRN shapes (T = 256*1024)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200 import _lib as L

DEV = "cuda"
T, K = 256 * 1024, 128
ws = torch.empty(L.lib.focr_wgrad_workspace_bytes(), dtype=torch.uint8, device=DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for N, ld in ((64, 64), (128, 128), (384, 384)):
    dy = torch.randn(T, ld, device=DEV).to(torch.bfloat16)
    x = torch.randn(T, K, device=DEV).to(torch.bfloat16)
    dw = torch.empty(N, K, device=DEV); db = torch.empty(N, device=DEV)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.focr_linear_wgrad_bias(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), T, K, N, ws.data_ptr(), ws.numel(), L.cur_stream()))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    by = T * (N + K) * 2
    t = min(ts[1:])
    print(f"N={N}: {t*1e3:.1f} us  {by/t/1e6:.0f} GB/s (algorithmic {by/1e6:.0f} MB)")
