#!/usr/bin/env python
"""Micro-benchmark of the 64->64 3x3 conv (tcgen05 implicit GEMM) at the TBSRN shape (B=256, 16x64)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200 import _lib as L
B, H, W, C = 256, 16, 64, 64
x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
w = torch.randn(C, C, 3, 3, device="cuda") / 24
b = torch.randn(C, device="cuda")
y = torch.empty(B, H, W, C, dtype=torch.bfloat16, device="cuda")
ws = torch.empty(L.lib.focr_conv2d_workspace_bytes(C, C, 3), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(L.lib.focr_conv2d_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 0, 0, B, H, W, C, C, 3, 0, ws.data_ptr(), ws.numel(), L.cur_stream()))
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = min(ts[2:])
fl = 2.0 * B * H * W * 576 * 64
print(f"conv3x3 64->64 B=256: {t*1e3:.1f} us  {fl/t/1e9:.0f} TFLOP/s (incl. weight prep launch)")
