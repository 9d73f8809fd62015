"""Decoder cross-attention at the 32x320 shape (B = 64, h = 4, d_k = 256, Tq = 30, Tk = 2560): device time of the row-split forward
and the row pass + key pass backward.  Usage: python scripts/mha_large.py [Tk] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fudanocr_b200.model import recog_ops as ops

Tk = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H, dk, Tq = 4, 256, 30
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
q = torch.randn(B * Tq, H * dk, device=dev, generator=g).to(torch.bfloat16)
k = torch.randn(B * Tk, H * dk, device=dev, generator=g).to(torch.bfloat16)
v = torch.randn(B * Tk, H * dk, device=dev, generator=g).to(torch.bfloat16)
do = torch.randn(B * Tq, H * dk, device=dev, generator=g).to(torch.bfloat16)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


t_f, (out, amap) = timed(lambda: ops.mha_fwd(q, k, v, B, H, dk, Tq, Tk, 0, 0.1, 5, 2))
t_b, _ = timed(lambda: ops.mha_bwd(q, k, v, do, amap, B, H, dk, Tq, Tk, 0, 0.1))
flop = 2.0 * B * H * Tq * Tk * dk
print(f"Tk={Tk} B={B}: fwd {t_f:.3f} ms ({2 * flop / t_f / 1e9:.1f} TFLOP/s), bwd {t_b:.3f} ms ({5 * flop / t_b / 1e9:.1f} TFLOP/s)")
