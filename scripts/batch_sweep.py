"""TBSRN train step (CUDA-graph replay) at the per-GPU batches of the strong-scaling curve (global 256 over N = 1, 2, 4, 8 GPUs):
ms/step and the speed-up a perfect exchange would give, t(256) / t(256 / N).  Usage: python scripts/batch_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fudanocr_b200.model.tbsrn import TBSRN
from fudanocr_b200.trainer import TBSRNTrainer

torch.manual_seed(1234)
res = {}
for B in (256, 128, 64, 32):
    m = TBSRN().cuda().train()
    tr = TBSRNTrainer(m)
    lr, hr = torch.rand(B, 3, 16, 64, device="cuda"), torch.rand(B, 3, 32, 128, device="cuda")
    for _ in range(5):
        tr.step(lr, hr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tr.step(lr, hr)
    e1.record()
    torch.cuda.synchronize()
    res[B] = e0.elapsed_time(e1) / 20
    del tr, m
    torch.cuda.empty_cache()
for B, t in res.items():
    print(f"batch {B:4d}: {t:7.3f} ms/step  {B / t * 1e3:8.0f} img/s/GPU   compute-only strong speed-up at N = {256 // B}: {res[256] / t:.2f}x")
