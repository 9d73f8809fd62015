#!/usr/bin/env python
"""Micro-benchmark of the HBM-bound normalisation kernels at the TBSRN shapes (T = 256*1024 tokens)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200 import _lib as L
DEV = "cuda"
T = 256 * 1024
lib = L.lib
st = L.cur_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=DEV)
def bench(name, fn, nbytes):
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    print(f"{name:14s} {t*1e3:7.1f} us  {nbytes/t/1e6:6.0f} GB/s  ({nbytes/1e6:.0f} MB)")
bf = lambda *s: torch.randn(*s, device=DEV).to(torch.bfloat16)
x128, dy128, y128 = bf(T, 128), bf(T, 128), bf(T, 128)
a, b = torch.ones(128, device=DEV), torch.zeros(128, device=DEV)
da, db = torch.empty(128, device=DEV), torch.empty(128, device=DEV)
bench("ln_fwd", lambda: L.check(lib.focr_layernorm_std_fwd(x128.data_ptr(), a.data_ptr(), b.data_ptr(), y128.data_ptr(), T, 1e-6, st)), T * 128 * 2 * 2)
bench("ln_bwd", lambda: L.check(lib.focr_layernorm_std_bwd(dy128.data_ptr(), x128.data_ptr(), a.data_ptr(), y128.data_ptr(), da.data_ptr(), db.data_ptr(), T, 1e-6, ws.data_ptr(), ws.numel(), st)), T * 128 * 2 * 3)
x64, dy64, y64 = bf(T, 64), bf(T, 64), bf(T, 64)
g, be = torch.ones(64, device=DEV), torch.zeros(64, device=DEV)
rm, rv = torch.zeros(64, device=DEV), torch.ones(64, device=DEV)
nbt = torch.zeros((), dtype=torch.long, device=DEV)
stats = torch.empty(4, 64, device=DEV)
dg, dbt = torch.empty(64, device=DEV), torch.empty(64, device=DEV)
bench("bn_fwd(mish)", lambda: L.check(lib.focr_bn_train_fwd(x64.data_ptr(), g.data_ptr(), be.data_ptr(), rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(), y64.data_ptr(), stats.data_ptr(), T, 64, 1, ws.data_ptr(), ws.numel(), st)), T * 64 * 2 * 3)
bench("bn_bwd(mish)", lambda: L.check(lib.focr_bn_bwd(dy64.data_ptr(), x64.data_ptr(), stats.data_ptr(), y64.data_ptr(), dg.data_ptr(), dbt.data_ptr(), T, 64, 1, ws.data_ptr(), ws.numel(), st)), T * 64 * 2 * 5)
bench("bn_bwd(none)", lambda: L.check(lib.focr_bn_bwd(dy64.data_ptr(), x64.data_ptr(), stats.data_ptr(), y64.data_ptr(), dg.data_ptr(), dbt.data_ptr(), T, 64, 0, ws.data_ptr(), ws.numel(), st)), T * 64 * 2 * 5)
