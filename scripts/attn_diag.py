"""Attention kernels: component-wise errors against a torch fp32 reference (small B) and CUDA-event timings at B=256.
Usage: python scripts/attn_diag.py [--time-only]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fudanocr_b200 import _lib as L
from oracle import dropout_rng as R

dev = "cuda"


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def ref_attn(qkv, B, keep, scale):
    q, k, v = [qkv[:, i * 128:(i + 1) * 128].view(B, 1024, 4, 32).transpose(1, 2) for i in range(3)]
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(32)
    p = F.softmax(s, dim=-1)
    if keep is not None:
        p = p * keep.to(p.dtype) * scale
    return torch.matmul(p, v).transpose(1, 2).reshape(B * 1024, 128)


def check(B, p_drop, use_bits):
    T = B * 1024
    g = torch.Generator(device=dev).manual_seed(B)
    qkv = (torch.randn(T, 384, device=dev, generator=g) * 1.2).to(torch.bfloat16)
    out = torch.zeros(T, 128, dtype=torch.bfloat16, device=dev)
    lse = torch.zeros(B * 4096, device=dev)
    bits = torch.zeros(L.lib.focr_mha_drop_bits_bytes(B) // 4, dtype=torch.int32, device=dev) if use_bits else None
    bp = bits.data_ptr() if use_bits else None
    st = L.cur_stream()
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, p_drop, 1234, 6, bp, st))
    L.check(L.lib.focr_sync_check(st))
    keep = R.attn_keep_mask(B, 1234, 3, p_drop).to(dev) if p_drop > 0 else None
    qr = qkv.float().requires_grad_(True)
    ref = ref_attn(qr, B, keep, R.attn_keep_scale(p_drop))
    msg = [f"B={B} p={p_drop} bits={use_bits}: out {rel(out, ref):.2e}"]
    q, k = [qkv.float()[:, i * 128:(i + 1) * 128].view(B, 1024, 4, 32).transpose(1, 2) for i in range(2)]
    lse_ref = torch.logsumexp(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(32), -1) / math.log(2.0)
    msg.append(f"lse maxabs {(lse.view(B, 4, 1024) - lse_ref).abs().max().item():.2e}")
    if use_bits and p_drop > 0:
        w = bits.view(B, 4, 32, 1024).long() & 0xFFFFFFFF
        kk = torch.arange(32, device=dev)
        sh = kk // 4 + 8 * (kk & 3)
        m = ((w[..., None] >> sh) & 1).permute(0, 1, 3, 2, 4).reshape(B, 4, 1024, 1024)
        msg.append(f"bits mismatches {(m.bool() != keep.bool()).sum().item()} keep-rate {m.float().mean().item():.4f}")
    d_out = torch.randn(T, 128, device=dev, generator=g).to(torch.bfloat16)
    ref.backward(d_out.float())
    dqkv = torch.zeros_like(qkv)
    bws = torch.empty(L.lib.focr_mha_bwd_workspace_bytes(B), dtype=torch.uint8, device=dev)
    L.check(L.lib.focr_mha_flash_bwd(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr(), bws.data_ptr(), bws.numel(),
                                     dqkv.data_ptr(), B, p_drop, 1234, 6, bp, st))
    L.check(L.lib.focr_sync_check(st))
    for i, nm in enumerate("qkv"):
        msg.append(f"d{nm} {rel(dqkv[:, i * 128:(i + 1) * 128], qr.grad[:, i * 128:(i + 1) * 128]):.2e}")
    # per-head / per-row-block breakdown of dK to localise layout faults
    dk, dkr = dqkv[:, 128:256].float().view(B, 1024, 4, 32), qr.grad[:, 128:256].view(B, 1024, 4, 32)
    blk = [(rel(dk[0, i * 64:(i + 1) * 64, 0], dkr[0, i * 64:(i + 1) * 64, 0])) for i in range(16)]
    msg.append("dk blocks(b0,h0) " + " ".join(f"{e:.1e}" for e in blk))
    print(" | ".join(msg), flush=True)


def timing(B=256, reps=5):
    T = B * 1024
    qkv = torch.randn(T, 384, device=dev).to(torch.bfloat16)
    out = torch.empty(T, 128, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B * 4096, device=dev)
    dout = torch.randn(T, 128, device=dev).to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    bws = torch.empty(L.lib.focr_mha_bwd_workspace_bytes(B), dtype=torch.uint8, device=dev)
    bits = torch.empty(L.lib.focr_mha_drop_bits_bytes(B) // 4, dtype=torch.int32, device=dev)
    st = L.cur_stream()
    for p, bp, tag in ((0.0, None, "no dropout"), (0.1, bits.data_ptr(), "dropout 0.1 + keep bits"), (0.1, None, "dropout 0.1 re-hash")):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for r in range(reps + 2):
            ev[0].record()
            L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, p, 1, 0, bp, st))
            ev[1].record()
            L.check(L.lib.focr_mha_flash_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), bws.data_ptr(), bws.numel(),
                                             dqkv.data_ptr(), B, p, 1, 0, bp, st))
            ev[2].record()
            torch.cuda.synchronize()
            if r >= 2:
                tf += ev[0].elapsed_time(ev[1])
                tb += ev[1].elapsed_time(ev[2])
        print(f"B={B} {tag}: fwd {tf / reps:.3f} ms  bwd(dq+dkv) {tb / reps:.3f} ms", flush=True)
    L.lib.focr_prof_enable(1, None)
    for r in range(3):
        L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, 0.1, 1, 0, bits.data_ptr(), st))
        L.check(L.lib.focr_mha_flash_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), bws.data_ptr(), bws.numel(),
                                         dqkv.data_ptr(), B, 0.1, 1, 0, bits.data_ptr(), st))
    print({k: (c, round(ms / c, 4)) for k, (c, ms) in L.prof_collect().items()}, flush=True)


if __name__ == "__main__":
    if "--time-only" not in sys.argv:
        check(1, 0.0, False)
        check(2, 0.1, True)
        check(1, 0.1, False)
    timing()
