"""Clock trace of one softmax warp of attn_fwd (block 0, warp 4) at B=256: average cycles per phase of a tile step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200 import _lib as L
B = 256; T = B * 1024; dev = "cuda"
qkv = torch.randn(T, 384, device=dev).to(torch.bfloat16)
out = torch.empty(T, 128, dtype=torch.bfloat16, device=dev); lse = torch.empty(B * 4096, device=dev)
bits = torch.empty(L.lib.focr_mha_drop_bits_bytes(B) // 4, dtype=torch.int32, device=dev)
trace = torch.zeros(8192, dtype=torch.int64, device=dev)
st = L.cur_stream()
for p, bp in ((0.1, bits.data_ptr()), (0.0, None)):
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, p, 1, 0, bp, st))
    torch.cuda.synchronize()
    trace.zero_()
    L.check(L.lib.focr_attn_set_trace(trace.data_ptr()))
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, p, 1, 0, bp, st))
    torch.cuda.synchronize()
    L.check(L.lib.focr_attn_set_trace(None))
    t = trace.cpu().tolist()
    per_tile = 16 * 2 + 16 * 10 + 2
    names = ["s_full", "ld0+", "cmp0", "p_empty", "st0+ld1+rel", "cmp1", "(none)", "st1", "publish", "loop"]
    print(f"p={p}: total cycles tile0 {t[per_tile - 1] - t[0]}, all 4 tiles {t[4 * per_tile - 1] - t[0]}")
    for it in range(4):
        base = it * per_tile
        p1 = t[base: base + 32]
        w1 = sum(p1[2 * j + 1] - p1[2 * j] for j in range(16)) / 16
        step1 = (p1[30] - p1[0]) / 15
        p2 = t[base + 32: base + 32 + 160]
        seg = [0.0] * 10
        for j in range(16):
            e = p2[10 * j: 10 * j + 10]
            nxt = p2[10 * j + 10] if j < 15 else t[base + 32 + 160]
            for k in range(9):
                seg[k] += (e[k + 1] - e[k]) / 16
            seg[9] += (nxt - e[9]) / 16
        step2 = (p2[150] - p2[0]) / 15
        ep = t[base + 32 + 160: base + 32 + 162]
        print(f" tile {it}: pass1 step {step1:.0f} (s_full wait {w1:.0f}) | pass2 step {step2:.0f}: " +
              " ".join(f"{n}={v:.0f}" for n, v in zip(names, seg)) + f" | o_full wait {ep[1] - ep[0]}")
