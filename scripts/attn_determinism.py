import os, sys
sys.path.insert(0, "/root/repo")
import torch
from fudanocr_b200 import _lib as L
B = 8; T = B * 1024; dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(T, 384, device=dev, generator=g).to(torch.bfloat16)
dout = torch.randn(T, 128, device=dev, generator=g).to(torch.bfloat16)
st = L.cur_stream()
res = []
for rep in range(6):
    out = torch.zeros(T, 128, dtype=torch.bfloat16, device=dev); lse = torch.zeros(B * 4096, device=dev)
    dqkv = torch.zeros_like(qkv); bws = torch.empty(L.lib.focr_mha_bwd_workspace_bytes(B), dtype=torch.uint8, device='cuda')
    bits = torch.zeros(L.lib.focr_mha_drop_bits_bytes(B) // 4, dtype=torch.int32, device=dev)
    p = 0.1 if rep >= 3 else 0.0
    bp = bits.data_ptr() if p > 0 else None
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, p, 5, 2, bp, st))
    L.check(L.lib.focr_mha_flash_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), bws.data_ptr(), bws.numel(), dqkv.data_ptr(), B, p, 5, 2, bp, st))
    torch.cuda.synchronize()
    res.append((out.clone(), lse.clone(), dqkv.clone()))
for a, b in ((0, 1), (1, 2), (3, 4), (4, 5)):
    print(a, b, [torch.equal(x, y) for x, y in zip(res[a], res[b])], [(x.float() - y.float()).abs().max().item() for x, y in zip(res[a], res[b])])
