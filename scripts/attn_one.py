"""One forward + backward launch of the attention kernels at B=256 with dropout + keep bits (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200 import _lib as L
B = int(os.environ.get("ATTN_B", "256")); T = B * 1024; dev = "cuda"
qkv = torch.randn(T, 384, device=dev).to(torch.bfloat16)
out = torch.empty(T, 128, dtype=torch.bfloat16, device=dev); lse = torch.empty(B * 4096, device=dev)
dout = torch.randn(T, 128, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); bws = torch.empty(L.lib.focr_mha_bwd_workspace_bytes(B), dtype=torch.uint8, device='cuda')
bits = torch.empty(L.lib.focr_mha_drop_bits_bytes(B) // 4, dtype=torch.int32, device=dev)
st = L.cur_stream()
for r in range(int(os.environ.get("ATTN_REPS", "1"))):
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, 0.1, 1, 0, bits.data_ptr(), st))
    L.check(L.lib.focr_mha_flash_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), bws.data_ptr(), bws.numel(), dqkv.data_ptr(), B, 0.1, 1, 0, bits.data_ptr(), st))
torch.cuda.synchronize(); print("ok")
