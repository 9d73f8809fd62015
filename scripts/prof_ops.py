"""Representative single launches of the hot kernels at the B=256 shapes (for `ncu --set full`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200 import _lib as L
B = 256; T = B * 1024
dev = "cuda"
def ws(n): return torch.empty(int(n), dtype=torch.uint8, device=dev)
qkv = (torch.randn(T, 384, device=dev) * 1.0).to(torch.bfloat16)
out = torch.empty(T, 128, dtype=torch.bfloat16, device=dev); lse = torch.empty(B * 4096, device=dev)
dout = torch.randn(T, 128, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); bws = torch.empty(L.lib.focr_mha_bwd_workspace_bytes(B), dtype=torch.uint8, device='cuda')
x128 = torch.randn(T, 128, device=dev).to(torch.bfloat16)
w = torch.randn(384, 128, device=dev) / 11; b = torch.randn(384, device=dev)
y = torch.empty(T, 384, dtype=torch.bfloat16, device=dev)
x64 = torch.randn(B, 16, 64, 64, device=dev).to(torch.bfloat16)
wc = torch.randn(64, 64, 3, 3, device=dev) / 24; bc = torch.randn(64, device=dev); yc = torch.empty_like(x64)
wsl = ws(L.lib.focr_linear_workspace_bytes(128, 384)); wsc = ws(L.lib.focr_conv2d_workspace_bytes(64, 64, 3)); wsg = ws(L.lib.focr_wgrad_workspace_bytes())
dw = torch.empty(384, 128, device=dev); dwc = torch.empty(64, 64, 3, 3, device=dev)
bits = torch.empty(L.lib.focr_mha_drop_bits_bytes(B) // 4, dtype=torch.int32, device=dev)
st = L.cur_stream()
for rep in range(2):
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, 0.1, 1, 0, bits.data_ptr(), st))
    L.check(L.lib.focr_mha_flash_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), bws.data_ptr(), bws.numel(), dqkv.data_ptr(), B, 0.1, 1, 0, bits.data_ptr(), st))
    L.check(L.lib.focr_linear_fwd(x128.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 0, T, 128, 384, 0, wsl.data_ptr(), wsl.numel(), st))
    L.check(L.lib.focr_conv2d_fwd(x64.data_ptr(), wc.data_ptr(), bc.data_ptr(), yc.data_ptr(), 0, 0, B, 16, 64, 64, 64, 3, 0, wsc.data_ptr(), wsc.numel(), st))
    L.check(L.lib.focr_linear_wgrad(y.data_ptr(), x128.data_ptr(), dw.data_ptr(), T, 128, 384, wsg.data_ptr(), wsg.numel(), st))
    L.check(L.lib.focr_conv2d_wgrad(yc.data_ptr(), x64.data_ptr(), dwc.data_ptr(), B, 16, 64, 0, wsg.data_ptr(), wsg.numel(), st))
    torch.cuda.synchronize()
print("ok")
