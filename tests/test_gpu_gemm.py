"""tcgen05 implicit-GEMM engine vs torch fp32 ops on bf16-rounded operands (GPU only)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from fudanocr_b200 import _lib as L
    return L


def _ws(nbytes):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device="cuda")


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def _nhwc(x):  # (B,C,H,W) fp32 -> (B,H,W,C) bf16 contiguous
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


@pytest.mark.parametrize("B,H,W,Ci,Co,ks", [(2, 16, 64, 64, 64, 3), (3, 16, 64, 64, 64, 1),
                                            (2, 16, 64, 128, 64, 3), (1, 32, 128, 64, 64, 3),
                                            (2, 16, 64, 64, 128, 3), (2, 16, 64, 64, 256, 3),
                                            # >= 2 tiles per SM: the resident-weights mode of the 64-wide kernel
                                            (40, 16, 64, 64, 64, 3), (11, 32, 128, 64, 64, 3), (37, 16, 64, 128, 64, 1)])
def test_conv_fwd(B, H, W, Ci, Co, ks):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, Ci, H, W, device="cuda", generator=g)
    w = torch.randn(Co, Ci, ks, ks, device="cuda", generator=g) / (Ci * ks * ks) ** 0.5
    b = torch.randn(Co, device="cuda", generator=g)
    xb = _nhwc(x)
    y = torch.empty(B, H, W, Co, dtype=torch.bfloat16, device="cuda")
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(Ci, Co, ks))
    L.check(L.lib.focr_conv2d_fwd(xb.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 0, 0, B, H, W, Ci, Co,
                                  ks, 0, ws.data_ptr(), ws.numel(), L.cur_stream()), "conv2d_fwd")
    L.check(L.lib.focr_sync_check(L.cur_stream()))
    # reference by unfold + matmul: cuDNN spends seconds runtime-compiling fp32 NHWC kernels for the larger shapes
    cols = F.unfold(xb.float().permute(0, 3, 1, 2).contiguous(), ks, padding=ks // 2)          # (B, Ci*ks*ks, H*W)
    ref = (w.to(torch.bfloat16).float().reshape(Co, -1) @ cols + b[None, :, None]).view(B, Co, H, W)
    err = _rel(y.permute(0, 3, 1, 2), ref)
    assert err < 1e-2, err


def test_conv_fwd_relu_residual():
    L = _lib()
    B, H, W, C = 2, 16, 64, 64
    x = torch.randn(B, C, H, W, device="cuda")
    w = torch.randn(C, C, 3, 3, device="cuda") / 24
    b = torch.randn(C, device="cuda")
    res = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    xb = _nhwc(x)
    y = torch.empty(B, H, W, C, dtype=torch.bfloat16, device="cuda")
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(C, C, 3))
    L.check(L.lib.focr_conv2d_fwd(xb.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 0, res.data_ptr(), B, H,
                                  W, C, C, 3, 1, ws.data_ptr(), ws.numel(), L.cur_stream()))
    L.check(L.lib.focr_sync_check(L.cur_stream()))
    ref = F.relu(F.conv2d(xb.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, padding=1))
    ref = ref + res.float().permute(0, 3, 1, 2)
    assert _rel(y.permute(0, 3, 1, 2), ref) < 1e-2


def test_conv_fwd_pixshuf_mish():
    L = _lib()
    B, H, W, Ci, Co = 2, 16, 64, 64, 256
    x = torch.randn(B, Ci, H, W, device="cuda")
    w = torch.randn(Co, Ci, 3, 3, device="cuda") / 24
    b = torch.randn(Co, device="cuda")
    xb = _nhwc(x)
    y = torch.empty(B, 2 * H, 2 * W, 64, dtype=torch.bfloat16, device="cuda")
    y2 = torch.empty_like(y)
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(Ci, Co, 3))
    L.check(L.lib.focr_conv2d_fwd(xb.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), y2.data_ptr(), 0, B, H,
                                  W, Ci, Co, 3, 2, ws.data_ptr(), ws.numel(), L.cur_stream()))
    L.check(L.lib.focr_sync_check(L.cur_stream()))
    pre = F.pixel_shuffle(F.conv2d(xb.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, padding=1), 2)
    assert _rel(y.permute(0, 3, 1, 2), pre) < 1e-2
    act = pre * torch.tanh(F.softplus(pre))
    assert _rel(y2.permute(0, 3, 1, 2), act) < 1e-2


@pytest.mark.parametrize("Ci,Co,ks,shuf", [(64, 64, 3, 0), (64, 64, 1, 0), (64, 256, 3, 1), (64, 128, 3, 0)])
def test_conv_dgrad(Ci, Co, ks, shuf):
    L = _lib()
    B, H, W = 2, 16, 64
    w = torch.randn(Co, Ci, ks, ks, device="cuda") / (Co * ks * ks) ** 0.5
    wq = w.to(torch.bfloat16).float()
    if shuf:
        dy = torch.randn(B, 64, 2 * H, 2 * W, device="cuda")
        dyb = _nhwc(dy)  # (B,2H,2W,64)
        dy_conv = F.pixel_unshuffle(dyb.float().permute(0, 3, 1, 2), 2)
    else:
        dy = torch.randn(B, Co, H, W, device="cuda")
        dyb = _nhwc(dy)
        dy_conv = dyb.float().permute(0, 3, 1, 2)
    ref = F.conv_transpose2d(dy_conv, wq, padding=ks // 2)
    dx = torch.empty(B, H, W, Ci, dtype=torch.bfloat16, device="cuda")
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(Ci, Co, ks))
    L.check(L.lib.focr_conv2d_dgrad(dyb.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, W, Ci, Co, ks, 2 * shuf,
                                    ws.data_ptr(), ws.numel(), L.cur_stream()))
    L.check(L.lib.focr_sync_check(L.cur_stream()))
    assert _rel(dx.permute(0, 3, 1, 2), ref) < 1e-2


@pytest.mark.parametrize("M,K,N,flags", [(1024, 128, 384, 0), (2048, 128, 128, 1), (1024, 128, 64, 0),
                                         (128, 64, 128, 0), (1024, 128, 128, 4)])
def test_linear_fwd(M, K, N, flags):
    L = _lib()
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(torch.bfloat16) if not (flags & 4) else None
    y = torch.empty(M, N, dtype=torch.float32 if flags & 4 else torch.bfloat16, device="cuda")
    ws = _ws(L.lib.focr_linear_workspace_bytes(K, N))
    L.check(L.lib.focr_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), L.ptr(res), M, K, N,
                                  flags, ws.data_ptr(), ws.numel(), L.cur_stream()))
    L.check(L.lib.focr_sync_check(L.cur_stream()))
    ref = x.float() @ w.to(torch.bfloat16).float().t() + b
    if flags & 1:
        ref = F.relu(ref)
    if res is not None:
        ref = ref + res.float()
    assert _rel(y, ref) < 1e-2


def test_linear_dgrad():
    L = _lib()
    M, K, N = 2048, 128, 384
    dy = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda") / N ** 0.5
    dx = torch.empty(M, K, dtype=torch.bfloat16, device="cuda")
    ws = _ws(L.lib.focr_linear_workspace_bytes(K, N))
    L.check(L.lib.focr_linear_dgrad(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), M, K, N, ws.data_ptr(), ws.numel(),
                                    L.cur_stream()))
    L.check(L.lib.focr_sync_check(L.cur_stream()))
    ref = dy.float() @ w.to(torch.bfloat16).float()
    assert _rel(dx, ref) < 1e-2
