"""Tensor-level wrappers over the C-ABI building blocks the trainable recognisers are assembled from (csrc/recog_ops.cu,
tc_gemm.cu, elementwise.cu).  Each function allocates its outputs with torch (caching allocator),
enqueues the kernels on the current stream and returns tensors; none of them computes anything in torch.  Activations are
bf16: feature maps NHWC (B, H, W, C), token matrices (rows, C).  CUDA tensors only - there is no CPU path."""
from __future__ import annotations

import os

import torch

from .. import _lib as L

BF = torch.bfloat16
# weight gradients of the encoder convs on the tcgen05 GEMM (transposed operands) instead of the streaming mma.sync kernel
TC_WGRAD = os.environ.get("FOCR_TC_WGRAD", "1") == "1"   # tuning knob: 0 = streaming kernel everywhere


def _dev(t):
    if not t.is_cuda:
        raise L.FocrError("focr recogniser ops run on CUDA tensors only (no CPU fallback)")
    return t.device


def require_cuda(t):
    _dev(t)


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def _call(fn, what, *args):
    L.check(fn(*args), what)


def _out(t, shape, dev):
    """the caller's gradient buffer (a view of a trainer's flat gradient: the kernel then writes the parameter gradient in place,
    no autograd accumulation pass) or a fresh fp32 tensor"""
    if t is None:
        return torch.empty(shape, dtype=torch.float32, device=dev)
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != int(torch.Size(shape).numel()):
        raise ValueError(f"gradient sink {tuple(t.shape)} {t.dtype} does not match a contiguous fp32 {tuple(shape)}")
    return t


# ---- convolutions ---------------------------------------------------------------------------------------------------
def conv_first_fwd(x_nchw, w, b):
    """3-channel fp32 NCHW image -> (B, H, W, Co) bf16 through im2col + tcgen05 GEMM"""
    dev = _dev(x_nchw)
    B, Ci, H, W = x_nchw.shape
    Co = w.shape[0]
    y = torch.empty(B, H, W, Co, dtype=BF, device=dev)
    ws = _ws(L.lib.focr_conv3x3_gemm_workspace_bytes(B, H, W, Ci, Co), dev)
    _call(L.lib.focr_conv3x3_gemm_fwd, "conv3x3_gemm_fwd", 0, x_nchw.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, H, W,
          Ci, Co, ws.data_ptr(), ws.numel(), L.cur_stream())
    return y


def conv_first_wgrad(dy, x_nchw, w_shape, out_w=None, out_b=None):
    dev = _dev(dy)
    B, Ci, H, W = x_nchw.shape
    Co = w_shape[0]
    dw = _out(out_w, w_shape, dev)
    db = _out(out_b, (Co,), dev)
    ws = _ws(L.lib.focr_conv3x3_gemm_workspace_bytes(B, H, W, Ci, Co), dev)
    _call(L.lib.focr_conv3x3_gemm_wgrad, "conv3x3_gemm_wgrad", dy.data_ptr(), 0, x_nchw.data_ptr(), dw.data_ptr(), db.data_ptr(), B,
          H, W, Ci, Co, ws.data_ptr(), ws.numel(), L.cur_stream())
    return dw, db


def conv_fwd(x, w, b):
    """(B, H, W, Ci) bf16 -> (B, H, W, Co) bf16, implicit GEMM on tcgen05 (Ci, Co multiples of 64; W in {16, 32, 64, 128})"""
    dev = _dev(x)
    B, H, W, Ci = x.shape
    Co = w.shape[0]
    y = torch.empty(B, H, W, Co, dtype=BF, device=dev)
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(Ci, Co, 3), dev)
    _call(L.lib.focr_conv2d_fwd, "conv2d_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 0, 0, B, H, W, Ci, Co, 3, 0,
          ws.data_ptr(), ws.numel(), L.cur_stream())
    return y


def conv_dgrad(dy, w):
    dev = _dev(dy)
    B, H, W, Co = dy.shape
    Ci = w.shape[1]
    dx = torch.empty(B, H, W, Ci, dtype=BF, device=dev)
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(Ci, Co, 3), dev)
    _call(L.lib.focr_conv2d_dgrad, "conv2d_dgrad", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, W, Ci, Co, 3, 0, ws.data_ptr(),
          ws.numel(), L.cur_stream())
    return dx


def conv_wgrad(dy, x, w_shape, out_w=None, out_b=None):
    """dW, db of a 3x3 convolution from dy (B,H,W,Co) and the layer input x (B,H,W,Ci): im2col, then dW = dY^T col on the tcgen05
    GEMM (focr_conv3x3_wgrad_tc: Co % 128 == 0 and B*H*W % 128 == 0) or on the streaming mma.sync kernel (the 64-channel stem)"""
    dev = _dev(dy)
    B, H, W, Ci = x.shape
    Co = w_shape[0]
    dw = _out(out_w, w_shape, dev)
    db = _out(out_b, (Co,), dev)
    if TC_WGRAD and Co % 128 == 0 and Ci % 64 == 0 and (B * H * W) % 128 == 0:
        ws = _ws(L.lib.focr_conv3x3_wgrad_tc_workspace_bytes(B, H, W, Ci, Co), dev)
        _call(L.lib.focr_conv3x3_wgrad_tc, "conv3x3_wgrad_tc", dy.data_ptr(), x.data_ptr(), 0, dw.data_ptr(), db.data_ptr(), B, H, W,
              Ci, Co, ws.data_ptr(), ws.numel(), L.cur_stream())
        return dw, db
    ws = _ws(L.lib.focr_conv3x3_gemm_workspace_bytes(B, H, W, Ci, Co), dev)
    _call(L.lib.focr_conv3x3_gemm_wgrad, "conv3x3_gemm_wgrad", dy.data_ptr(), x.data_ptr(), 0, dw.data_ptr(), db.data_ptr(), B, H, W,
          Ci, Co, ws.data_ptr(), ws.numel(), L.cur_stream())
    return dw, db


# ---- BatchNorm / element-wise ------------------------------------------------------------------------------------------
ACT_NONE, ACT_RELU = 0, 2


def bn_train_fwd(x, gamma, beta, rm, rv, nbt, act):
    """x (..., C) bf16; updates the running buffers in place; returns (y, stats fp32 [4][C])"""
    dev = _dev(x)
    C = x.shape[-1]
    T = x.numel() // C
    y = torch.empty_like(x)
    stats = torch.empty(4, C, dtype=torch.float32, device=dev)
    ws = _ws(L.lib.focr_bn_workspace_bytes(), dev)
    _call(L.lib.focr_bn_train_fwd, "bn_train_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
          nbt.data_ptr(), y.data_ptr(), stats.data_ptr(), T, C, act, ws.data_ptr(), ws.numel(), L.cur_stream())
    return y, stats


def bn_eval_fwd(x, gamma, beta, rm, rv, act):
    dev = _dev(x)
    C = x.shape[-1]
    T = x.numel() // C
    y = torch.empty_like(x)
    stats = torch.empty(4, C, dtype=torch.float32, device=dev)
    _call(L.lib.focr_bn_eval_fwd, "bn_eval_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
          y.data_ptr(), stats.data_ptr(), T, C, act, L.cur_stream())
    return y


def bn_bwd(dy, x, stats, act, out_g=None, out_b=None):
    dev = _dev(dy)
    C = x.shape[-1]
    T = x.numel() // C
    dx = torch.empty_like(x)
    dg = _out(out_g, (C,), dev)
    db = _out(out_b, (C,), dev)
    ws = _ws(L.lib.focr_bn_workspace_bytes(), dev)
    _call(L.lib.focr_bn_bwd, "bn_bwd", dy.data_ptr(), x.data_ptr(), stats.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), T, C,
          act, ws.data_ptr(), ws.numel(), L.cur_stream())
    return dx, dg, db


def add_relu(a, b):
    _dev(a)
    y = torch.empty_like(a)
    _call(L.lib.focr_add_relu, "add_relu", a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), L.cur_stream())
    return y


def relu_bwd(dy, y):
    _dev(dy)
    dx = torch.empty_like(dy)
    _call(L.lib.focr_relu_bwd, "relu_bwd", dy.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.numel(), L.cur_stream())
    return dx


def maxpool_fwd(x):
    dev = _dev(x)
    B, H, W, C = x.shape
    y = torch.empty(B, H // 2, W // 2, C, dtype=BF, device=dev)
    _call(L.lib.focr_maxpool2x2_fwd, "maxpool2x2_fwd", x.data_ptr(), y.data_ptr(), B, H, W, C, L.cur_stream())
    return y


def maxpool_bwd(x, y, dy):
    _dev(x)
    B, H, W, C = x.shape
    dx = torch.empty_like(x)
    _call(L.lib.focr_maxpool2x2_bwd, "maxpool2x2_bwd", x.data_ptr(), y.data_ptr(), dy.data_ptr(), dx.data_ptr(), B, H, W, C,
          L.cur_stream())
    return dx


def dropout(x, p, seed, sid):
    _dev(x)
    y = torch.empty_like(x)
    _call(L.lib.focr_dropout, "dropout", x.data_ptr(), y.data_ptr(), x.numel(), float(p), int(seed), int(sid), L.cur_stream())
    return y


# ---- token-matrix linears ------------------------------------------------------------------------------------------------
def linear_fwd(x, w, b, relu=False, fp32_out=False):
    """x (M, K) bf16, w (N, K) fp32, b (N) fp32 -> (M, N) bf16 (or fp32); M % 128 == 0, K and N multiples of 64"""
    dev = _dev(x)
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.float32 if fp32_out else BF, device=dev)
    ws = _ws(L.lib.focr_linear_workspace_bytes(K, N), dev)
    _call(L.lib.focr_linear_fwd, "linear_fwd", x.data_ptr(), w.data_ptr(), L.ptr(b), y.data_ptr(), 0, M, K, N,
          (1 if relu else 0) | (4 if fp32_out else 0), ws.data_ptr(), ws.numel(), L.cur_stream())
    return y


def linear_dgrad(dy, w):
    dev = _dev(dy)
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty(M, K, dtype=BF, device=dev)
    ws = _ws(L.lib.focr_linear_workspace_bytes(K, N), dev)
    _call(L.lib.focr_linear_dgrad, "linear_dgrad", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), M, K, N, ws.data_ptr(), ws.numel(),
          L.cur_stream())
    return dx


def linear_wgrad(dy, x, out_w=None, out_b=None):
    """dw (N, K) fp32 = dy^T x and db (N) = column sums of dy"""
    dev = _dev(dy)
    M, N = dy.shape
    K = x.shape[1]
    dw = _out(out_w, (N, K), dev)
    db = _out(out_b, (N,), dev)
    ws = _ws(L.lib.focr_wgrad_workspace_bytes(), dev)
    _call(L.lib.focr_linear_wgrad, "linear_wgrad", dy.data_ptr(), x.data_ptr(), dw.data_ptr(), M, K, N, ws.data_ptr(), ws.numel(),
          L.cur_stream())
    _call(L.lib.focr_bias_grad, "bias_grad", dy.data_ptr(), db.data_ptr(), M, N, ws.data_ptr(), ws.numel(), L.cur_stream())
    return dw, db


# ---- decoder pieces --------------------------------------------------------------------------------------------------
def mha_fwd(q, k, v, B, H, dk, Tq, Tk, causal, p, seed, sid):
    """q (>= B*Tq rows, H*dk) / k, v (>= B*Tk rows, H*dk) bf16 (row-strided views allowed) ->
    out (rows of q, H*dk) bf16 with the padding rows zero, map fp32 (B, H, Tq, Tk)"""
    dev = _dev(q)
    out = torch.zeros(q.shape[0], H * dk, dtype=BF, device=dev)
    amap = torch.empty(B, H, Tq, Tk, dtype=torch.float32, device=dev)
    _call(L.lib.focr_mha_small_fwd, "mha_small_fwd", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
          out.data_ptr(), out.stride(0), amap.data_ptr(), B, H, dk, Tq, Tk, int(causal), float(p), int(seed), int(sid),
          L.cur_stream())
    return out, amap


def mha_bwd(q, k, v, d_out, amap, B, H, dk, Tq, Tk, causal, p):
    dev = _dev(q)
    dq = torch.zeros(q.shape[0], H * dk, dtype=BF, device=dev)
    dk_ = torch.zeros(k.shape[0], H * dk, dtype=BF, device=dev)
    dv = torch.zeros(v.shape[0], H * dk, dtype=BF, device=dev)
    ws = _ws(L.lib.focr_mha_small_bwd_workspace_bytes(B, H, Tq, Tk), dev)
    _call(L.lib.focr_mha_small_bwd_ws, "mha_small_bwd", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
          d_out.data_ptr(), d_out.stride(0), amap.data_ptr(), dq.data_ptr(), dq.stride(0), dk_.data_ptr(), dk_.stride(0),
          dv.data_ptr(), dv.stride(0), B, H, dk, Tq, Tk, int(causal), float(p), ws.data_ptr(), ws.numel(), L.cur_stream())
    return dq, dk_, dv


def ln_fwd(x, res, a, b, eps=1e-6):
    """y = LN(x + res) (reference LayerNorm: unbiased std, eps on the std); returns (x + res, y)"""
    _dev(x)
    T, C = x.shape
    xs = torch.empty_like(x)
    y = torch.empty_like(x)
    _call(L.lib.focr_layernorm_wide_fwd, "layernorm_wide_fwd", x.data_ptr(), res.data_ptr(), a.data_ptr(), b.data_ptr(), xs.data_ptr(),
          y.data_ptr(), T, C, eps, L.cur_stream())
    return xs, y


def ln_bwd(dy, xs, a, eps=1e-6, out_a=None, out_b=None):
    dev = _dev(dy)
    T, C = xs.shape
    dx = torch.empty_like(xs)
    da = _out(out_a, (C,), dev)
    db = _out(out_b, (C,), dev)
    ws = _ws(L.lib.focr_layernorm_wide_workspace_bytes(C), dev)
    _call(L.lib.focr_layernorm_wide_bwd, "layernorm_wide_bwd", dy.data_ptr(), xs.data_ptr(), a.data_ptr(), dx.data_ptr(), da.data_ptr(),
          db.data_ptr(), T, C, eps, ws.data_ptr(), ws.numel(), L.cur_stream())
    return dx, da, db


def embed_fwd(idx, lut, rows_pad, p, seed, sid):
    """idx (B, T) int64 -> (rows_pad, 2E) bf16 = [lut[idx] * sqrt(E) | dropout(pe[t])], padding rows zero"""
    dev = _dev(lut)
    B, T = idx.shape
    vocab, E = lut.shape
    out = torch.empty(rows_pad, 2 * E, dtype=BF, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _call(L.lib.focr_text_embed_fwd, "text_embed_fwd", idx.data_ptr(), lut.data_ptr(), vocab, E, B, T, rows_pad, out.data_ptr(),
          float(p), int(seed), int(sid), status.data_ptr(), L.cur_stream())
    if L.status_checks() and int(status.item()):   # nn.Embedding raises IndexError on such an index
        raise IndexError(f"text_embed_fwd: token index outside [0, {vocab}) (the kernel maps it to row 0)")
    return out


def embed_bwd(idx, d_out, vocab, E, out=None):
    dev = _dev(d_out)
    B, T = idx.shape
    d_lut = _out(out, (vocab, E), dev)
    _call(L.lib.focr_text_embed_bwd, "text_embed_bwd", idx.data_ptr(), d_out.data_ptr(), vocab, E, B, T, d_lut.data_ptr(),
          L.cur_stream())
    return d_lut


def packed_ce(logits, B, T, C, length, gt, gscale=1.0, want_grad=True):
    """logits fp32 (rows_pad, ld) -> (loss scalar tensor, d_logits bf16 (rows_pad, ld) or None)"""
    dev = _dev(logits)
    ld = logits.shape[1]
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d = torch.zeros(logits.shape, dtype=BF, device=dev) if want_grad else None
    ws = _ws(L.lib.focr_packed_ce_workspace_bytes(B), dev)
    _call(L.lib.focr_packed_ce, "packed_ce", logits.data_ptr(), ld, B, T, C, length.data_ptr(), gt.data_ptr(), float(gscale),
          loss.data_ptr(), L.ptr(d), ld, ws.data_ptr(), ws.numel(), L.cur_stream())
    return loss[0], d


def l2norm_fwd(x):
    """x fp32 (T, C) -> (y bf16 (T, C) = x / ||x||, inv fp32 (T))"""
    dev = _dev(x)
    T, C = x.shape
    y = torch.empty(T, C, dtype=BF, device=dev)
    inv = torch.empty(T, dtype=torch.float32, device=dev)
    _call(L.lib.focr_l2norm_rows_fwd, "l2norm_rows_fwd", x.data_ptr(), x.stride(0), y.data_ptr(), inv.data_ptr(), T, C, L.cur_stream())
    return y, inv


def l2norm_bwd(dy, y, inv):
    _dev(dy)
    T, C = y.shape
    dx = torch.empty(T, C, dtype=BF, device=dy.device)
    _call(L.lib.focr_l2norm_rows_bwd, "l2norm_rows_bwd", dy.data_ptr(), y.data_ptr(), inv.data_ptr(), dx.data_ptr(), C, T, C,
          L.cur_stream())
    return dx


def packed_feat_mse(y, B, T, length, gt, feats, gscale=1.0, want_grad=True):
    """mean over the valid rows (t < length[b]) and columns of (y - feats[gt])^2 -> (loss scalar tensor, d_y bf16 or None)"""
    dev = _dev(y)
    C = y.shape[1]
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d = torch.zeros(y.shape, dtype=BF, device=dev) if want_grad else None
    ws = _ws(L.lib.focr_packed_ce_workspace_bytes(B), dev)
    _call(L.lib.focr_packed_feat_mse, "packed_feat_mse", y.data_ptr(), B, T, C, length.data_ptr(), gt.data_ptr(), feats.data_ptr(),
          feats.shape[0], float(gscale), loss.data_ptr(), L.ptr(d), ws.data_ptr(), ws.numel(), L.cur_stream())
    return loss[0], d


def adadelta_step(table, n_chunks, gscale, lr, rho, eps, weight_decay):
    _call(L.lib.focr_adadelta_step, "adadelta_step", table.data_ptr(), n_chunks, float(gscale), float(lr), float(rho), float(eps),
          float(weight_decay), L.cur_stream())
