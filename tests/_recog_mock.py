"""TEST INFRASTRUCTURE ONLY.  torch-fp32 CPU stand-ins with the signatures of fudanocr_b200/model/recog_ops.py, so that the
ASSEMBLY of the recogniser (autograd wiring, layouts, parameter mapping, padding rows, packing) can be checked against the
oracle on the GPU-less build box.  The kernels themselves are checked on the B200 (tests/test_gpu_recog_ops.py,
tests/test_gpu_sld.py); nothing in the product imports this file."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import dropout_rng as R

BF = torch.float32
ACT_NONE, ACT_RELU = 0, 2


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _into(out, t):
    """the kernels' output-buffer contract (recog_ops._out): write the result into the caller's gradient view when one is given"""
    if out is None:
        return t
    out.copy_(t.reshape(out.shape))
    return out


def conv_first_fwd(x, w, b):
    return _nhwc(F.conv2d(x, w, b, padding=1))


def conv_first_wgrad(dy, x, wshape, out_w=None, out_b=None):
    g = _nchw(dy)
    return _into(out_w, torch.nn.grad.conv2d_weight(x, wshape, g, padding=1)), _into(out_b, g.sum((0, 2, 3)))


def conv_fwd(x, w, b):
    return _nhwc(F.conv2d(_nchw(x), w, b, padding=1))


def conv_dgrad(dy, w):
    return _nhwc(F.conv_transpose2d(_nchw(dy), w, padding=1))


def conv_wgrad(dy, x, wshape, out_w=None, out_b=None):
    g = _nchw(dy)
    return _into(out_w, torch.nn.grad.conv2d_weight(_nchw(x), wshape, g, padding=1)), _into(out_b, g.sum((0, 2, 3)))


def bn_train_fwd(x, gamma, beta, rm, rv, nbt, act):
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    mean, var = x2.mean(0), x2.var(0, unbiased=False)
    n = x2.shape[0]
    with torch.no_grad():
        rm.mul_(0.9).add_(0.1 * mean)
        rv.mul_(0.9).add_(0.1 * var * n / (n - 1))
        nbt.add_(1)
    invstd = (var + 1e-5).rsqrt()
    scale = gamma * invstd
    shift = beta - mean * scale
    y = x * scale + shift
    if act == ACT_RELU:
        y = F.relu(y)
    return y, torch.stack([mean, invstd, scale, shift]).detach()


def bn_eval_fwd(x, gamma, beta, rm, rv, act):
    scale = gamma * (rv + 1e-5).rsqrt()
    y = x * scale + (beta - rm * scale)
    return F.relu(y) if act == ACT_RELU else y


def bn_bwd(dy, x, stats, act, out_g=None, out_b=None):
    mean, invstd, scale, shift = stats
    if act == ACT_RELU:
        dy = dy * ((x * scale + shift) > 0)
    C = x.shape[-1]
    xh = ((x - mean) * invstd).reshape(-1, C)
    g = dy.reshape(-1, C)
    dgamma, dbeta = (g * xh).sum(0), g.sum(0)
    n = g.shape[0]
    dx = scale * (g - dbeta / n - xh * dgamma / n)
    return dx.reshape(x.shape), _into(out_g, dgamma), _into(out_b, dbeta)


def add_relu(a, b):
    return F.relu(a + b)


def relu_bwd(dy, y):
    return dy * (y > 0)


def maxpool_fwd(x):
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))


def maxpool_bwd(x, y, dy):
    xr = _nchw(x).detach().clone().requires_grad_(True)
    with torch.enable_grad():
        F.max_pool2d(xr, 2, 2).backward(_nchw(dy))
    return _nhwc(xr.grad)


def _keep(n, p, seed, sid):
    return torch.from_numpy(R._keep(R.drop_key(seed, sid), np.arange(n, dtype=np.uint64), R.thresh16(p)))


def dropout(x, p, seed, sid):
    if p <= 0:
        return x.clone()
    return x * _keep(x.numel(), p, seed, sid).reshape(x.shape) * R.keep_scale(p)


def linear_fwd(x, w, b, relu=False, fp32_out=False):
    y = F.linear(x, w, b)   # b may be None
    return F.relu(y) if relu else y


def linear_dgrad(dy, w):
    return dy @ w


def linear_wgrad(dy, x, out_w=None, out_b=None):
    return _into(out_w, dy.t() @ x), _into(out_b, dy.sum(0))


def _heads(t, B, T, H, dk):
    return t[:B * T].reshape(B, T, H, dk).transpose(1, 2)


def _attn(q, k, v, B, H, dk, Tq, Tk, causal, keep, ks):
    s = _heads(q, B, Tq, H, dk) @ _heads(k, B, Tk, H, dk).transpose(-1, -2) / math.sqrt(dk)
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Tq, Tk, dtype=torch.bool), 1), float("-inf"))
    pm = F.softmax(s, -1)
    if keep is not None:
        pm = pm * keep * ks
    return (pm @ _heads(v, B, Tk, H, dk)).transpose(1, 2).reshape(B * Tq, H * dk), pm


def mha_fwd(q, k, v, B, H, dk, Tq, Tk, causal, p, seed, sid):
    keep = _keep(B * H * Tq * Tk, p, seed, sid).reshape(B, H, Tq, Tk) if p > 0 else None
    o, pm = _attn(q, k, v, B, H, dk, Tq, Tk, causal, keep, R.keep_scale(p) if p > 0 else 1.0)
    out = torch.zeros(q.shape[0], H * dk)
    out[:B * Tq] = o
    return out, pm.contiguous()


def mha_bwd(q, k, v, d_out, amap, B, H, dk, Tq, Tk, causal, p):
    keep = (amap != 0) if p > 0 else None
    qr, kr, vr = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    with torch.enable_grad():
        o, _ = _attn(qr, kr, vr, B, H, dk, Tq, Tk, causal, keep, R.keep_scale(p) if p > 0 else 1.0)
        o.backward(d_out[:B * Tq])
    return qr.grad, kr.grad, vr.grad


def _ln(x, a, b, eps):
    mean = x.mean(-1, keepdim=True)
    return a * (x - mean) / (x.std(-1, keepdim=True) + eps) + b


def ln_fwd(x, res, a, b, eps=1e-6):
    xs = x + res
    return xs, _ln(xs, a, b, eps)


def ln_bwd(dy, xs, a, eps=1e-6, out_a=None, out_b=None):
    xr, ar = xs.detach().clone().requires_grad_(True), a.detach().clone().requires_grad_(True)
    br = torch.zeros_like(ar).requires_grad_(True)
    with torch.enable_grad():
        _ln(xr, ar, br, eps).backward(dy)
    return torch.nan_to_num(xr.grad), _into(out_a, torch.nan_to_num(ar.grad)), _into(out_b, br.grad)   # constant (padding) rows: the kernel emits 0


def embed_fwd(idx, lut, rows_pad, p, seed, sid):
    B, T = idx.shape
    vocab, E = lut.shape
    pe = torch.zeros(T, E)
    pos = torch.arange(0, T).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, E, 2).float() * -(math.log(10000.0) / E))
    pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
    pe = pe.unsqueeze(0).expand(B, T, E)
    if p > 0:
        pe = pe * _keep(B * T * E, p, seed, sid).reshape(B, T, E) * R.keep_scale(p)
    out = torch.zeros(rows_pad, 2 * E)
    out[:B * T] = torch.cat([lut[idx] * math.sqrt(E), pe], 2).reshape(B * T, 2 * E)
    return out


def embed_bwd(idx, d_out, vocab, E, out=None):
    return _into(out, torch.zeros(vocab, E).index_add_(0, idx.reshape(-1), d_out[:idx.numel(), :E]) * math.sqrt(E))


def packed_ce(logits, B, T, C, length, gt, gscale=1.0, want_grad=True):
    x = logits.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        x3 = x[:B * T].view(B, T, -1)[:, :, :C]
        loss = F.cross_entropy(torch.cat([x3[b, :int(length[b])] for b in range(B)], 0), gt)
        if want_grad:
            (loss * gscale).backward()
    return loss.detach(), (x.grad if want_grad else None)


def l2norm_fwd(x):
    n = x.norm(dim=1, keepdim=True)
    inv = torch.where(n > 0, 1.0 / n, torch.zeros_like(n))
    return x * inv, inv[:, 0]


def l2norm_bwd(dy, y, inv):
    return inv[:, None] * (dy - y * (dy * y).sum(1, keepdim=True))


def packed_feat_mse(y, B, T, length, gt, feats, gscale=1.0, want_grad=True):
    x = y.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        x3 = x[:B * T].view(B, T, -1)
        packed = torch.cat([x3[b, :int(length[b])] for b in range(B)], 0)
        loss = F.mse_loss(packed, feats[gt])
        if want_grad:
            (loss * gscale).backward()
    return loss.detach(), (x.grad if want_grad else None)


def install(monkeypatch):
    """route fudanocr_b200.model.recog_ops through the stand-ins above (pytest monkeypatch: undone after the test)"""
    from fudanocr_b200.model import recog_ops as ops
    for name, fn in globals().items():
        if callable(fn) and not name.startswith("_") and name != "install" and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(ops, "BF", torch.float32)
    monkeypatch.setattr(ops, "require_cuda", lambda t: None)
