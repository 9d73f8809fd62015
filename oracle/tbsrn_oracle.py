"""ORACLE — test infrastructure only (never imported by the product path).

A CPU/fp32 functional restatement, in plain torch ops, of the reference's TBSRN training hot path
(FudanVI/FudanOCR, scene-text-telescope).  Each function cites the reference file:line it
follows.  It is pinned against the *real* reference modules by tests/golden/*.pt, which are
produced by oracle/make_golden.py (that script imports /root/reference; this file does not).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---------------------------------------------------------------------------------------------
# small pieces
# ---------------------------------------------------------------------------------------------
def mish(x: Tensor) -> Tensor:
    """model/tbsrn.py:277-285  x * tanh(softplus(x))."""
    return x * torch.tanh(F.softplus(x))


def layer_norm_std(x: Tensor, a: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """model/tbsrn.py:23-36: unbiased std, eps added to the std (NOT nn.LayerNorm semantics)."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a * (x - mean) / (std + eps) + b


def positionalencoding2d(d_model: int, height: int, width: int) -> Tensor:
    """model/tbsrn.py:39-61."""
    pe = torch.zeros(d_model, height, width)
    d = d_model // 2
    div_term = torch.exp(torch.arange(0.0, d, 2) * -(math.log(10000.0) / d))
    pos_w = torch.arange(0.0, width).unsqueeze(1)
    pos_h = torch.arange(0.0, height).unsqueeze(1)
    pe[0:d:2] = torch.sin(pos_w * div_term).transpose(0, 1).unsqueeze(1).repeat(1, height, 1)
    pe[1:d:2] = torch.cos(pos_w * div_term).transpose(0, 1).unsqueeze(1).repeat(1, height, 1)
    pe[d::2] = torch.sin(pos_h * div_term).transpose(0, 1).unsqueeze(2).repeat(1, 1, width)
    pe[d + 1::2] = torch.cos(pos_h * div_term).transpose(0, 1).unsqueeze(2).repeat(1, 1, width)
    return pe


def batch_norm_train(x: Tensor, sd: Dict[str, Tensor], pre: str, new_stats: Optional[dict],
                     training: bool, eps: float = 1e-5, momentum: float = 0.1) -> Tensor:
    """nn.BatchNorm2d / BatchNorm1d semantics (torch, third-party; parity unpinned by the reference):
    train: biased batch variance for normalisation, unbiased for the running update."""
    w, b = sd[pre + ".weight"], sd[pre + ".bias"]
    if not training:
        return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"], w, b, False, momentum, eps)
    rm, rv = sd[pre + ".running_mean"].clone(), sd[pre + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, w, b, True, momentum, eps)
    if new_stats is not None:
        new_stats[pre + ".running_mean"] = rm
        new_stats[pre + ".running_var"] = rv
        new_stats[pre + ".num_batches_tracked"] = sd[pre + ".num_batches_tracked"] + 1
    return y


# ---------------------------------------------------------------------------------------------
# FeatureEnhancer  (model/tbsrn.py:63-163)
# ---------------------------------------------------------------------------------------------
def multi_head_attention(sd, pre, x, attn_keep: Optional[Tensor], p_drop: float, h: int = 4,
                         keep_scale: Optional[float] = None):
    """model/tbsrn.py:95-150: q,k,v,out = 4 x Linear(128,128); softmax(QK^T/sqrt(d_k)); dropout on P.
    keep_scale: the 1/(1-p) rescale of kept probabilities when the caller's mask was drawn at a quantised rate
    (the CUDA attention draws at 3277/32768 for p = 0.1); default 1/(1-p_drop) as nn.Dropout."""
    B, S, D = x.shape
    dk = D // h
    q, k, v = [F.linear(x, sd[f"{pre}.linears.{i}.weight"], sd[f"{pre}.linears.{i}.bias"])
               .view(B, S, h, dk).transpose(1, 2) for i in range(3)]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    p = F.softmax(scores, dim=-1)
    if attn_keep is not None:
        p = p * attn_keep.to(p.dtype) * (keep_scale if keep_scale is not None else 1.0 / (1.0 - p_drop))
    o = torch.matmul(p, v).transpose(1, 2).contiguous().view(B, S, D)
    return F.linear(o, sd[f"{pre}.linears.3.weight"], sd[f"{pre}.linears.3.bias"])


def feature_enhancer(sd, pre, conv_feature: Tensor, masks: Optional[dict], p_drop: float = 0.1) -> Tensor:
    """model/tbsrn.py:76-92.  conv_feature: (B,64,1024) -> (B,64,1024)."""
    B = conv_feature.shape[0]
    pe = positionalencoding2d(64, 16, 64).to(conv_feature).view(1, 64, 1024).repeat(B, 1, 1)
    x = torch.cat([conv_feature, pe], 1).permute(0, 2, 1).contiguous()  # (B,1024,128)
    attn_keep = masks.get(pre + ".attn") if masks else None
    ffn_keep = masks.get(pre + ".ffn") if masks else None
    attn_scale = masks.get(pre + ".attn_scale") if masks else None
    y = layer_norm_std(x + multi_head_attention(sd, pre + ".multihead", x, attn_keep, p_drop, keep_scale=attn_scale),
                       sd[pre + ".mul_layernorm1.a_2"], sd[pre + ".mul_layernorm1.b_2"])
    hdn = F.relu(F.linear(y, sd[pre + ".pff.w_1.weight"], sd[pre + ".pff.w_1.bias"]))
    if ffn_keep is not None:
        hdn = hdn * ffn_keep.to(hdn.dtype) / (1.0 - p_drop)
    z = layer_norm_std(y + F.linear(hdn, sd[pre + ".pff.w_2.weight"], sd[pre + ".pff.w_2.bias"]),
                       sd[pre + ".mul_layernorm3.a_2"], sd[pre + ".mul_layernorm3.b_2"])
    out = F.linear(z, sd[pre + ".linear.weight"], sd[pre + ".linear.bias"])
    return out.permute(0, 2, 1).contiguous()


def srb(sd, pre, x, new_stats, training, masks, taps=None):
    """RecurrentResidualBlock.forward, model/tbsrn.py:246-257 (gru1/gru2 exist but are never called)."""
    c1 = F.conv2d(x, sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], padding=1)
    a1 = mish(batch_norm_train(c1, sd, pre + ".bn1", new_stats, training))
    c2 = F.conv2d(a1, sd[pre + ".conv2.weight"], sd[pre + ".conv2.bias"], padding=1)
    r = batch_norm_train(c2, sd, pre + ".bn2", new_stats, training)
    size = r.shape
    r = feature_enhancer(sd, pre + ".feature_enhancer", r.view(size[0], size[1], -1), masks)
    out = x + r.reshape(size)
    if taps is not None:
        taps.update({pre + ".c1": c1.detach(), pre + ".a1": a1.detach(), pre + ".c2": c2.detach(),
                     pre + ".out": out.detach()})
    return out


# ---------------------------------------------------------------------------------------------
# STN head + TPS  (model/stn_head.py:25-99, model/tps_spatial_transformer.py:54-112)
# ---------------------------------------------------------------------------------------------
def stn_head(sd, x, new_stats, training, pre="stn_head", taps=None):
    pools = {0: (2, 2), 2: (2, 2), 4: (2, 2), 6: (2, 2), 8: ((1, 2), (1, 2))}
    for i in range(0, 11, 2):  # conv3x3_block indices 0,2,4,6,8,10 of the Sequential
        p = f"{pre}.stn_convnet.{i}"
        x = F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1)
        if taps is not None:
            taps[f"stn.ypre{i // 2}"] = x.detach()
        x = F.relu(batch_norm_train(x, sd, p + ".1", new_stats, training))
        if taps is not None:
            taps[f"stn.yact{i // 2}"] = x.detach()
        if i in pools:
            x = F.max_pool2d(x, kernel_size=pools[i][0], stride=pools[i][1])
    x = x.reshape(x.shape[0], -1)
    f = F.linear(x, sd[pre + ".stn_fc1.0.weight"], sd[pre + ".stn_fc1.0.bias"])
    if taps is not None:
        taps["stn.f1pre"] = f.detach()
    f = F.relu(batch_norm_train(f, sd, pre + ".stn_fc1.1", new_stats, training))
    if taps is not None:
        taps["stn.f1"] = f.detach()
    c = F.linear(0.1 * f, sd[pre + ".stn_fc2.weight"], sd[pre + ".stn_fc2.bias"])  # stn_head.py:93
    return c.view(-1, 20, 2)


def tps_transform(sd, x, ctrl, pre="tps"):
    """tps_spatial_transformer.py:97-112; grid_sample bilinear/zeros/align_corners=False (torch>=1.3 default)."""
    B = ctrl.shape[0]
    Y = torch.cat([ctrl, sd[pre + ".padding_matrix"].expand(B, 3, 2)], 1)
    mapping = torch.matmul(sd[pre + ".inverse_kernel"], Y)
    src = torch.matmul(sd[pre + ".target_coordinate_repr"], mapping)
    H, W = x.shape[-2:]
    grid = torch.clamp(src.view(-1, H, W, 2), 0, 1)
    grid = 2.0 * grid - 1.0
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


# ---------------------------------------------------------------------------------------------
# whole network and the training step
# ---------------------------------------------------------------------------------------------
def tbsrn_forward(sd: Dict[str, Tensor], x: Tensor, training: bool = True, stn: bool = True,
                  srb_nums: int = 5, masks: Optional[dict] = None, new_stats: Optional[dict] = None,
                  taps: Optional[dict] = None) -> Tensor:
    """TBSRN.forward, model/tbsrn.py:214-226."""
    if stn and training:
        ctrl = stn_head(sd, x, new_stats, training, taps=taps)
        x = tps_transform(sd, x, ctrl)
        if taps is not None:
            taps["ctrl"], taps["x_tps"] = ctrl.detach(), x.detach()
    b1 = F.conv2d(x, sd["block1.0.weight"], sd["block1.0.bias"], padding=4)
    b1 = F.prelu(b1, sd["block1.1.weight"])
    cur = b1
    for i in range(srb_nums):
        cur = srb(sd, f"block{i + 2}", cur, new_stats, training, masks, taps)
    k = srb_nums + 2
    cur = F.conv2d(cur, sd[f"block{k}.0.weight"], sd[f"block{k}.0.bias"], padding=1)
    cur = batch_norm_train(cur, sd, f"block{k}.1", new_stats, training)
    k = srb_nums + 3
    u = F.conv2d(b1 + cur, sd[f"block{k}.0.conv.weight"], sd[f"block{k}.0.conv.bias"], padding=1)
    u = mish(F.pixel_shuffle(u, 2))
    out = F.conv2d(u, sd[f"block{k}.1.weight"], sd[f"block{k}.1.bias"], padding=4)
    if taps is not None:
        taps.update({"b1": b1.detach(), "s7": (b1 + cur).detach(), "u": u.detach(), "opre": out.detach()})
    return torch.tanh(out)


def is_buffer(key: str) -> bool:
    """state-dict entries that are registered buffers, not nn.Parameters (BN running stats, TPS matrices)."""
    return "running_" in key or key.endswith("num_batches_tracked") or key.startswith("tps.")


def clip_grad_norm(grads, max_norm: float = 0.25, eps: float = 1e-6):
    """torch.nn.utils.clip_grad_norm_ (interfaces/super_resolution.py:83): global L2, coef clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + eps), max=1.0)
    return {k: g * coef for k, g in grads.items()}, total


def adam_step(params, grads, state, lr=1e-4, betas=(0.5, 0.999), eps=1e-8):
    """torch.optim.Adam (interfaces/base.py:194-198: lr 1e-4, betas (0.5, 0.999), no weight decay)."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    b1, b2 = betas
    out = {}
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            out[k] = p
            continue
        m = state.setdefault("m." + k, torch.zeros_like(p))
        v = state.setdefault("v." + k, torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        out[k] = p - (lr / (1 - b1 ** t)) * m / denom
    return out


def train_step(sd: Dict[str, Tensor], lr_img: Tensor, hr_img: Tensor, opt_state: dict,
               masks: Optional[dict] = None, stn: bool = True, srb_nums: int = 5, taps: Optional[dict] = None):
    """The step body of TextSR.train (interfaces/super_resolution.py:60-84) with the MSE-only
    image_crit (loss/text_focus_loss.py:84-103, text_focus off): forward, mse, x100, backward,
    clip_grad_norm_(0.25), Adam.  Returns (new_sd, info)."""
    float_keys = [k for k, v in sd.items() if v.is_floating_point() and not is_buffer(k)]
    leaf = {k: (sd[k].detach().clone().requires_grad_(True) if k in float_keys else sd[k]) for k in sd}
    new_stats: dict = {}
    sr = tbsrn_forward(leaf, lr_img, training=True, stn=stn, srb_nums=srb_nums, masks=masks, new_stats=new_stats,
                       taps=taps)
    mse = F.mse_loss(sr, hr_img)
    (mse * 100).backward()
    grads = {k: leaf[k].grad for k in float_keys if leaf[k].grad is not None}
    clipped, gnorm = clip_grad_norm(grads)
    params = {k: sd[k] for k in float_keys}
    new_params = adam_step(params, clipped, opt_state)
    new_sd = dict(sd)
    new_sd.update(new_params)
    new_sd.update(new_stats)
    return new_sd, {"sr": sr.detach(), "mse": mse.detach(), "grad_norm": gnorm, "grads": grads}


# ---------------------------------------------------------------------------------------------
# architecture-determined buffers
# ---------------------------------------------------------------------------------------------
def tps_buffers(height: int = 16, width: int = 64, n: int = 20, margins=(0.05, 0.05)) -> Dict[str, Tensor]:
    """TPSSpatialTransformer.__init__, model/tps_spatial_transformer.py:54-95 (registered buffers)."""
    import numpy as np

    def partial_repr(inp, ctrl):  # :22-34  phi(r) = 0.5 r^2 log r^2, NaN (0*log 0) -> 0
        d = inp.view(-1, 1, 2) - ctrl.view(1, -1, 2)
        dist = (d * d).sum(-1)
        r = 0.5 * dist * torch.log(dist)
        return torch.where(r != r, torch.zeros_like(r), r)

    k = n // 2
    xs = np.linspace(margins[0], 1.0 - margins[0], k)
    top = np.stack([xs, np.ones(k) * margins[1]], axis=1)
    bot = np.stack([xs, np.ones(k) * (1.0 - margins[1])], axis=1)
    tcp = torch.Tensor(np.concatenate([top, bot], axis=0))
    fk = torch.zeros(n + 3, n + 3)
    fk[:n, :n] = partial_repr(tcp, tcp)
    fk[:n, -3] = 1
    fk[-3, :n] = 1
    fk[:n, -2:] = tcp
    fk[-2:, :n] = tcp.t()
    inv = torch.inverse(fk)
    ys, xs_ = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32),
                             indexing="ij")
    coord = torch.stack([xs_.reshape(-1) / (width - 1), ys.reshape(-1) / (height - 1)], dim=1)
    rep = torch.cat([partial_repr(coord, tcp), torch.ones(height * width, 1), coord], dim=1)
    return {"tps.inverse_kernel": inv, "tps.padding_matrix": torch.zeros(3, 2),
            "tps.target_coordinate_repr": rep, "tps.target_control_points": tcp}
