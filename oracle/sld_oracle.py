"""ORACLE (test infrastructure only).  torch-fp32 functional restatement of the stroke-level-decomposition recogniser and
its training step, each function citing the reference lines it follows (paths relative to
/root/reference/stroke-level-decomposition):
  ResNet encoder [3,4,6,3], pools 2-4 commented out      model/transformer.py:77-164 (BasicBlock :43-73)
  Embeddings * sqrt(512) | PositionalEncoding(zeros)     :277-286, :168-186, :346-348
  Decoder (masked self-MHA, cross-MHA, FFN, 3 LayerNorms) :289-317; attention :227-241; LayerNorm :244-254
  Generator, packing of the valid positions              :266-274, :361-373
  CrossEntropyLoss, Adadelta(lr 1, rho 0.9)              train.py:32-41, :63-77
Arithmetic lives in torch (pinned torch==1.4.0 in requirement.txt; this container: 2.11).  Pinned by
oracle/make_golden_sld.py, which runs the UNMODIFIED reference module on the same synthetic weights and asserts equality
of logits, attention map, encoder features and every parameter gradient, then records tests/golden/sld_b3.pt.
Dropout: `drop` = None runs without dropout (parity mode); a dict of keep-masks reproduces the kernels' masks."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

ALPHABET = "<12345$"      # util.py:14
LAYERS = (("layer1", 3, 128, 256), ("layer2", 4, 256, 256), ("layer3", 6, 256, 512), ("layer4", 3, 512, 512))


def _bn(sd, p, x, train, stats_out=None):
    rm, rv = sd[p + ".running_mean"], sd[p + ".running_var"]
    if stats_out is not None:   # functional update of the running buffers (momentum 0.1, unbiased variance)
        rm, rv = rm.clone(), rv.clone()
        stats_out[p + ".running_mean"], stats_out[p + ".running_var"] = rm, rv
    elif train:
        rm, rv = rm.clone(), rv.clone()
    return F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], training=train, momentum=0.1, eps=1e-5)


def _conv(sd, p, x):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)


def basic_block(sd, p, x, train, stats_out=None):                      # transformer.py:57-73
    out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x), train, stats_out))
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out), train, stats_out)
    res = x
    if p + ".downsample.0.weight" in sd:
        res = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x), train, stats_out)
    return F.relu(out + res)


def encoder(sd, image, train=True, stats_out=None, taps=None):         # transformer.py:126-164
    """taps (optional dict): receives (input, output) of every stage - 'stem' (conv1 + bn1 + relu + pool), 'conv2', each
    BasicBlock 'layerL.i' and each transition conv 'layerL_conv' - for teacher-forced per-stage tests"""
    e = "encoder."

    def tap(name, xin, xout):
        if taps is not None:
            taps[name] = (xin.detach(), xout.detach())
        return xout
    x = tap("stem", image, F.max_pool2d(F.relu(_bn(sd, e + "bn1", _conv(sd, e + "conv1", image), train, stats_out)), (2, 2), (2, 2)))
    x = tap("conv2", x, F.relu(_bn(sd, e + "bn2", _conv(sd, e + "conv2", x), train, stats_out)))
    for name, n, _, _ in LAYERS:
        for i in range(n):
            x = tap(f"{name}.{i}", x, basic_block(sd, f"{e}{name}.{i}", x, train, stats_out))
        tail = name + ("_conv2" if name == "layer4" else "_conv")
        bn = name + ("_conv2_bn" if name == "layer4" else "_bn")
        x = tap(tail, x, F.relu(_bn(sd, e + bn, _conv(sd, e + tail, x), train, stats_out)))
    return x


def layer_norm(x, a, b, eps=1e-6):                                     # transformer.py:251-254
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a * (x - mean) / (std + eps) + b


def positional_encoding(T, d_model, device):                           # transformer.py:173-180
    pe = torch.zeros(T, d_model, device=device)
    position = torch.arange(0, T, device=device).unsqueeze(1).float()
    div_term = torch.exp(torch.arange(0, d_model, 2, device=device).float() * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def mha(sd, p, query, key, value, mask, h=4, keep=None, keep_scale=1.0):   # transformer.py:205-223, :227-241
    B = query.shape[0]
    d_k = query.shape[-1] // h
    lin = lambda i, x: F.linear(x, sd[f"{p}.linears.{i}.weight"], sd[f"{p}.linears.{i}.bias"])
    q, k, v = (lin(i, x).view(B, -1, h, d_k).transpose(1, 2) for i, x in enumerate((query, key, value)))
    scores = q @ k.transpose(-2, -1) / math.sqrt(d_k)
    if mask is not None:
        scores = scores.masked_fill(mask == 0, float("-inf"))
    p_attn = F.softmax(scores, dim=-1)
    if keep is not None:
        p_attn = p_attn * keep * keep_scale
    x = (p_attn @ v).transpose(1, 2).contiguous().view(B, -1, h * d_k)
    return lin(3, x), p_attn


def _ln_params(sd, p):
    """(a, b) of a LayerNorm: named a / b in stroke-level-decomposition (:247-248), a_2 / b_2 in image-ids-CTR (:253-254)"""
    return (sd[p + ".a"], sd[p + ".b"]) if p + ".a" in sd else (sd[p + ".a_2"], sd[p + ".b_2"])


def decoder(sd, text, conv_feature, drop: Optional[Dict] = None):      # transformer.py:303-317
    d = "decoder."
    T = text.shape[1]
    mask = torch.tril(torch.ones(1, 1, T, T, device=text.device)) != 0     # subsequent_mask :226-229
    ks = drop["scale"] if drop else 1.0
    sa, _ = mha(sd, d + "mask_multihead", text, text, text, mask, keep=drop and drop["self"], keep_scale=ks)
    result = layer_norm(text + sa, *_ln_params(sd, d + "mul_layernorm1"))
    b, c, hh, ww = conv_feature.shape
    feat = conv_feature.view(b, c, hh * ww).permute(0, 2, 1).contiguous()
    ca, attention_map = mha(sd, d + "multihead", result, feat, feat, None, keep=drop and drop["cross"], keep_scale=ks)
    result = layer_norm(result + ca, *_ln_params(sd, d + "mul_layernorm2"))
    hdn = F.relu(F.linear(result, sd[d + "pff.w_1.weight"], sd[d + "pff.w_1.bias"]))   # :262-263
    if drop:
        hdn = hdn * drop["ffn"] * ks
    ff = F.linear(hdn, sd[d + "pff.w_2.weight"], sd[d + "pff.w_2.bias"])
    result = layer_norm(result + ff, *_ln_params(sd, d + "mul_layernorm3"))
    return result, attention_map


def forward(sd, image, text_input, train=True, conv_feature=None, drop: Optional[Dict] = None, stats_out=None):
    """-> (logits (B, T, 7), attention map (B, 4, T, tokens), conv feature (B, 1024, H/2, W/2))   transformer.py:339-377"""
    if conv_feature is None:
        conv_feature = encoder(sd, image, train, stats_out)
    emb = F.embedding(text_input, sd["embedding_word.lut.weight"]) * math.sqrt(512)
    pe = positional_encoding(text_input.shape[1], 512, emb.device).unsqueeze(0).expand(emb.shape[0], -1, -1)
    if drop:
        pe = pe * drop["pe"] * drop["scale"]
    x = torch.cat([emb, pe], 2)
    x, amap = decoder(sd, x, conv_feature, drop)
    logits = F.linear(x, sd["generator_word.proj.weight"], sd["generator_word.proj.bias"])
    return logits, amap, conv_feature


def pack(logits, length):                                              # transformer.py:361-373
    return torch.cat([logits[b, :int(n)] for b, n in enumerate(length)], 0)


def loss_fn(sd, image, length, text_input, text_gt, drop=None, stats_out=None):   # train.py:68-71
    logits, amap, conv = forward(sd, image, text_input, True, None, drop, stats_out)
    return F.cross_entropy(pack(logits, length), text_gt), logits, amap, conv


def adadelta_update(p, g, sq, acc, lr=1.0, rho=0.9, eps=1e-6, wd=0.0):   # torch.optim.Adadelta (train.py:32-36)
    if wd:
        g = g + wd * p
    sq = rho * sq + (1 - rho) * g * g
    delta = (acc + eps).sqrt() / (sq + eps).sqrt() * g
    acc = rho * acc + (1 - rho) * delta * delta
    return p - lr * delta, sq, acc


def converter_stroke(stroke_strings):                                   # util.py:90-116 with the stroke alphabet
    """['12$', '3$', ...] (each already ending in '$') -> (length, text_input, text_gt) exactly as util.converter builds them"""
    a2n = {c: i for i, c in enumerate(ALPHABET)}
    length = torch.tensor([len(s) for s in stroke_strings], dtype=torch.long)
    T = int(length.max())
    text_input = torch.zeros(len(stroke_strings), T, dtype=torch.long)
    for i, s in enumerate(stroke_strings):
        for j in range(len(s) - 1):
            text_input[i, j + 1] = a2n[s[j]]
    text_gt = torch.tensor([a2n[c] for s in stroke_strings for c in s], dtype=torch.long)
    return length, text_input, text_gt


def synth_batch(B: int, seed: int = 1234, size: int = 32):
    """images in [-1, 1] (lmdbReader: sub_(0.5).div_(0.5)) with smooth structure, stroke strings of length 2..8"""
    import numpy as np
    rs = np.random.RandomState(seed + 5)
    yy, xx = np.meshgrid(np.linspace(0, 1, size), np.linspace(0, 1, size), indexing="ij")
    img = np.zeros((B, 3, size, size), np.float64)
    for b in range(B):
        for c in range(3):
            f = np.zeros_like(yy)
            for _ in range(5):
                f += rs.uniform(0.3, 1.0) * np.cos(2 * np.pi * (rs.uniform(0.5, 4) * xx + rs.uniform(0.5, 4) * yy) + rs.uniform(0, 6.28))
            for _ in range(3):
                x0, wd = rs.uniform(0.1, 0.9), rs.uniform(0.02, 0.06)
                f += rs.uniform(1, 2) * np.exp(-((xx - x0) / wd) ** 2)
            f = (f - f.min()) / (f.max() - f.min() + 1e-9)
            img[b, c] = 2 * f - 1 + 0.02 * rs.standard_normal(f.shape)
    strings = ["".join(rs.choice(list("12345"), size=int(rs.randint(1, 8)))) + "$" for _ in range(B)]
    return torch.from_numpy(np.clip(img, -1, 1).astype(np.float32)), strings
