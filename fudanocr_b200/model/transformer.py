"""Stroke-level-decomposition recogniser on the focr engine - drop-in for ``model.transformer.Transformer`` of
stroke-level-decomposition (model/transformer.py:320-377): ResNet encoder [3,4,6,3] without the later pooling steps
(:77-164, 40 3x3 convolutions with train-mode BatchNorm), one Transformer decoder layer with h = 4, d_model = 1024 (:289-317)
and a linear generator over the 7-symbol stroke alphabet ``'<12345$'`` (util.py:14).  SURVEY.md §8 row A21.

Same constructor (``Transformer(mode)``), same ``state_dict`` keys (314 entries incl. the dead ``compress_attention_linear``
layers and the ``pe.pe`` buffer), same ``forward(image, text_length, text_input, conv_feature=None, test=False)`` contract and
return dict.  The submodules are parameter containers only; the arithmetic is a chain of ``torch.autograd.Function`` nodes
whose forward / backward bodies are C-ABI kernel calls (model/recog_ops.py): tcgen05 implicit-GEMM convolutions and linears,
im2col GEMM weight gradients, fused BatchNorm(+ReLU), the decoder attention / LayerNorm / embedding / cross-entropy kernels
of csrc/recog_ops.cu.  PyTorch contributes the autograd tape, memory and a handful of views - no torch kernel touches an
activation on the fused path (``SLDTrainer``).  CUDA only; there is no CPU fallback."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import recog_ops as ops

ALPHABET_STROKE = "<12345$"  # stroke-level-decomposition/util.py:14


def get_alphabet(mode: str, alphabet=None):
    """util.get_alphabet (stroke-level-decomposition/util.py:119-123): the 7 stroke symbols, or - mode 'character' - the caller's
    character alphabet ('<' + the 3755 GB2312 level-1 characters + '$' in the reference, util.py:21: data, not code, so it is
    passed in rather than restated here)"""
    if mode == "stroke":
        return ALPHABET_STROKE
    if mode == "character":
        if alphabet is None:
            raise ValueError("focr Transformer(mode='character') needs alphabet=<the character alphabet of util.py:21>")
        return alphabet
    raise ValueError(f"unknown mode {mode!r} (config.py:5: 'character' / 'stroke')")


def _bf(t):
    return t if t.dtype == ops.BF else t.to(ops.BF)


# ---- autograd nodes: bodies are kernel calls ----------------------------------------------------------------------------
def _map_tiles(h: int, w: int) -> bool:
    """can the implicit-GEMM convs (csrc/tc_gemm.cu: tc_gemm_launch) cut an (h, w) map into 128-pixel boxes?"""
    if w in (16, 32, 64, 128):
        rows = min(128 // w, h)
        return h % rows == 0
    return any(w % c == 0 and h % (128 // c) == 0 for c in (128, 64, 32, 16))


# Gradient sinks.  Autograd adds every parameter gradient a Function returns into `p.grad` with one at::add kernel per parameter
# (193 of them per step here).  The fused trainers (trainer_sld.py) own a flat gradient buffer that `p.grad` already views; inside
# their step they publish {parameter data pointer: gradient view} and the backward bodies below hand those views to the kernels as
# OUTPUT buffers and return None for the parameter - the gradient lands in place, nothing is accumulated.  Each parameter of these
# models is used once per forward, so overwrite == accumulate into the zeroed buffer.  Outside a trainer step (the reference's own
# loop: loss.backward(); optimizer.step()) the table is None and autograd's semantics - accumulation included - are untouched.
_GRAD_SINKS = None


class grad_sinks:
    """context manager: parameter gradients of every backward run inside are written straight into `table[param.data_ptr()]`"""

    def __init__(self, table):
        self.table = table

    def __enter__(self):
        global _GRAD_SINKS
        self.prev, _GRAD_SINKS = _GRAD_SINKS, self.table
        return self

    def __exit__(self, *exc):
        global _GRAD_SINKS
        _GRAD_SINKS = self.prev
        return False


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _sink(ptr):
    return None if (_GRAD_SINKS is None or not ptr) else _GRAD_SINKS.get(ptr)


def _ret(sink, grad):
    return None if sink is not None else grad


class _ConvFirst(Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x)
        ctx.wshape = tuple(w.shape)
        ctx.ptrs = (_ptr(w), _ptr(b))
        return ops.conv_first_fwd(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        sw, sb = _sink(ctx.ptrs[0]), _sink(ctx.ptrs[1])
        dw, db = ops.conv_first_wgrad(_bf(dy).contiguous(), x, ctx.wshape, sw, sb)
        return None, _ret(sw, dw), _ret(sb, db)


class _Conv(Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.ptrs = (_ptr(w), _ptr(b))
        return ops.conv_fwd(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _bf(dy).contiguous()
        dx = ops.conv_dgrad(dy, w) if ctx.needs_input_grad[0] else None
        sw, sb = _sink(ctx.ptrs[0]), _sink(ctx.ptrs[1])
        dw, db = ops.conv_wgrad(dy, x, tuple(w.shape), sw, sb)
        return dx, _ret(sw, dw), _ret(sb, db)


class _BNTrain(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, rm, rv, nbt, act):
        y, stats = ops.bn_train_fwd(x, gamma, beta, rm, rv, nbt, act)
        ctx.save_for_backward(x, stats)
        ctx.act = act
        ctx.ptrs = (_ptr(gamma), _ptr(beta))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        sg, sb = _sink(ctx.ptrs[0]), _sink(ctx.ptrs[1])
        dx, dg, db = ops.bn_bwd(_bf(dy).contiguous(), x, stats, ctx.act, sg, sb)
        return dx, _ret(sg, dg), _ret(sb, db), None, None, None, None


class _AddRelu(Function):
    @staticmethod
    def forward(ctx, a, b):
        y = ops.add_relu(a, b)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        g = ops.relu_bwd(_bf(dy).contiguous(), y)
        return g, g


class _MaxPool(Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.maxpool_fwd(x)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        return ops.maxpool_bwd(x, y, _bf(dy).contiguous())


class _Linear(Function):
    @staticmethod
    def forward(ctx, x, w, b, relu, fp32_out):
        y = ops.linear_fwd(x, w, b, relu, fp32_out)
        ctx.relu = relu
        ctx.ptrs = (_ptr(w), _ptr(b))
        ctx.save_for_backward(x, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _bf(dy).contiguous()
        if ctx.relu:
            dy = ops.relu_bwd(dy, y)
        dx = ops.linear_dgrad(dy, w) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:   # frozen operands (the CLIP text features of image-ids-CTR) skip this
            sw = _sink(ctx.ptrs[0]) if ctx.needs_input_grad[1] else None
            sb = _sink(ctx.ptrs[1]) if ctx.needs_input_grad[2] else None
            dw, db = ops.linear_wgrad(dy, x, sw, sb)
            dw, db = _ret(sw, dw), _ret(sb, db)
        return dx, dw, (db if ctx.needs_input_grad[2] else None), None, None


class _MHA(Function):
    @staticmethod
    def forward(ctx, q, k, v, B, H, dk, Tq, Tk, causal, p, seed, sid):
        out, amap = ops.mha_fwd(q, k, v, B, H, dk, Tq, Tk, causal, p, seed, sid)
        ctx.save_for_backward(q, k, v, amap)
        ctx.cfg = (B, H, dk, Tq, Tk, causal, p)
        ctx.mark_non_differentiable(amap)
        return out, amap

    @staticmethod
    def backward(ctx, d_out, _d_map):
        q, k, v, amap = ctx.saved_tensors
        B, H, dk, Tq, Tk, causal, p = ctx.cfg
        dq, dk_, dv = ops.mha_bwd(q, k, v, _bf(d_out).contiguous(), amap, B, H, dk, Tq, Tk, causal, p)
        return (dq, dk_, dv) + (None,) * 9


class _LN(Function):
    @staticmethod
    def forward(ctx, x, res, a, b):
        xs, y = ops.ln_fwd(x, res, a, b)
        ctx.save_for_backward(xs, a)
        ctx.ptrs = (_ptr(a), _ptr(b))
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, a = ctx.saved_tensors
        sa, sb = _sink(ctx.ptrs[0]), _sink(ctx.ptrs[1])
        dx, da, db = ops.ln_bwd(_bf(dy).contiguous(), xs, a, out_a=sa, out_b=sb)
        return dx, dx, _ret(sa, da), _ret(sb, db)


class _Embed(Function):
    @staticmethod
    def forward(ctx, idx, lut, rows_pad, p, seed, sid):
        ctx.save_for_backward(idx)
        ctx.shape = tuple(lut.shape)
        ctx.ptrs = (_ptr(lut),)
        return ops.embed_fwd(idx, lut, rows_pad, p, seed, sid)

    @staticmethod
    def backward(ctx, d_out):
        (idx,) = ctx.saved_tensors
        sl = _sink(ctx.ptrs[0])
        return None, _ret(sl, ops.embed_bwd(idx, _bf(d_out).contiguous(), *ctx.shape, out=sl)), None, None, None, None


class _Dropout(Function):
    @staticmethod
    def forward(ctx, x, p, seed, sid):
        ctx.cfg = (p, seed, sid)
        return ops.dropout(x, p, seed, sid)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(_bf(dy).contiguous(), *ctx.cfg), None, None, None


class _PackedCE(Function):
    """mean cross entropy over the positions t < length[b] - the packing loop of Transformer.forward (transformer.py:361-373)
    followed by nn.CrossEntropyLoss (train.py:41,71) - value and logits gradient from one kernel"""

    @staticmethod
    def forward(ctx, logits, B, T, C, length, gt):
        loss, d = ops.packed_ce(logits, B, T, C, length, gt, 1.0, True)
        ctx.save_for_backward(d)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g.to(d.dtype), None, None, None, None, None


class _L2Norm(Function):
    """rows of an fp32 matrix scaled to unit length (image-ids-CTR/train.py:76)"""

    @staticmethod
    def forward(ctx, x):
        y, inv = ops.l2norm_fwd(x)
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        return ops.l2norm_bwd(_bf(dy).contiguous(), y, inv)


class _FeatMSE(Function):
    """nn.MSELoss(pred_n, text_features[text_gt]) over the packed valid rows (image-ids-CTR/train.py:66-69,79)"""

    @staticmethod
    def forward(ctx, y, B, T, length, gt, feats):
        loss, d = ops.packed_feat_mse(y, B, T, length, gt, feats, 1.0, True)
        ctx.save_for_backward(d)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g.to(d.dtype), None, None, None, None, None


# ---- parameter containers with the reference's attribute names -----------------------------------------------------------
class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is a parameter container of the focr Transformer; call the model itself")


class _BasicBlock(_Container):  # transformer.py:43-73
    def __init__(self, inplanes, planes, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=1, padding=1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample


class _ResNet(_Container):  # transformer.py:77-164
    def __init__(self, num_in, layers):
        super().__init__()
        self.conv1 = nn.Conv2d(num_in, 64, 3, 1, 1)
        self.bn1 = nn.BatchNorm2d(64)
        self.conv2 = nn.Conv2d(64, 128, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(128)
        self.layer1 = self._make_layer(128, 256, layers[0])
        self.layer1_conv = nn.Conv2d(256, 256, 3, 1, 1)
        self.layer1_bn = nn.BatchNorm2d(256)
        self.layer2 = self._make_layer(256, 256, layers[1])
        self.layer2_conv = nn.Conv2d(256, 256, 3, 1, 1)
        self.layer2_bn = nn.BatchNorm2d(256)
        self.layer3 = self._make_layer(256, 512, layers[2])
        self.layer3_conv = nn.Conv2d(512, 512, 3, 1, 1)
        self.layer3_bn = nn.BatchNorm2d(512)
        self.layer4 = self._make_layer(512, 512, layers[3])
        self.layer4_conv2 = nn.Conv2d(512, 1024, 3, 1, 1)
        self.layer4_conv2_bn = nn.BatchNorm2d(1024)

    @staticmethod
    def _make_layer(inplanes, planes, blocks):
        ds = nn.Sequential(nn.Conv2d(inplanes, planes, 3, 1, 1), nn.BatchNorm2d(planes)) if inplanes != planes else None
        return nn.Sequential(_BasicBlock(inplanes, planes, ds), *[_BasicBlock(planes, planes, None) for _ in range(1, blocks)])


class _Embeddings(_Container):
    def __init__(self, d_model, vocab):
        super().__init__()
        self.lut = nn.Embedding(vocab, d_model)
        self.d_model = d_model


class _PositionalEncoding(_Container):  # transformer.py:168-186 (buffer kept for state_dict parity; the kernel evaluates it)
    def __init__(self, d_model, max_len=7000):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len).unsqueeze(1).float()
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class _MultiHeadedAttention(_Container):  # transformer.py:189-223
    def __init__(self, h, d_model):
        super().__init__()
        self.h, self.d_k = h, d_model // h
        self.linears = nn.ModuleList([nn.Linear(d_model, d_model) for _ in range(4)])
        self.compress_attention_linear = nn.Linear(h, 1)  # constructed, never used by the reference forward


class _LayerNorm(_Container):  # transformer.py:244-254 (parameters a / b; image-ids-CTR names them a_2 / b_2)
    def __init__(self, features, eps=1e-6, names=("a", "b")):
        super().__init__()
        self._names = names
        setattr(self, names[0], nn.Parameter(torch.ones(features)))
        setattr(self, names[1], nn.Parameter(torch.zeros(features)))
        self.eps = eps

    @property
    def scale(self):
        return getattr(self, self._names[0])

    @property
    def shift(self):
        return getattr(self, self._names[1])


class _PositionwiseFeedForward(_Container):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_ff)
        self.w_2 = nn.Linear(d_ff, d_model)


class _Generator(_Container):
    def __init__(self, d_model, vocab):
        super().__init__()
        self.proj = nn.Linear(d_model, vocab)


class _Decoder(_Container):  # transformer.py:289-317
    def __init__(self, ln_names=("a", "b")):
        super().__init__()
        self.mask_multihead = _MultiHeadedAttention(4, 1024)
        self.mul_layernorm1 = _LayerNorm(1024, names=ln_names)
        self.multihead = _MultiHeadedAttention(4, 1024)
        self.mul_layernorm2 = _LayerNorm(1024, names=ln_names)
        self.pff = _PositionwiseFeedForward(1024, 2048)
        self.mul_layernorm3 = _LayerNorm(1024, names=ln_names)


def _pad128(n: int) -> int:
    return (n + 127) // 128 * 128


class Transformer(nn.Module):
    """drop-in for stroke-level-decomposition/model/transformer.py:320-377"""

    DROPOUT = 0.1  # transformer.py:292,295,297,326 and PositionwiseFeedForward default

    def __init__(self, mode: str = "stroke", alphabet=None):
        super().__init__()
        self.mode = mode
        self.word_n_class = len(get_alphabet(mode, alphabet))
        self.embedding_word = _Embeddings(512, self.word_n_class)
        self.pe = _PositionalEncoding(512)
        self.encoder = _ResNet(3, [3, 4, 6, 3])
        self.decoder = _Decoder()
        self.generator_word = _Generator(1024, self.word_n_class)
        self.attribute = None
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.dropout_p = self.DROPOUT
        # dropout stream: follows torch.manual_seed (reproducible runs); the trainers mix the rank in (reseed_dropout)
        self._seed = (0x5EED ^ int(torch.initial_seed())) & 0x7FFFFFFF

    # -- encoder ------------------------------------------------------------------------------------------------------
    def _bn(self, x, bn: nn.BatchNorm2d, act: int):
        if self.training:
            return _BNTrain.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, act)
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("focr Transformer: backward through eval-mode BatchNorm is not built (the reference evaluates under "
                               "torch.no_grad(), stroke-level-decomposition/train.py:80); call model.train() or wrap in no_grad")
        return ops.bn_eval_fwd(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, act)

    def _block(self, x, blk: _BasicBlock):
        out = self._bn(_Conv.apply(x, blk.conv1.weight, blk.conv1.bias), blk.bn1, ops.ACT_RELU)
        out = self._bn(_Conv.apply(out, blk.conv2.weight, blk.conv2.bias), blk.bn2, ops.ACT_NONE)
        res = x
        if blk.downsample is not None:
            res = self._bn(_Conv.apply(x, blk.downsample[0].weight, blk.downsample[0].bias), blk.downsample[1], ops.ACT_NONE)
        return _AddRelu.apply(out, res)

    def encode(self, image: torch.Tensor) -> torch.Tensor:
        """(B, 3, H, W) fp32 -> (B, H/2, W/2, 1024) bf16 NHWC feature map (ResNet.forward, transformer.py:126-164)"""
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] % 2 or image.shape[3] % 2 or \
                not _map_tiles(image.shape[2] // 2, image.shape[3] // 2):
            raise ValueError(f"focr Transformer: image must be (B,3,H,W) whose (H/2, W/2) feature map cuts into 128-pixel TMA boxes "
                             f"(W/2 in {{16,32,64,128}}, or a multiple of 32 / 64 with H/2 a multiple of 4 / 2 - e.g. 32x320), got "
                             f"{tuple(image.shape)}")
        ops.require_cuda(image)
        e = self.encoder
        x = _ConvFirst.apply(image.float().contiguous(), e.conv1.weight, e.conv1.bias)
        x = _MaxPool.apply(self._bn(x, e.bn1, ops.ACT_RELU))
        x = self._bn(_Conv.apply(x, e.conv2.weight, e.conv2.bias), e.bn2, ops.ACT_RELU)
        for name, tail in (("layer1", "layer1"), ("layer2", "layer2"), ("layer3", "layer3"), ("layer4", "layer4_conv2")):
            for blk in getattr(e, name):
                x = self._block(x, blk)
            conv = getattr(e, tail + ("_conv" if name != "layer4" else ""))
            bn = getattr(e, tail + "_bn")
            x = self._bn(_Conv.apply(x, conv.weight, conv.bias), bn, ops.ACT_RELU)
        return x

    # -- decoder ------------------------------------------------------------------------------------------------------
    def _lin(self, x, lin: nn.Linear, relu=False):
        return _Linear.apply(x, lin.weight, lin.bias, relu, False)

    def decode_hidden(self, feat: torch.Tensor, text_input: torch.Tensor):
        """feat (B, h, w, 1024) bf16, text_input (B, T) int64 -> (decoder output bf16 (rows_pad, 1024) with rows b*T + t,
        map (B,4,T,h*w)): embedding | PE, then Decoder.forward (transformer.py:346-351, :303-317)"""
        B, T = text_input.shape
        n_tok = feat.shape[1] * feat.shape[2]
        rows_pad = _pad128(B * T)
        p = self.dropout_p if self.training else 0.0
        self._seed = (self._seed * 1103515245 + 12345) & 0x7FFFFFFF
        seed = self._seed
        d = self.decoder
        x0 = _Embed.apply(text_input.contiguous(), self.embedding_word.lut.weight, rows_pad, p, seed, 0)
        mm = d.mask_multihead
        q, k, v = (self._lin(x0, mm.linears[i]) for i in range(3))
        a, _ = _MHA.apply(q, k, v, B, mm.h, mm.d_k, T, T, 1, p, seed, 1)
        r1 = _LN.apply(self._lin(a, mm.linears[3]), x0, d.mul_layernorm1.scale, d.mul_layernorm1.shift)
        mh = d.multihead
        img = feat.reshape(B * n_tok, feat.shape[3])
        q2 = self._lin(r1, mh.linears[0])
        k2, v2 = self._lin(img, mh.linears[1]), self._lin(img, mh.linears[2])
        a2, amap = _MHA.apply(q2, k2, v2, B, mh.h, mh.d_k, T, n_tok, 0, p, seed, 2)
        r2 = _LN.apply(self._lin(a2, mh.linears[3]), r1, d.mul_layernorm2.scale, d.mul_layernorm2.shift)
        hdn = self._lin(r2, d.pff.w_1, relu=True)
        if p > 0:
            hdn = _Dropout.apply(hdn, p, seed, 3)
        r3 = _LN.apply(self._lin(hdn, d.pff.w_2), r2, d.mul_layernorm3.scale, d.mul_layernorm3.shift)
        return r3, amap

    def decode(self, feat: torch.Tensor, text_input: torch.Tensor):
        """-> (logits fp32 (rows_pad, n_class padded to a GEMM tile): generator over the alphabet, map)"""
        r3, amap = self.decode_hidden(feat, text_input)
        g = self.generator_word.proj
        n = g.weight.shape[0]
        npad = (-n) % (64 if n <= 64 else 128)     # one 64-column tile for the stroke alphabet, 128-column tiles beyond
        logits = _Linear.apply(r3, F.pad(g.weight, (0, 0, 0, npad)), F.pad(g.bias, (0, npad)), False, True)
        return logits, amap

    # -- reference forward contract ---------------------------------------------------------------------------------------
    def forward(self, image, text_length, text_input, conv_feature: Optional[torch.Tensor] = None, test: bool = False):
        if conv_feature is None:
            feat = self.encode(image)
        else:  # what this module returned as 'conv' earlier: an NCHW-shaped view of the NHWC map
            feat = conv_feature.permute(0, 2, 3, 1).contiguous()
        conv_out = feat.permute(0, 3, 1, 2)
        if text_length is None:
            return {"conv": conv_out}
        B, T = text_input.shape
        logits, amap = self.decode(feat, text_input)
        full = logits[:B * T].view(B, T, -1)[:, :, :self.word_n_class]
        if test:
            return {"pred": full, "map": amap, "conv": conv_out}
        keep = torch.arange(T, device=full.device)[None, :] < text_length.to(full.device)[:, None]
        return {"pred": full[keep], "map": amap, "conv": conv_out}   # rows in (b, t) order = the reference's packing loop

    def loss(self, image, text_length, text_input, text_gt):
        """fused criterion of train.py:68-71: CE over the packed positions straight from the padded logits"""
        feat = self.encode(image)
        B, T = text_input.shape
        logits, _ = self.decode(feat, text_input)
        return _PackedCE.apply(logits, B, T, self.word_n_class, text_length.contiguous(), text_gt.contiguous())
