"""ORACLE (test infrastructure only - never imported by the product path).

torch-fp32 functional restatement of text-gestalt's stroke-focus loss:
  * loss/stroke_focus_loss.py:12-18   to_gray_tensor
  * loss/stroke_focus_loss.py:49-80   label_stroke_encoder
  * loss/stroke_focus_loss.py:83-122  StrokeFocusLoss.forward  (mse + stroke_lambda * L1(attention maps))
  * loss/transformer_english_decomposition.py:70-168   ResNet encoder, layers [1,2,5,3], eval-mode BN
  * loss/transformer_english_decomposition.py:276-304  Decoder (masked MHA -> LN -> cross MHA -> LN -> FFN -> LN)
  * loss/transformer_english_decomposition.py:343-398  Transformer.forward (embedding*sqrt(512) || PE, generator, packing)
The recogniser is frozen and in eval() mode (stroke_focus_loss.py:45): BatchNorm uses running statistics, dropout is off.
State-dict keys are the reference's (without the DataParallel 'module.' prefix).  oracle/make_golden_focus.py pins
this file against the unmodified reference classes; parity of the CUDA path is then checked against it.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
STROKE_ALPHABET = "0123456789"      # transformer_english_decomposition.py:8
LAYERS = [1, 2, 5, 3]               # Encoder: ResNet(num_in=1, block=BasicBlock, layers=[1,2,5,3])  (:337-341)
PLANES = [(128, 256), (256, 256), (256, 512), (512, 512)]


def to_gray_tensor(t: Tensor) -> Tensor:
    return 0.299 * t[:, 0:1] + 0.587 * t[:, 1:2] + 0.114 * t[:, 2:3]


def positional_encoding(d_model: int, max_len: int) -> Tensor:
    """PositionalEncoding buffer 'pe' (1, max_len, d_model)  (:199-213)"""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len).unsqueeze(1).float()
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def layer_norm_std(x: Tensor, a: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """the reference's LayerNorm: unbiased std, eps added to std  (:222-234)"""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a * (x - mean) / (std + eps) + b


class Numerics:
    """How the restatement evaluates the recogniser.

    Default (`Numerics()`): plain fp32, conv followed by eval-mode BatchNorm - the form pinned against the reference.
    `Numerics(fold=True)`: BatchNorm folded into the conv weights/bias (algebraically identical; pinned against the
    golden fixture on CPU).  `Numerics.bf16_emulation()`: the folded form with every tensor the sm_100a engine stores in
    bf16 rounded to bf16 at the same point (weights, activations after the fused epilogue, gradients after each input-
    gradient GEMM), arithmetic still fp32.  The recogniser with random weights is chaotic (a 2^-9 relative perturbation of
    the weights moves the encoder output by ~10-30 %), so end-to-end agreement with fp32 says little about a bf16
    kernel; agreement with this emulation (same rounding points => same trajectory) is the sharp test."""

    def __init__(self, fold: bool = False, round_act=None, round_grad: bool = False, teacher: Optional[Dict[str, Tensor]] = None):
        self.fold = fold
        self._ra = round_act
        self.round_grad = round_grad
        self.teacher = teacher

    @staticmethod
    def bf16_emulation() -> "Numerics":
        return Numerics(fold=True, round_act=lambda t: t.to(torch.bfloat16).to(torch.float32), round_grad=True)

    @staticmethod
    def teacher_forced(acts: Dict[str, Tensor]) -> "Numerics":
        """Linearise the restatement AT THE ENGINE'S OWN ACTIVATIONS: every stored activation named in `acts` replaces
        the restatement's value (straight-through), with the ReLU mask taken from the engine's value, so autograd then
        evaluates exactly the linear map the engine's input-gradient chain implements (same masks, same pooling argmax),
        free of the chaotic forward divergence."""
        return Numerics(fold=True, teacher=acts)

    def w(self, t: Tensor) -> Tensor:           # a weight the engine keeps in bf16
        return t if self._ra is None else self._ra(t)

    def act(self, pre: Tensor, relu: bool, key: Optional[str] = None) -> Tensor:
        """activation epilogue of one layer: (ReLU) then the storage rounding / teacher replacement"""
        if self.teacher is not None and key in self.teacher:
            eng = self.teacher[key]
            lin = pre * (eng > 0).to(pre.dtype) if relu else pre
            return lin + (eng - lin).detach()
        return self.a(F.relu(pre) if relu else pre)

    def a(self, t: Tensor) -> Tensor:           # an activation the engine stores in bf16 (straight-through rounding)
        if self._ra is None:
            return t
        y = t + (self._ra(t) - t).detach()
        if self.round_grad and y.requires_grad:
            y.register_hook(lambda g: g.to(torch.bfloat16).to(torch.float32))
        return y

    def g(self, t: Tensor) -> Tensor:           # identity whose gradient is rounded (a gradient buffer of the engine)
        if not (self.round_grad and t.requires_grad):
            return t
        y = t.clone()
        y.register_hook(lambda g: g.to(torch.bfloat16).to(torch.float32))
        return y


FP32 = Numerics()


def _conv_bn_pre(sd, conv: str, bn: str, x: Tensor, nm: Numerics, fp32_weights: bool = False) -> Tensor:
    """conv3x3 + eval-mode BatchNorm, before any activation"""
    if not nm.fold:
        y = F.conv2d(x, sd[conv + ".weight"], sd[conv + ".bias"], stride=1, padding=1)
        return F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                            training=False, eps=1e-5)
    sc = sd[bn + ".weight"] * torch.rsqrt(sd[bn + ".running_var"] + 1e-5)
    w = sd[conv + ".weight"] * sc.view(-1, 1, 1, 1)
    b = (sd[conv + ".bias"] - sd[bn + ".running_mean"]) * sc + sd[bn + ".bias"]
    return F.conv2d(x, w if fp32_weights else nm.w(w), b, stride=1, padding=1)


def _conv_bn(sd, conv: str, bn: str, x: Tensor, relu: bool, nm: Numerics = FP32, fp32_weights: bool = False) -> Tensor:
    return nm.act(_conv_bn_pre(sd, conv, bn, x, nm, fp32_weights), relu, key=conv)


def _basic_block(sd, pre: str, x: Tensor, has_down: bool, nm: Numerics = FP32) -> Tensor:
    """BasicBlock.forward (:317-334)"""
    out = _conv_bn(sd, pre + ".conv1", pre + ".bn1", x, True, nm)
    res = _conv_bn(sd, pre + ".downsample.0", pre + ".downsample.1", nm.g(x), False, nm) if has_down else x
    out = _conv_bn_pre(sd, pre + ".conv2", pre + ".bn2", out, nm)
    return nm.act(out + res, True, key=pre + ".conv2")


def resnet_encoder(sd: Dict[str, Tensor], gray: Tensor, pre: str = "encoder.cnn", nm: Numerics = FP32) -> Tensor:
    """ResNet.forward (:130-168); only the first two max-pools are active.  (B,1,32,128) -> (B,1024,8,32)"""
    x = _conv_bn(sd, f"{pre}.conv1", f"{pre}.bn1", gray, True, nm, fp32_weights=True)   # the engine's stem is fp32 SIMT
    x = F.max_pool2d(x, 2, 2)
    x = _conv_bn(sd, f"{pre}.conv2", f"{pre}.bn2", nm.g(x), True, nm)
    x = F.max_pool2d(x, 2, 2)
    x = nm.g(x)
    for li, (nblk, (cin, cout)) in enumerate(zip(LAYERS, PLANES), start=1):
        for bi in range(nblk):
            x = _basic_block(sd, f"{pre}.layer{li}.{bi}", x, has_down=(bi == 0 and cin != cout), nm=nm)
        if li < 4:
            x = _conv_bn(sd, f"{pre}.layer{li}_conv", f"{pre}.layer{li}_bn", x, True, nm)
        else:
            x = _conv_bn(sd, f"{pre}.layer4_conv2", f"{pre}.layer4_conv2_bn", x, True, nm)
    return x


def _lin(sd, name: str, x: Tensor, nm: Numerics) -> Tensor:
    return F.linear(x, nm.w(sd[name + ".weight"]), sd[name + ".bias"])


def _mha(sd, pre: str, query: Tensor, key: Tensor, value: Tensor, mask: Optional[Tensor], h: int = 16, nm: Numerics = FP32,
         need_out: bool = True):
    """MultiHeadedAttention.forward + attention() (:26-79), dropout off (eval)"""
    nb, d = query.size(0), query.size(-1)
    dk = d // h
    q, k, v = [nm.act(_lin(sd, f"{pre}.linears.{i}", x, nm), False, key=f"{pre}.linears.{i}").view(nb, -1, h, dk).transpose(1, 2)
               for i, x in enumerate((query, key, value))]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    if mask is not None:
        scores = scores.masked_fill(mask.unsqueeze(1) == 0, float("-inf"))
    p = F.softmax(scores, dim=-1)
    if nm.teacher is not None and f"{pre}.map" in nm.teacher:
        p = p + (nm.teacher[f"{pre}.map"] - p).detach()
    if not need_out:
        return None, p
    x = nm.act(torch.matmul(p, v).transpose(1, 2).contiguous().view(nb, -1, h * dk), False, key=f"{pre}.ctx")
    return _lin(sd, f"{pre}.linears.3", x, nm), p


def text_embedding(sd, text_input: Tensor, nm: Numerics = FP32) -> Tensor:
    """Transformer.forward (:365-369): lut(x)*sqrt(512) || pe(zeros)  -> (B,T,1024)"""
    emb = F.embedding(text_input, sd["embedding_word_with_upperword.lut.weight"]) * math.sqrt(512)
    T = text_input.shape[1]
    pe = sd["pe.pe"][:, :T].expand(emb.shape[0], -1, -1)
    return nm.a(torch.cat([emb, pe], 2))


def decoder_query(sd, text: Tensor, pre: str = "decoder", nm: Numerics = FP32) -> Tensor:
    """the image-independent part of Decoder.forward (:289-293): LN1(text + masked self-attention)"""
    T = text.shape[1]
    mask = torch.tril(torch.ones(1, T, T, dtype=torch.bool, device=text.device))
    att, _ = _mha(sd, f"{pre}.mask_multihead", text, text, text, mask, nm=nm)
    return nm.a(layer_norm_std(nm.a(text + att), sd[f"{pre}.mul_layernorm1.a_2"], sd[f"{pre}.mul_layernorm1.b_2"]))


def decoder(sd, text: Tensor, conv_feature: Tensor, pre: str = "decoder", nm: Numerics = FP32,
            map_only: bool = False) -> Tuple[Optional[Tensor], Tensor]:
    """Decoder.forward (:289-304) -> (result (B,T,1024), attention_map (B,16,T,256))"""
    result = decoder_query(sd, text, pre, nm)
    b, c, hh, ww = conv_feature.shape
    tokens = conv_feature.view(b, c, hh * ww).permute(0, 2, 1).contiguous()
    align, amap = _mha(sd, f"{pre}.multihead", result, tokens, tokens, None, nm=nm, need_out=not map_only)
    if map_only:
        return None, amap
    x2 = nm.act(result + align, False, key=f"{pre}.x2")
    result = nm.act(layer_norm_std(x2, sd[f"{pre}.mul_layernorm2.a_2"], sd[f"{pre}.mul_layernorm2.b_2"]), False, key=f"{pre}.r2")
    hidden = nm.act(_lin(sd, f"{pre}.pff.w_1", result, nm), True, key=f"{pre}.pff.w_1")
    x3 = nm.act(result + _lin(sd, f"{pre}.pff.w_2", hidden, nm), False, key=f"{pre}.x3")
    result = nm.act(layer_norm_std(x3, sd[f"{pre}.mul_layernorm3.a_2"], sd[f"{pre}.mul_layernorm3.b_2"]), False, key=f"{pre}.r3")
    return result, amap


def attention_map(sd, image: Tensor, text_input: Tensor, nm: Numerics = FP32) -> Tensor:
    """the word-attention map alone (what the stroke-focus loss reads), under the numerics mode `nm`"""
    feat = resnet_encoder(sd, image, nm=nm)
    _, amap = decoder(sd, text_embedding(sd, text_input, nm), feat, nm=nm, map_only=True)
    return amap


def transformer_forward(sd, image: Tensor, text_length: Tensor, text_input: Tensor, nm: Numerics = FP32):
    """Transformer.forward with test=False (:356-393) -> (probs_res (sum len, 10), attention map, correct_list)"""
    feat = resnet_encoder(sd, image, nm=nm)
    text = text_embedding(sd, text_input, nm)
    res, amap = decoder(sd, text, feat, nm=nm)
    logits = _lin(sd, "generator_word_with_upperword.proj", res, nm)
    rows, correct = [], []
    for i, ln in enumerate(text_length.tolist()):
        r = logits[i, :ln]
        rows.append(r)
        correct.append(bool((r.max(1)[1][:-1] == text_input[i][1:ln]).all()))
    return torch.cat(rows, 0), amap, correct


def label_stroke_encoder(labels: Sequence[str], dic: Dict[str, str]):
    """StrokeFocusLoss.label_stroke_encoder (:49-80) -> (length (B,), input (B,Tmax) right-shifted digits, text_gt)"""
    seqs = ["".join(dic[c] for c in lab if c in dic) + "0" for lab in labels]
    length = [len(s) for s in seqs]
    tmax = max(length)
    inp = torch.zeros(len(seqs), tmax, dtype=torch.long)
    for i, s in enumerate(seqs):
        for j in range(length[i] - 1):
            inp[i, j + 1] = STROKE_ALPHABET.index(s[j])
    gt = torch.tensor([STROKE_ALPHABET.index(c) for s in seqs for c in s], dtype=torch.long)
    return torch.tensor(length, dtype=torch.long), inp, gt


def stroke_focus_loss(sd, sr_img: Tensor, hr_img: Tensor, labels: Sequence[str], dic: Dict[str, str],
                      stroke_lambda: float = 50.0, nm: Optional[Numerics] = None):
    """StrokeFocusLoss.forward with args.text_focus on (:83-118) -> (loss, mse, attention_loss, info).
    nm=None: the full Transformer.forward of the reference (pinned form); otherwise only the attention maps under `nm`."""
    mse = F.mse_loss(sr_img, hr_img)
    length, inp, _ = label_stroke_encoder(labels, dic)
    length, inp = length.to(sr_img.device), inp.to(sr_img.device)
    if nm is None:
        _, map_hr, _ = transformer_forward(sd, to_gray_tensor(hr_img), length, inp)
        _, map_sr, _ = transformer_forward(sd, to_gray_tensor(sr_img), length, inp)
    else:
        with torch.no_grad():
            map_hr = attention_map(sd, to_gray_tensor(hr_img), inp, nm)
        map_sr = attention_map(sd, to_gray_tensor(sr_img), inp, nm)
    att = F.l1_loss(map_hr, map_sr)
    return mse + att * stroke_lambda, mse, att, {"map_hr": map_hr, "map_sr": map_sr, "text_input": inp, "length": length}


def synth_decomposition() -> Dict[str, str]:
    """english_decomposition.txt is git-ignored in the reference: a deterministic stand-in with the
    same format (character -> stroke-digit string over 1..9), 1-4 strokes per character"""
    chars = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    dic = {}
    for i, c in enumerate(chars):
        n = 1 + (i * 7) % 4
        dic[c] = "".join(str(1 + (i * 3 + 5 * k) % 9) for k in range(n))
    return dic


def synth_recogniser_state_dict(spec: Dict[str, list], bn_stats: Optional[Dict[str, Tensor]] = None, seed: int = 777):
    """Deterministic synthetic recogniser weights: oracle.synth rules, the last BatchNorm gain of every residual block
    scaled by 0.25 (near-identity blocks, as in a trained ResNet; keeps the random network from being needlessly
    chaotic), and - when given - the calibrated BatchNorm running statistics stored in the golden fixture."""
    import re
    from oracle import synth
    sd = synth.synth_state_dict(spec, seed=seed, computed={"pe.pe": positional_encoding(512, 5000)})
    for k in sd:
        if re.search(r"layer\d\.\d+\.bn2\.weight$", k):
            sd[k] = sd[k] * 0.25
    if bn_stats is not None:
        sd.update(bn_stats)
    return sd
