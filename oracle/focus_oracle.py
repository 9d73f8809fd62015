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


def _conv_bn(sd, conv: str, bn: str, x: Tensor, relu: bool) -> Tensor:
    y = F.conv2d(x, sd[conv + ".weight"], sd[conv + ".bias"], stride=1, padding=1)
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     training=False, eps=1e-5)
    return F.relu(y) if relu else y


def _basic_block(sd, pre: str, x: Tensor, has_down: bool) -> Tensor:
    """BasicBlock.forward (:317-334)"""
    out = _conv_bn(sd, pre + ".conv1", pre + ".bn1", x, True)
    out = _conv_bn(sd, pre + ".conv2", pre + ".bn2", out, False)
    res = _conv_bn(sd, pre + ".downsample.0", pre + ".downsample.1", x, False) if has_down else x
    return F.relu(out + res)


def resnet_encoder(sd: Dict[str, Tensor], gray: Tensor, pre: str = "encoder.cnn") -> Tensor:
    """ResNet.forward (:130-168); only the first two max-pools are active.  (B,1,32,128) -> (B,1024,8,32)"""
    x = _conv_bn(sd, f"{pre}.conv1", f"{pre}.bn1", gray, True)
    x = F.max_pool2d(x, 2, 2)
    x = _conv_bn(sd, f"{pre}.conv2", f"{pre}.bn2", x, True)
    x = F.max_pool2d(x, 2, 2)
    for li, (nblk, (cin, cout)) in enumerate(zip(LAYERS, PLANES), start=1):
        for bi in range(nblk):
            x = _basic_block(sd, f"{pre}.layer{li}.{bi}", x, has_down=(bi == 0 and cin != cout))
        if li < 4:
            x = _conv_bn(sd, f"{pre}.layer{li}_conv", f"{pre}.layer{li}_bn", x, True)
        else:
            x = _conv_bn(sd, f"{pre}.layer4_conv2", f"{pre}.layer4_conv2_bn", x, True)
    return x


def _mha(sd, pre: str, query: Tensor, key: Tensor, value: Tensor, mask: Optional[Tensor], h: int = 16):
    """MultiHeadedAttention.forward + attention() (:26-79), dropout off (eval)"""
    nb, d = query.size(0), query.size(-1)
    dk = d // h
    q, k, v = [F.linear(x, sd[f"{pre}.linears.{i}.weight"], sd[f"{pre}.linears.{i}.bias"]).view(nb, -1, h, dk).transpose(1, 2)
               for i, x in enumerate((query, key, value))]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    if mask is not None:
        scores = scores.masked_fill(mask.unsqueeze(1) == 0, float("-inf"))
    p = F.softmax(scores, dim=-1)
    x = torch.matmul(p, v).transpose(1, 2).contiguous().view(nb, -1, h * dk)
    return F.linear(x, sd[f"{pre}.linears.3.weight"], sd[f"{pre}.linears.3.bias"]), p


def text_embedding(sd, text_input: Tensor) -> Tensor:
    """Transformer.forward (:365-369): lut(x)*sqrt(512) || pe(zeros)  -> (B,T,1024)"""
    emb = F.embedding(text_input, sd["embedding_word_with_upperword.lut.weight"]) * math.sqrt(512)
    T = text_input.shape[1]
    pe = sd["pe.pe"][:, :T].expand(emb.shape[0], -1, -1)
    return torch.cat([emb, pe], 2)


def decoder_query(sd, text: Tensor, pre: str = "decoder") -> Tensor:
    """the image-independent part of Decoder.forward (:289-293): LN1(text + masked self-attention)"""
    T = text.shape[1]
    mask = torch.tril(torch.ones(1, T, T, dtype=torch.bool, device=text.device))
    att, _ = _mha(sd, f"{pre}.mask_multihead", text, text, text, mask)
    return layer_norm_std(text + att, sd[f"{pre}.mul_layernorm1.a_2"], sd[f"{pre}.mul_layernorm1.b_2"])


def decoder(sd, text: Tensor, conv_feature: Tensor, pre: str = "decoder") -> Tuple[Tensor, Tensor]:
    """Decoder.forward (:289-304) -> (result (B,T,1024), attention_map (B,16,T,256))"""
    result = decoder_query(sd, text, pre)
    b, c, hh, ww = conv_feature.shape
    tokens = conv_feature.view(b, c, hh * ww).permute(0, 2, 1).contiguous()
    align, amap = _mha(sd, f"{pre}.multihead", result, tokens, tokens, None)
    result = layer_norm_std(result + align, sd[f"{pre}.mul_layernorm2.a_2"], sd[f"{pre}.mul_layernorm2.b_2"])
    ff = F.linear(F.relu(F.linear(result, sd[f"{pre}.pff.w_1.weight"], sd[f"{pre}.pff.w_1.bias"])),
                  sd[f"{pre}.pff.w_2.weight"], sd[f"{pre}.pff.w_2.bias"])
    result = layer_norm_std(result + ff, sd[f"{pre}.mul_layernorm3.a_2"], sd[f"{pre}.mul_layernorm3.b_2"])
    return result, amap


def transformer_forward(sd, image: Tensor, text_length: Tensor, text_input: Tensor):
    """Transformer.forward with test=False (:356-393) -> (probs_res (sum len, 10), attention map, correct_list)"""
    feat = resnet_encoder(sd, image)
    text = text_embedding(sd, text_input)
    res, amap = decoder(sd, text, feat)
    logits = F.linear(res, sd["generator_word_with_upperword.proj.weight"], sd["generator_word_with_upperword.proj.bias"])
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
                      stroke_lambda: float = 50.0):
    """StrokeFocusLoss.forward with args.text_focus on (:83-118) -> (loss, mse, attention_loss, info)"""
    mse = F.mse_loss(sr_img, hr_img)
    length, inp, _ = label_stroke_encoder(labels, dic)
    length, inp = length.to(sr_img.device), inp.to(sr_img.device)
    _, map_hr, _ = transformer_forward(sd, to_gray_tensor(hr_img), length, inp)
    _, map_sr, _ = transformer_forward(sd, to_gray_tensor(sr_img), length, inp)
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
