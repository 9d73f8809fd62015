"""Label encoders of the trainable recognisers - host-side mirrors of ``util.converter`` of stroke-level-decomposition
(util.py:90-116) and image-ids-CTR (util.py:101-127).  The reference builds its three tensors ON the GPU one element at a time from
Python (a device write + sync per character); here they are assembled on the host with numpy and uploaded once, non-blocking.
Same results, same return order: ``(length, text_input, text_gt, labels)``.

  text_input[i, 0] = 0 (start) and text_input[i, j + 1] = index of the j-th symbol, for all but the last symbol;
  text_gt = the symbols of all samples, concatenated (image-ids-CTR forces the last symbol of each sample to 'END').
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Sequence

import numpy as np
import torch

ALPHABET_STROKE = "<12345$"                                            # stroke-level-decomposition/util.py:14


def load_stroke_table(path: str) -> Dict[str, str]:
    """data/decompose-stroke-3755.txt -> {character: stroke string}  (stroke-level-decomposition/util.py:26-30)"""
    table = {}
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            word, _, strokes = line.split()
            table[word] = strokes.strip()
    return table


def _assemble(symbols: Sequence[Sequence[int]], device):
    B = len(symbols)
    length = np.array([len(s) for s in symbols], np.int64)
    if B == 0 or length.min() < 1:
        raise ValueError("converter: every label needs at least one symbol")
    T = int(length.max())
    text_input = np.zeros((B, T), np.int64)
    for i, s in enumerate(symbols):
        text_input[i, 1:len(s)] = s[:-1]
    text_gt = np.concatenate([np.asarray(s, np.int64) for s in symbols])
    out = [torch.from_numpy(a) for a in (length, text_input, text_gt)]
    if device is not None:
        out = [t.pin_memory().to(device, non_blocking=True) if torch.cuda.is_available() else t.to(device) for t in out]
    return out


def converter_sld(mode: str, label: Sequence[str], character_to_strokelist: Optional[Mapping[str, str]] = None,
                  alp2num_character: Optional[Mapping[str, int]] = None, device="cuda"):
    """stroke-level-decomposition/util.py:90-116.  mode 'stroke': label[i][0] is decomposed through `character_to_strokelist`
    and terminated by '$'; mode 'character': the label string itself over `alp2num_character`."""
    if mode == "stroke":
        if character_to_strokelist is None:
            raise ValueError("converter_sld('stroke', ...) needs the character -> stroke table")
        strings = [character_to_strokelist[i[0]] + "$" for i in label]
        a2n = {c: i for i, c in enumerate(ALPHABET_STROKE)}
    elif mode == "character":
        if alp2num_character is None:
            raise ValueError("converter_sld('character', ...) needs alp2num_character")
        strings, a2n = [i for i in label], alp2num_character
    else:
        raise ValueError(f"unknown mode {mode!r}")
    length, text_input, text_gt = _assemble([[a2n[c] for c in s] for s in strings], device)
    return length, text_input, text_gt, label


def converter_ids(label: Sequence[str], alp2num_character: Mapping[str, int], device="cuda"):
    """image-ids-CTR/util.py:101-127: the last position of every label stands for 'END' in the target (whatever character the
    string carries there), and is never fed to the decoder"""
    end = alp2num_character["END"]
    symbols = []
    for s in label:
        idx = [alp2num_character[c] for c in s[:-1]]
        symbols.append(idx + [end])
    length, text_input, text_gt = _assemble(symbols, device)
    return length, text_input, text_gt, label


def cosine_warm_restarts_lr(epoch: int, base_lr: float = 1.0, T_0: int = 10, T_mult: int = 1, eta_min: float = 0.0) -> float:
    """learning rate of torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=10, T_mult=1) after `epoch` calls of
    scheduler.step() (image-ids-CTR/train.py:29): the value to pass as ``IDSTrainer.step(..., lr=...)``"""
    import math
    if T_mult == 1:
        t_cur, t_i = epoch % T_0, T_0
    else:
        n = int(math.log(epoch / T_0 * (T_mult - 1) + 1, T_mult)) if epoch >= T_0 else 0
        t_cur = epoch - T_0 * (T_mult ** n - 1) // (T_mult - 1)
        t_i = T_0 * T_mult ** n
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2


@torch.no_grad()
def greedy_decode_sld(model, image: torch.Tensor, max_length: int = 30):
    """The autoregressive test-time loop of stroke-level-decomposition/train.py:110-121 on the drop-in module: encoder once
    (features re-entered through ``conv_feature``), then per step the decoder on the growing prefix, arg-max of the last position.
    Returns (pred (B, max_length + 1) int64 with the start symbol 0 in column 0, prob (B, max_length) fp32 = the winning softmax
    probability of each step, sequences (list of B index lists cut after the first '$' = last alphabet entry, start dropped),
    overall_prob (B) = product of the step probabilities up to the cut, train.py:123-137)."""
    B = image.shape[0]
    dev = image.device
    pred = torch.zeros(B, 1, dtype=torch.long, device=dev)
    prob = torch.zeros(B, max_length, dtype=torch.float32, device=dev)
    feats = None
    for i in range(max_length):
        length = torch.full((B,), i + 1, dtype=torch.long, device=dev)
        result = model(image, length, pred, conv_feature=feats, test=True)
        last = torch.softmax(result["pred"][:, -1, :].float(), 1)
        p, now = last.max(1)
        prob[:, i] = p
        pred = torch.cat((pred, now.view(-1, 1)), 1)
        feats = result["conv"]
    end = model.word_n_class - 1
    pred_h, prob_h = pred.cpu(), prob.cpu()
    sequences, overall = [], []
    for b in range(B):
        row = pred_h[b].tolist()
        cut = next((j for j in range(max_length) if row[j] == end), max_length - 1)     # train.py:126-131 scans columns 0..max-1
        kept = row[:cut + 1]
        sequences.append(kept[1:])
        overall.append(float(torch.prod(prob_h[b, :max(len(kept) - 1, 0)])))
    return pred, prob, sequences, overall


@torch.no_grad()
def greedy_decode_ids(model, image: torch.Tensor, text_features: torch.Tensor, max_length: int):
    """The test-time loop of image-ids-CTR/train.py:118-134 on the drop-in module: per step the last position's 2048-d prediction
    is L2-normalised and matched against the text features; arg-max / winning softmax probability as in the reference.
    Returns (pred (B, max_length + 1) int64 with the start symbol in column 0, prob (B, max_length) fp32)."""
    B = image.shape[0]
    dev = image.device
    tf = text_features.to(dev).float()
    pred = torch.zeros(B, 1, dtype=torch.long, device=dev)
    prob = torch.zeros(B, max_length, dtype=torch.float32, device=dev)
    feats = None
    for i in range(max_length):
        length = torch.full((B,), i + 1, dtype=torch.long, device=dev)
        result = model(image, length, pred, conv_feature=feats, test=True)
        prediction = result["pred"][:, -1, :].float()
        prediction = prediction / prediction.norm(dim=1, keepdim=True)
        sm = torch.softmax(prediction @ tf.t(), 1)
        p, now = sm.max(1)
        prob[:, i] = p
        pred = torch.cat((pred, now.view(-1, 1)), 1)
        feats = result["conv"]
    return pred, prob


# ---- KV-cached decode (csrc/decode.cu): the same loops, one launch sequence on the device, one read-back -------------------------
def _decoder_param_table(model):
    """the 29 fp32 tensors focr_recog_decode_prepare takes, in its order"""
    d = model.decoder
    t = [model.embedding_word.lut.weight]
    for lin in d.mask_multihead.linears:
        t += [lin.weight, lin.bias]
    t += [d.mul_layernorm1.scale, d.mul_layernorm1.shift]
    for lin in d.multihead.linears:
        t += [lin.weight, lin.bias]
    t += [d.mul_layernorm2.scale, d.mul_layernorm2.shift, d.pff.w_1.weight, d.pff.w_1.bias, d.pff.w_2.weight, d.pff.w_2.bias,
          d.mul_layernorm3.scale, d.mul_layernorm3.shift, model.generator_word.proj.weight, model.generator_word.proj.bias]
    return [x.detach().float().contiguous() for x in t]


@torch.no_grad()
def _cached_decode(model, image: torch.Tensor, max_length: int, text_features: Optional[torch.Tensor]):
    import ctypes as C
    from . import _lib as L
    if not image.is_cuda:
        raise L.FocrError("focr decode runs on CUDA tensors only (no CPU fallback)")
    was_training = model.training
    model.eval()
    try:
        feat = model.encode(image)                                   # (B, h, w, 1024) bf16, encoder once
    finally:
        model.train(was_training)
    B, n_tok = feat.shape[0], feat.shape[1] * feat.shape[2]
    dev = image.device
    rows = B * n_tok
    rows_pad = (rows + 127) // 128 * 128
    f2 = feat.reshape(rows, 1024)
    if rows_pad != rows:
        f2 = torch.cat([f2, torch.zeros(rows_pad - rows, 1024, dtype=f2.dtype, device=dev)], 0)
    f2 = f2.contiguous()
    params = _decoder_param_table(model)
    vocab = params[0].shape[0]
    n_out = params[27].shape[0]
    tf = None if text_features is None else text_features.to(dev).float().contiguous()
    n_feat = 0 if tf is None else tf.shape[0]
    table = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
    blob = torch.empty(L.lib.focr_recog_decode_prepared_bytes(vocab, n_out, n_feat), dtype=torch.uint8, device=dev)
    st = L.cur_stream()
    with torch.cuda.device(dev):
        L.check(L.lib.focr_recog_decode_prepare(table, vocab, n_out, None if tf is None else tf.data_ptr(), n_feat, blob.data_ptr(),
                                                blob.numel(), st), "recog_decode_prepare")
        ws = torch.empty(L.lib.focr_recog_decode_workspace_bytes(B, n_tok, max_length, n_out, n_feat), dtype=torch.uint8, device=dev)
        pred = torch.empty(B, max_length + 1, dtype=torch.long, device=dev)
        prob = torch.empty(B, max_length, dtype=torch.float32, device=dev)
        L.check(L.lib.focr_recog_decode(blob.data_ptr(), blob.numel(), vocab, n_out, n_feat, f2.data_ptr(), B, n_tok, max_length,
                                        pred.data_ptr(), prob.data_ptr(), ws.data_ptr(), ws.numel(), st), "recog_decode")
    return pred, prob


def greedy_decode_sld_cached(model, image: torch.Tensor, max_length: int = 30):
    """`greedy_decode_sld` with the decoder run incrementally on the device (K / V caches, arg-max and end bookkeeping in kernels):
    same return values.  stroke-level-decomposition/train.py:110-137."""
    pred, prob = _cached_decode(model, image, max_length, None)
    end = model.word_n_class - 1
    pred_h, prob_h = pred.cpu(), prob.cpu()                          # the one read-back
    sequences, overall = [], []
    for b in range(pred_h.shape[0]):
        row = pred_h[b].tolist()
        cut = next((j for j in range(max_length) if row[j] == end), max_length - 1)
        kept = row[:cut + 1]
        sequences.append(kept[1:])
        overall.append(float(torch.prod(prob_h[b, :max(len(kept) - 1, 0)])))
    return pred, prob, sequences, overall


def greedy_decode_ids_cached(model, image: torch.Tensor, text_features: torch.Tensor, max_length: int):
    """`greedy_decode_ids` on the KV-cached device loop: (pred (B, max_length + 1), prob (B, max_length)).
    image-ids-CTR/train.py:118-134."""
    return _cached_decode(model, image, max_length, text_features)
