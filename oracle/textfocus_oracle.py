"""ORACLE (test infrastructure only - never imported by the product path).

torch-fp32 functional restatement of scene-text-telescope's text-focus loss:
  * loss/text_focus_loss.py:24-37    str_filt
  * loss/text_focus_loss.py:62-81    TextFocusLoss.label_encoder
  * loss/text_focus_loss.py:84-103   TextFocusLoss.forward: mse + 10 * L1(attention maps) + 5e-4 * weighted CE(sr logits)
  * loss/weight_ce_loss.py:10-33     load_confuse_matrix  (37 x 37 weight table from the 62 x 62 confusion counts)
  * loss/weight_ce_loss.py:36-45     weight_cross_entropy
  * loss/transformer.py              the recogniser; the same network as text-gestalt's (oracle/focus_oracle.py restates it)
    with a 37-symbol alphabet and the members named embedding_word / generator_word.
oracle/make_golden_textfocus.py pins this file against the unmodified reference modules."""
from __future__ import annotations

import string
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle import focus_oracle as FO

Tensor = torch.Tensor
ALPHABET = "-0123456789abcdefghijklmnopqrstuvwxyz"                                      # loss/transformer.py:8
LABEL_ALPHABET = "-0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"      # text_focus_loss.py:47


def str_filt(s: str, voc_type: str = "lower") -> str:
    keep = {"digit": string.digits, "lower": string.digits + string.ascii_lowercase,
            "upper": string.digits + string.ascii_letters, "all": string.digits + string.ascii_letters + string.punctuation}
    if voc_type == "lower":
        s = s.lower()
    return "".join(c for c in s if c in keep[voc_type]).lower()


def label_encoder(labels: Sequence[str]):
    """TextFocusLoss.label_encoder on labels already filtered and '-'-terminated (:62-81)"""
    d = {c: i for i, c in enumerate(LABEL_ALPHABET)}
    length = [len(s) for s in labels]
    inp = torch.zeros(len(labels), max(length), dtype=torch.long)
    for i, s in enumerate(labels):
        for j in range(length[i] - 1):
            inp[i, j + 1] = d[s[j]]
    gt = torch.tensor([d[c] for s in labels for c in s], dtype=torch.long)
    return torch.tensor(length, dtype=torch.long), inp, gt


def confuse_weight_table(data: np.ndarray) -> Tensor:
    """load_confuse_matrix (:10-33) on a 62 x 62 confusion-count matrix ordered digits, upper, lower"""
    number, upper, lower = data[:10], data[10:36], data[36:]
    end = np.ones((1, 62))
    pad = np.ones((63, 1))
    re = np.concatenate((end, number, lower, upper), axis=0)
    re = np.concatenate((pad, re), axis=1)
    with np.errstate(divide="ignore"):
        re = 1 / re
    re[re == np.inf] = 1
    t = torch.tensor(re, dtype=torch.float32)
    low = "abcdefghijklmnopqrstuvwxyz"
    for i in range(63):
        for j in range(63):
            if i != j and LABEL_ALPHABET[j] in low:
                t[i][j] = max(t[i][j], t[i][j + 26])
    return t[:37, :37]


def synth_confuse_counts(seed: int = 99) -> np.ndarray:
    """confuse.pkl is git-ignored in the reference: a deterministic stand-in (positive counts, heavy diagonal, some zeros)"""
    rs = np.random.RandomState(seed)
    m = rs.randint(0, 40, size=(62, 62)).astype(np.float64)
    m[rs.random_sample((62, 62)) < 0.15] = 0.0
    m += np.diag(rs.randint(200, 2000, size=62).astype(np.float64))
    return m


def weight_cross_entropy(pred: Tensor, gt: Tensor, table: Tensor) -> Tensor:
    """weight_cross_entropy (:36-45), vectorised; pred (N, 37) logits, gt (N,)"""
    w = table.to(pred.device)[gt]
    pe = w * torch.exp(pred)
    return -(torch.log(pe.gather(1, gt[:, None])[:, 0] / pe.sum(1))).sum() / gt.shape[0]


def _rename(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """view of an STT state dict under the text-gestalt member names that focus_oracle reads"""
    out = dict(sd)
    out["embedding_word_with_upperword.lut.weight"] = sd["embedding_word.lut.weight"]
    out["generator_word_with_upperword.proj.weight"] = sd["generator_word.proj.weight"]
    out["generator_word_with_upperword.proj.bias"] = sd["generator_word.proj.bias"]
    return out


def text_focus_loss(sd, sr_img: Tensor, hr_img: Tensor, labels: Sequence[str], table: Tensor,
                    lambda_attn: float = 10.0, lambda_ce: float = 0.0005, nm: FO.Numerics = FO.FP32,
                    map_hr: Optional[Tensor] = None):
    """TextFocusLoss.forward with args.text_focus on (:84-99) -> (loss, mse, attention_loss, recognition_loss, info).
    nm: numerics mode of the recogniser (focus_oracle.Numerics); map_hr: use this HR attention map instead of computing it"""
    sd = _rename(sd)
    mse = F.mse_loss(sr_img, hr_img)
    labels = [str_filt(s, "lower") + "-" for s in labels]
    length, inp, gt = label_encoder(labels)
    length, inp, gt = length.to(sr_img.device), inp.to(sr_img.device), gt.to(sr_img.device)
    if map_hr is None:
        with torch.no_grad():
            _, map_hr, _ = FO.transformer_forward(sd, FO.to_gray_tensor(hr_img), length, inp, nm)
    sr_pred, map_sr, _ = FO.transformer_forward(sd, FO.to_gray_tensor(sr_img), length, inp, nm)
    att = F.l1_loss(map_hr, map_sr)
    rec = weight_cross_entropy(sr_pred, gt, table)
    loss = mse + att * lambda_attn + rec * lambda_ce
    return loss, mse, att, rec, {"map_hr": map_hr, "map_sr": map_sr, "sr_pred": sr_pred, "text_input": inp, "length": length,
                                 "text_gt": gt}
