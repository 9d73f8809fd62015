"""ORACLE — test infrastructure only.  Torch-fp32 restatement of the reference's frozen CRNN evaluator and the
greedy CTC decode that scores the super-resolved images (scene-text-telescope; text-gestalt is byte-identical):

  parse_crnn_data   interfaces/base.py:319-325       bicubic (32,100) + 0.299R+0.587G+0.114B
  CRNN.forward      model/crnn/crnn.py:25-80          7 convs (+3 eval BN) + 4 max-pools -> (26,B,512) -> 2 x BiLSTM+Linear
  get_crnn_pred     interfaces/super_resolution.py:143-158   argmax, collapse repeats, drop blank (index 0)
  strLabelConverter.decode   utils/utils_crnn.py:54-89        the same collapse on a raw index path

Pinned against the real modules by oracle/make_golden_crnn.py -> tests/golden/crnn_b2.pt."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
ALPHABET = "-0123456789abcdefghijklmnopqrstuvwxyz"  # index 0 = CTC blank (super_resolution.py:144)


def parse_crnn_data(imgs: Tensor) -> Tensor:
    """interfaces/base.py:319-325 (torch bicubic: A = -0.75, align_corners=False)."""
    x = F.interpolate(imgs[:, :3], (32, 100), mode="bicubic")
    return 0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3]


def _bn_eval(x, sd, pre, eps=1e-5):
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"], sd[pre + ".weight"], sd[pre + ".bias"],
                        False, 0.1, eps)


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """nn.LSTM single direction, gate order i, f, g, o.  x: (T,B,nIn) -> (T,B,H)."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = F.linear(x[t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return torch.stack(out, 0)


def _bilstm(x, sd, pre):
    """BidirectionalLSTM, crnn.py:6-22: nn.LSTM(bidirectional) then Linear on the concatenated states."""
    p = pre + ".rnn."
    fwd = _lstm_dir(x, sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"], sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"], False)
    bwd = _lstm_dir(x, sd[p + "weight_ih_l0_reverse"], sd[p + "weight_hh_l0_reverse"], sd[p + "bias_ih_l0_reverse"],
                    sd[p + "bias_hh_l0_reverse"], True)
    rec = torch.cat([fwd, bwd], 2)
    T, B, h = rec.shape
    return F.linear(rec.view(T * B, h), sd[pre + ".embedding.weight"], sd[pre + ".embedding.bias"]).view(T, B, -1)


def crnn_forward(sd: Dict[str, Tensor], x: Tensor, taps: dict = None) -> Tensor:
    """CRNN(32, 1, 37, 256).forward in eval mode: (B,1,32,100) -> (26,B,37)."""
    def conv(i, x, pad):
        return F.conv2d(x, sd[f"cnn.conv{i}.weight"], sd[f"cnn.conv{i}.bias"], padding=pad)
    x = F.max_pool2d(F.relu(conv(0, x, 1)), 2, 2)
    x = F.max_pool2d(F.relu(conv(1, x, 1)), 2, 2)
    x = F.relu(_bn_eval(conv(2, x, 1), sd, "cnn.batchnorm2"))
    x = F.max_pool2d(F.relu(conv(3, x, 1)), (2, 2), (2, 1), (0, 1))
    x = F.relu(_bn_eval(conv(4, x, 1), sd, "cnn.batchnorm4"))
    x = F.max_pool2d(F.relu(conv(5, x, 1)), (2, 2), (2, 1), (0, 1))
    x = F.relu(_bn_eval(conv(6, x, 0), sd, "cnn.batchnorm6"))
    assert x.shape[2] == 1
    seq = x.squeeze(2).permute(2, 0, 1)  # (W, B, C)
    if taps is not None:
        taps["cnn"] = seq.detach()
    y = _bilstm(seq, sd, "rnn.0")
    if taps is not None:
        taps["rnn0"] = y.detach()
    return _bilstm(y, sd, "rnn.1")


def greedy_path(logits_tbc: Tensor) -> Tensor:
    """argmax over classes, lowest index on ties (torch.max semantics): (T,B,C) -> (B,T) int64"""
    return logits_tbc.permute(1, 0, 2).max(2)[1]


def ctc_collapse(path: List[int]) -> List[int]:
    """both reference decoders reduce to this: drop repeats, then blanks (utils_crnn.py:76-79)"""
    out, prev = [], -1
    for i in path:
        if i != 0 and i != prev:
            out.append(int(i))
        prev = i
    return out


def get_crnn_pred(outputs_btc: Tensor) -> List[str]:
    """interfaces/super_resolution.py:143-158 restated (tracks the last EMITTED char, reset on blank)."""
    res = []
    for output in outputs_btc:
        idx = output.max(1)[1].tolist()
        s, last = "", ""
        for i in idx:
            if ALPHABET[i] != last:
                if i != 0:
                    s += ALPHABET[i]
                    last = ALPHABET[i]
                else:
                    last = ""
        res.append(s)
    return res
