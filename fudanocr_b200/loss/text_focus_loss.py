"""TextFocusLoss on the focr sm_100a engine — drop-in for ``loss.text_focus_loss.TextFocusLoss`` of scene-text-telescope
(scene-text-telescope/loss/text_focus_loss.py:40-103).

``TextFocusLoss(args)`` with ``args.text_focus``; ``forward(sr_img, hr_img, label) -> (loss, mse_loss, attention_loss,
recognition_loss)`` = ``mse + 10 * L1(attention maps) + 0.0005 * weight_cross_entropy(sr logits, gt)``.  Value and gradient
w.r.t. ``sr_img`` come from one C-ABI call (``focr_text_focus_loss``); no PyTorch fallback."""
from __future__ import annotations

import string
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .. import _lib as L
from .stroke_focus_loss import _FocusBase, _MseFn
from .weight_ce_loss import confuse_weight_table, load_confuse_matrix

__all__ = ["TextFocusLoss", "str_filt"]


def str_filt(str_: str, voc_type: str) -> str:           # text_focus_loss.py:24-37
    alpha_dict = {"digit": string.digits, "lower": string.digits + string.ascii_lowercase,
                  "upper": string.digits + string.ascii_letters, "all": string.digits + string.ascii_letters + string.punctuation}
    if voc_type == "lower":
        str_ = str_.lower()
    for char in str_:
        if char not in alpha_dict[voc_type]:
            str_ = str_.replace(char, "")
    return str_.lower()


class _TextFocusFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sr, hr, owner, enc):
        losses, d_sr = owner._run_text(sr, hr, enc, 1.0)
        ctx.save_for_backward(d_sr)
        loss, mse, att, rec = [losses[i].clone() for i in range(4)]
        ctx.mark_non_differentiable(mse, att, rec)
        return loss, mse, att, rec

    @staticmethod
    def backward(ctx, g0, g1, g2, g3):
        (d_sr,) = ctx.saved_tensors
        return d_sr * g0, None, None, None


class TextFocusLoss(_FocusBase):
    variant = "stt"
    lambda_attn, lambda_ce = 10.0, 0.0005                    # text_focus_loss.py:97

    def __init__(self, args, confuse_counts: Optional[np.ndarray] = None,
                 transformer_state_dict: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__()
        self.args = args
        self.english_alphabet = "-0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
        self.english_dict = {c: i for i, c in enumerate(self.english_alphabet)}
        table = load_confuse_matrix() if confuse_counts is None else confuse_weight_table(np.asarray(confuse_counts, dtype=np.float64))
        self.register_buffer("weight_table", table)
        self.build_up_transformer(transformer_state_dict)

    def build_up_transformer(self, transformer_state_dict=None):   # text_focus_loss.py:54-60
        self._init_engine(transformer_state_dict, "./dataset/mydata/pretrain_transformer.pth")

    def label_encoder(self, label: Sequence[str], device=None):
        """text_focus_loss.py:62-81 -> (length (B,), right-shifted indices (B,Tmax), text_gt (sum len,)), int64"""
        length = [len(i) for i in label]
        inp = torch.zeros(len(label), max(length), dtype=torch.long)
        for i, s in enumerate(label):
            for j in range(length[i] - 1):
                inp[i, j + 1] = self.english_dict[s[j]]
        gt = torch.tensor([self.english_dict[c] for s in label for c in s], dtype=torch.long)
        length_t = torch.tensor(length, dtype=torch.long)
        if device is not None:
            length_t, inp, gt = length_t.to(device), inp.to(device), gt.to(device)
        return length_t, inp, gt

    def _run_text(self, sr, hr, enc, gscale: float, d_sr: Optional[torch.Tensor] = None, outputs: bool = False):
        if not (sr.is_cuda and hr.is_cuda):
            raise L.FocrError("focr losses run on CUDA tensors only (no CPU fallback)")
        dev = sr.device
        length, text_input, text_gt = [t.to(dev).contiguous() for t in enc]
        sr_c, hr_c = sr.detach().contiguous().float(), hr.detach().contiguous().float()
        B, T = text_input.shape
        assert sr_c.shape == (B, 3, 32, 128) and hr_c.shape == sr_c.shape, (sr_c.shape, hr_c.shape)
        blob = self._prepare(dev)
        ws = self._workspace(B, T, dev)
        if d_sr is None:
            d_sr = torch.empty_like(sr_c)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        nc = self.transformer.n_class
        mh = torch.empty(B, 16, T, 256, dtype=torch.float32, device=dev) if outputs else None
        ms = torch.empty(B, 16, T, 256, dtype=torch.float32, device=dev) if outputs else None
        pred = torch.empty(int(text_gt.numel()), nc, dtype=torch.float32, device=dev) if outputs else None
        table = self.weight_table.to(dev).contiguous()
        L.check(L.lib.focr_text_focus_loss(blob.data_ptr(), blob.numel(), nc, sr_c.data_ptr(), hr_c.data_ptr(),
                                           text_input.data_ptr(), length.data_ptr(), text_gt.data_ptr(), table.data_ptr(),
                                           B, T, self.lambda_attn, self.lambda_ce, float(gscale), d_sr.data_ptr(),
                                           losses.data_ptr(), L.ptr(mh), L.ptr(ms), L.ptr(pred), ws.data_ptr(), ws.numel(),
                                           L.cur_stream()), "text_focus_loss")
        if outputs:
            return losses, d_sr, mh, ms, pred
        return losses, d_sr

    def forward(self, sr_img, hr_img, label):
        if not self.args.text_focus:                                 # text_focus_loss.py:99-103
            mse = _MseFn.apply(sr_img, hr_img)
            return mse, mse, -1, -1
        label = [str_filt(i, "lower") + "-" for i in label]
        enc = self.label_encoder(label, sr_img.device)
        return _TextFocusFn.apply(sr_img, hr_img, self, enc)

    def loss_and_grad(self, sr_img, hr_img, label, gscale: float, d_sr: torch.Tensor):
        """fused-trainer entry: losses (device, [loss, mse, attention, recognition]); d_sr = gscale * dloss/dsr in place"""
        label = [str_filt(i, "lower") + "-" for i in label]
        enc = self.label_encoder(label, sr_img.device)
        losses, _ = self._run_text(sr_img, hr_img, enc, gscale, d_sr=d_sr)
        return losses
