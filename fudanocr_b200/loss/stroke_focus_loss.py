"""StrokeFocusLoss on the focr sm_100a engine — drop-in for ``loss.stroke_focus_loss.StrokeFocusLoss`` of text-gestalt
(text-gestalt/loss/stroke_focus_loss.py:20-122).

Same constructor (``StrokeFocusLoss(args)`` with ``args.text_focus`` / ``args.stroke_lambda``), same
``forward(sr_img, hr_img, label) -> (loss, mse_loss, attention_loss, recognition_loss)``, same label encoder.  The value
AND the gradient w.r.t. ``sr_img`` come from one C-ABI call (``focr_focus_loss``): HR forward, SR forward and the SR
input-gradient chain of the frozen recogniser run as hand-written kernels; ``loss.backward()`` then just scales the
stored gradient.  The weight gradients the reference's autograd also produces for the frozen recogniser (and never
reads) are not computed.  No PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import torch
from torch import nn

from .. import _lib as L
from .transformer_english_decomposition import TG_ALPHABET, Transformer

__all__ = ["StrokeFocusLoss", "to_gray_tensor"]


class _ToGray(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        if not t.is_cuda:
            raise L.FocrError("to_gray_tensor runs on CUDA tensors only (no CPU fallback)")
        if t.dim() != 4 or t.shape[1] < 3:
            raise ValueError(f"to_gray_tensor expects (B, >=3, H, W), got {tuple(t.shape)}")
        x = t.detach().float().contiguous()
        b, c, h, w = x.shape
        out = torch.empty(b, 1, h, w, dtype=torch.float32, device=x.device)
        L.check(L.lib.focr_to_gray(x.data_ptr(), out.data_ptr(), b, c, h * w, L.cur_stream()), "to_gray")
        ctx.shape = (b, c, h, w)
        return out

    @staticmethod
    def backward(ctx, g):
        b, c, h, w = ctx.shape
        g = g.float().contiguous()
        d = torch.empty(b, c, h, w, dtype=torch.float32, device=g.device)
        L.check(L.lib.focr_to_gray_bwd(g.data_ptr(), d.data_ptr(), b, c, h * w, L.cur_stream()), "to_gray_bwd")
        return d


def to_gray_tensor(tensor: torch.Tensor) -> torch.Tensor:
    """0.299 R + 0.587 G + 0.114 B, (B, >=3, H, W) -> (B, 1, H, W)  (stroke_focus_loss.py:12-18; inside the fused loss the same
    arithmetic is part of the recogniser stem, conv1_fwd_kernel)"""
    return _ToGray.apply(tensor)


class _FocusFn(torch.autograd.Function):
    """loss triple from focr_focus_loss; backward scales the gradient the same call produced"""

    @staticmethod
    def forward(ctx, sr, hr, owner, text_input, lam):
        losses, d_sr = owner._run(sr, hr, text_input, lam, 1.0)
        ctx.save_for_backward(d_sr)
        loss, mse, att = losses[0].clone(), losses[1].clone(), losses[2].clone()
        ctx.mark_non_differentiable(mse, att)
        return loss, mse, att

    @staticmethod
    def backward(ctx, g0, g1, g2):
        (d_sr,) = ctx.saved_tensors
        return d_sr * g0, None, None, None, None


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sr, hr):
        if not sr.is_cuda:
            raise L.FocrError("focr losses run on CUDA tensors only (no CPU fallback)")
        sr_c, hr_c = sr.contiguous().float(), hr.contiguous().float()
        d_sr = torch.empty_like(sr_c)
        loss = torch.empty(1, dtype=torch.float32, device=sr.device)
        scratch = torch.empty(1 << 16, dtype=torch.uint8, device=sr.device)
        L.check(L.lib.focr_mse_loss_grad(sr_c.data_ptr(), hr_c.data_ptr(), d_sr.data_ptr(), loss.data_ptr(), sr_c.numel(),
                                         1.0, scratch.data_ptr(), scratch.numel(), L.cur_stream()), "mse_loss_grad")
        ctx.save_for_backward(d_sr)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_sr,) = ctx.saved_tensors
        return d_sr * g, None


class _FocusBase(nn.Module):
    """shared machinery of StrokeFocusLoss (text-gestalt) and the attention term of TextFocusLoss (scene-text-telescope)"""
    variant = "tg"

    def _init_engine(self, transformer_state_dict: Optional[Dict[str, torch.Tensor]], weights_path: str):
        transformer = Transformer(self.variant)
        sd = transformer_state_dict
        if sd is None:
            sd = torch.load(weights_path, map_location="cpu")      # the reference's asset (DataParallel-prefixed)
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
        transformer.load_state_dict(sd)
        transformer.eval()
        for p in transformer.parameters():
            p.requires_grad_(False)                                 # frozen: the engine never forms its weight gradients
        self.transformer = transformer
        self._prepared: Optional[torch.Tensor] = None
        self._ws: Optional[torch.Tensor] = None
        self._ws_key = None

    def _prepare(self, dev: torch.device) -> torch.Tensor:
        if self._prepared is None or self._prepared.device != dev:
            self.transformer.to(dev)
            tensors = self.transformer.slot_tensors()
            table = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
            nbytes = L.lib.focr_strokenet_prepared_bytes(self.transformer.n_class)
            blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            L.check(L.lib.focr_strokenet_prepare(table, self.transformer.n_class, blob.data_ptr(), nbytes, L.cur_stream()),
                    "strokenet_prepare")
            self._prepared = blob
        return self._prepared

    def _workspace(self, B: int, T: int, dev: torch.device) -> torch.Tensor:
        key = (B, T, dev)
        if self._ws_key != key:
            need = L.lib.focr_focus_loss_workspace_bytes(B, T)
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self._ws_key = key
        return self._ws

    def _run(self, sr: torch.Tensor, hr: torch.Tensor, text_input: torch.Tensor, lam: float, gscale: float,
             d_sr: Optional[torch.Tensor] = None, maps: bool = False):
        if not (sr.is_cuda and hr.is_cuda):
            raise L.FocrError("focr losses run on CUDA tensors only (no CPU fallback)")
        dev = sr.device
        sr_c, hr_c = sr.detach().contiguous().float(), hr.detach().contiguous().float()
        B, T = text_input.shape
        assert sr_c.shape == (B, 3, 32, 128) and hr_c.shape == sr_c.shape, (sr_c.shape, hr_c.shape)
        blob = self._prepare(dev)
        ws = self._workspace(B, T, dev)
        if d_sr is None:
            d_sr = torch.empty_like(sr_c)
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        text_input = text_input.to(dev).contiguous()
        mh = torch.empty(B, 16, T, 256, dtype=torch.float32, device=dev) if maps else None
        ms = torch.empty(B, 16, T, 256, dtype=torch.float32, device=dev) if maps else None
        L.check(L.lib.focr_focus_loss(blob.data_ptr(), blob.numel(), self.transformer.n_class, sr_c.data_ptr(),
                                      hr_c.data_ptr(), text_input.data_ptr(), B, T, float(lam), float(gscale),
                                      d_sr.data_ptr(), losses.data_ptr(), L.ptr(mh), L.ptr(ms), ws.data_ptr(), ws.numel(),
                                      L.cur_stream()), "focus_loss")
        if maps:
            return losses, d_sr, mh, ms
        return losses, d_sr


class StrokeFocusLoss(_FocusBase):
    variant = "tg"

    def __init__(self, args, decomposition: Optional[Dict[str, str]] = None,
                 transformer_state_dict: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__()
        self.args = args
        self.english_stroke_alphabet = TG_ALPHABET
        self.english_stroke_dict = {c: i for i, c in enumerate(self.english_stroke_alphabet)}
        if decomposition is None:                                    # stroke_focus_loss.py:32-38
            decomposition = {}
            with open("./dataset/mydata/english_decomposition.txt", "r") as f:
                for line in f.readlines():
                    character, sequence = line.strip().split()
                    decomposition[character] = sequence
        self.dic = decomposition
        self.build_up_transformer(transformer_state_dict)

    def build_up_transformer(self, transformer_state_dict=None):   # stroke_focus_loss.py:42-47
        self._init_engine(transformer_state_dict, "./dataset/mydata/pretrain_transformer_stroke_decomposition.pth")

    def label_stroke_encoder(self, label: Sequence[str], device=None):
        """stroke_focus_loss.py:49-80 -> (length (B,), right-shifted stroke digits (B,Tmax), text_gt (sum len,)), int64"""
        seqs = ["".join(self.dic[c] for c in one if c in self.dic) + "0" for one in label]
        length = [len(s) for s in seqs]
        tmax = max(length)
        inp = torch.zeros(len(seqs), tmax, dtype=torch.long)
        for i, s in enumerate(seqs):
            for j in range(length[i] - 1):
                inp[i, j + 1] = self.english_stroke_dict[s[j]]
        gt = torch.tensor([self.english_stroke_dict[c] for s in seqs for c in s], dtype=torch.long)
        length_t = torch.tensor(length, dtype=torch.long)
        if device is not None:
            length_t, inp, gt = length_t.to(device), inp.to(device), gt.to(device)
        return length_t, inp, gt

    def forward(self, sr_img, hr_img, label):
        if not self.args.text_focus:                                 # stroke_focus_loss.py:118-122
            mse = _MseFn.apply(sr_img, hr_img)
            return mse, mse, -1, -1
        _, text_input, _ = self.label_stroke_encoder(label, sr_img.device)
        loss, mse, att = _FocusFn.apply(sr_img, hr_img, self, text_input, float(self.args.stroke_lambda))
        return loss, mse, att, -1

    def loss_and_grad(self, sr_img, hr_img, label, gscale: float, d_sr: torch.Tensor):
        """fused-trainer entry: losses (device, [loss, mse, attention]) and d_sr = gscale * dloss/dsr written in place"""
        if not self.args.text_focus:
            raise RuntimeError("loss_and_grad is the text_focus path; the MSE-only step uses focr_mse_loss_grad")
        _, text_input, _ = self.label_stroke_encoder(label, sr_img.device)
        losses, _ = self._run(sr_img, hr_img, text_input, float(self.args.stroke_lambda), gscale, d_sr=d_sr)
        return losses
