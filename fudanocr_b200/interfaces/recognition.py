"""Evaluation-side helpers of the reference's TextSR loop on the focr engine:
``parse_crnn_data`` (interfaces/base.py:319-325), ``get_crnn_pred`` (interfaces/super_resolution.py:143-158) and
``strLabelConverter.decode`` (utils/utils_crnn.py:54-89).  The index arithmetic (argmax, repeat/blank collapse)
runs on the GPU and returns INT32 arrays; only the final index -> character join happens on the host."""
from __future__ import annotations

from typing import List, Tuple

import torch

from .. import _lib as L

ALPHABET = "-0123456789abcdefghijklmnopqrstuvwxyz"  # index 0 = CTC blank


def parse_crnn_data(imgs_input: torch.Tensor) -> torch.Tensor:
    x = imgs_input[:, :3].detach().contiguous().float()
    if not x.is_cuda or tuple(x.shape[1:]) != (3, 32, 128):
        raise L.FocrError("parse_crnn_data: expected a CUDA (B,3,32,128) tensor")
    out = torch.empty(x.shape[0], 1, 32, 100, dtype=torch.float32, device=x.device)
    L.check(L.lib.focr_bicubic_gray_32x100(x.data_ptr(), out.data_ptr(), x.shape[0], L.cur_stream()))
    return out


def ctc_greedy_decode(logits_tbc: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """logits (T,B,C) -> (path (B,T), labels (B,T) padded with -1, lengths (B)), all int32 on the device."""
    x = logits_tbc.detach().contiguous().float()
    if not x.is_cuda:
        raise L.FocrError("ctc_greedy_decode runs on CUDA only")
    T, B, Cc = x.shape
    path = torch.empty(B, T, dtype=torch.int32, device=x.device)
    out = torch.empty(B, T, dtype=torch.int32, device=x.device)
    ln = torch.empty(B, dtype=torch.int32, device=x.device)
    L.check(L.lib.focr_ctc_greedy_decode(x.data_ptr(), T, B, Cc, path.data_ptr(), out.data_ptr(), ln.data_ptr(),
                                         L.cur_stream()))
    return path, out, ln


def get_crnn_pred(outputs_btc: torch.Tensor) -> List[str]:
    """same call shape as the reference: outputs = crnn_output.permute(1, 0, 2) -> list of strings"""
    _, out, ln = ctc_greedy_decode(outputs_btc.permute(1, 0, 2))
    out, ln = out.cpu().tolist(), ln.cpu().tolist()
    return ["".join(ALPHABET[i] for i in row[:n]) for row, n in zip(out, ln)]


def str_filt(str_: str, voc_type: str = "lower") -> str:
    """utils/util.py:12-24 (the comparison key of the accuracy count, super_resolution.py:204)"""
    import string
    keep = {"digit": string.digits, "lower": string.digits + string.ascii_lowercase,
            "upper": string.digits + string.ascii_letters, "all": string.digits + string.ascii_letters + string.punctuation}
    if voc_type == "lower":
        str_ = str_.lower()
    return "".join(c for c in str_ if c in keep[voc_type])


def evaluate_batch(model, recognizer, images_lr: torch.Tensor, images_hr: torch.Tensor, label_strs) -> dict:
    """One iteration of the reference's validation loop (TextSR.eval, interfaces/super_resolution.py:178-207) on the engine:
    SR forward (eval mode) -> PSNR / SSIM -> bicubic + gray -> CRNN -> greedy CTC decode -> exact-match count.
    Everything up to the decoded index arrays stays on the device; one D2H copy (indices + the two metric scalars)."""
    from ..utils.ssim_psnr import psnr_ssim
    with torch.no_grad():
        images_sr = model(images_lr)
        psnr, ssim_avg = psnr_ssim(images_sr, images_hr)
        logits = recognizer(parse_crnn_data(images_sr[:, :3]))            # (26, B, 37)
        pred = get_crnn_pred(logits.permute(1, 0, 2).contiguous())
    n_correct = sum(1 for p, t in zip(pred, label_strs) if p == str_filt(t, "lower"))
    return {"images_sr": images_sr, "psnr": psnr, "ssim": ssim_avg, "pred": pred, "n_correct": n_correct}
