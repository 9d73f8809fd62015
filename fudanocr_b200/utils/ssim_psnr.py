"""PSNR / SSIM on the focr engine — drop-in for ``utils.ssim_psnr`` of scene-text-telescope / text-gestalt
(scene-text-telescope/utils/ssim_psnr.py: ``calculate_psnr`` :9-15, ``SSIM`` :54-78, ``ssim`` :81-89), the metrics the
validation loop computes per batch (interfaces/super_resolution.py:191-192, base.py:62-63).  One fused kernel
(``focr_psnr_ssim``) instead of ~25 launches and ten full-size temporaries; CUDA tensors only (no CPU fallback)."""
from __future__ import annotations

from math import exp

import torch

from .. import _lib as L

__all__ = ["calculate_psnr", "SSIM", "ssim", "psnr_ssim"]


def gaussian(window_size: int, sigma: float) -> torch.Tensor:           # ssim_psnr.py:18-20
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size: int = 11) -> torch.Tensor:                # ssim_psnr.py:23-28 (one channel: all are equal)
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    return w1.mm(w1.t()).float().contiguous()


_WINDOWS = {}


def psnr_ssim(img1: torch.Tensor, img2: torch.Tensor, per_image: bool = False):
    """(psnr, ssim_mean[, ssim_per_image]) of fp32 NCHW (B, >=3, 32, 128) batches in [0,1]; device tensors, no sync"""
    if not (img1.is_cuda and img2.is_cuda):
        raise L.FocrError("focr metrics run on CUDA tensors only (no CPU fallback)")
    if img1.shape != img2.shape or img1.dim() != 4 or img1.shape[1] < 3 or tuple(img1.shape[2:]) != (32, 128):
        raise ValueError(f"psnr_ssim expects two (B,>=3,32,128) batches, got {tuple(img1.shape)} / {tuple(img2.shape)}")
    a, b = img1.detach().contiguous().float(), img2.detach().contiguous().float()
    dev = a.device
    win = _WINDOWS.get(dev)
    if win is None:
        win = _WINDOWS[dev] = create_window(11).to(dev)
    B = a.shape[0]
    out = torch.empty(2, dtype=torch.float32, device=dev)
    per = torch.empty(B, dtype=torch.float32, device=dev) if per_image else None
    ws = torch.empty(L.lib.focr_psnr_ssim_workspace_bytes(B), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib.focr_psnr_ssim(a.data_ptr(), b.data_ptr(), B, a.shape[1], win.data_ptr(), out.data_ptr(), L.ptr(per),
                                     ws.data_ptr(), ws.numel(), L.cur_stream()), "psnr_ssim")
    return (out[0], out[1], per) if per_image else (out[0], out[1])


def calculate_psnr(img1, img2):
    return psnr_ssim(img1, img2)[0]


class SSIM(torch.nn.Module):
    def __init__(self, window_size: int = 11, size_average: bool = True):
        super().__init__()
        if window_size != 11:
            raise NotImplementedError("focr SSIM: window_size 11 (the reference's only use)")
        self.window_size, self.size_average = window_size, size_average

    def forward(self, img1, img2):
        if self.size_average:
            return psnr_ssim(img1, img2)[1]
        return psnr_ssim(img1, img2, per_image=True)[2]


def ssim(img1, img2, window_size: int = 11, size_average: bool = True):
    return SSIM(window_size, size_average)(img1, img2)
