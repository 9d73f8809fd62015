"""ORACLE (test infrastructure only).  Restatement of scene-text-telescope/utils/ssim_psnr.py: calculate_psnr (:9-15),
gaussian / create_window (:18-28), _ssim (:31-51), SSIM.forward (:54-78).  Pinned against the unmodified reference module by
oracle/make_golden_metrics.py (tests/golden/metrics.pt)."""
from math import exp

import torch
import torch.nn.functional as F


def calculate_psnr(img1, img2):
    mse = ((img1[:, :3] * 255 - img2[:, :3] * 255) ** 2).mean()
    if mse == 0:
        return torch.tensor(float("inf"))
    return 20 * torch.log10(255.0 / torch.sqrt(mse))


def create_window(window_size: int, channel: int):
    g = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size: int = 11, size_average: bool = True):
    img1, img2 = img1[:, :3], img2[:, :3]
    c = img1.shape[1]
    w = create_window(window_size, c).to(img1)
    pad = window_size // 2
    mu1, mu2 = F.conv2d(img1, w, padding=pad, groups=c), F.conv2d(img2, w, padding=pad, groups=c)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=pad, groups=c) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=pad, groups=c) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=pad, groups=c) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)


def synth_pair(B: int, seed: int):
    """deterministic (sr, hr) test batches with a 4th (mask) channel the metrics must ignore"""
    from oracle import synth
    lr, hr = synth.synth_images(B, seed=seed)
    sr = F.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False)
    sr = (sr + 0.03 * torch.randn(sr.shape, generator=torch.Generator().manual_seed(seed))).clamp(0, 1)
    sr4 = torch.cat([sr, torch.rand(B, 1, 32, 128, generator=torch.Generator().manual_seed(1))], 1)
    hr4 = torch.cat([hr, torch.rand(B, 1, 32, 128, generator=torch.Generator().manual_seed(2))], 1)
    return sr4.contiguous(), hr4.contiguous()
