"""ORACLE (test infrastructure only).  numpy restatement of the image resize the reference's collate functions apply to every
crop: `img.resize(size, Image.BICUBIC)` followed by `transforms.ToTensor()` (scene-text-telescope/dataset/dataset.py:136-152
resizeNormalize, :231-270 alignCollate_syn / alignCollate_real).  The arithmetic lives in a third-party dependency, Pillow
(src/libImaging/Resample.c; the reference pins Pillow==6.1.0 in requirement.txt, this container has 12.2.0 - the 8-bit
resampling path is the same algorithm): for each axis, per output position, a window of `bicubic(a = -0.5)` weights whose
support is scaled by the down-sampling factor (antialiasing), normalised in double precision and quantised to 22-bit fixed
point; a horizontal pass into a uint8 intermediate, then a vertical pass, each output = clip8((2^21 + sum k_i p_i) >> 22).
Pinned against PIL itself (tests/golden/resize.npz and, where PIL is importable, directly) - bit-exact uint8."""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box -> (bounds (out,2) int, kk (out,ksize) int)"""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """img (H, W, C) uint8 -> (oh, ow, C) uint8 with size = (ow, oh) as PIL's Image.resize(size, Image.BICUBIC)"""
    ow, oh = size
    h, w, _ = img.shape
    cur = img.astype(np.int64)
    if w != ow:
        bounds, kk = precompute_coeffs(w, ow)
        out = np.zeros((h, ow, img.shape[2]), np.int64)
        for xx in range(ow):
            x0, n = bounds[xx]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(cur[:, x0:x0 + n, :], kk[xx, :n], axes=([1], [0]))
            out[:, xx, :] = _clip8(acc)
        cur = out
    if h != oh:
        bounds, kk = precompute_coeffs(h, oh)
        out = np.zeros((oh, cur.shape[1], img.shape[2]), np.int64)
        for yy in range(oh):
            y0, n = bounds[yy]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[yy, :n], cur[y0:y0 + n], axes=([0], [0]))
            out[yy] = _clip8(acc)
        cur = out
    return cur.astype(np.uint8)


def resize_normalize(img: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """resizeNormalize.__call__ (dataset.py:143-152, mask=False): resize + ToTensor -> float32 (C, oh, ow) in [0, 1]"""
    r = resize_bicubic_u8(img, size)
    return (r.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)


def synth_crops(n: int, seed: int):
    """deterministic ragged batch of uint8 RGB crops with TextZoom-like sizes (up- and down-sampling on both axes)"""
    rs = np.random.RandomState(seed)
    out = []
    for i in range(n):
        h = int(rs.randint(9, 70))
        w = int(rs.randint(20, 400))
        base = rs.randint(0, 256, size=(h, w, 3)).astype(np.float64)
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        smooth = 127 + 100 * np.sin(xx / (3.0 + i % 5))[..., None] * np.cos(yy / 2.5)[..., None]
        out.append(np.clip(0.5 * base + 0.5 * smooth, 0, 255).astype(np.uint8))
    out.append(rs.randint(0, 256, size=(32, 128, 3)).astype(np.uint8))   # already HR-sized: both passes skipped for (128, 32)
    out.append(rs.randint(0, 256, size=(16, 200, 3)).astype(np.uint8))   # height already LR-sized
    return out
