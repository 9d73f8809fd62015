"""ORACLE support — numpy twin of the device dropout RNG (fudanocr_b200/csrc/common.cuh:
drop_hash32 / drop_key).  Lets the oracle reproduce the exact keep-masks the kernels draw, so parity
can be checked with dropout ON (test infrastructure only)."""
from __future__ import annotations

import numpy as np
import torch

M32 = np.uint64(0xFFFFFFFF)


def drop_key(seed: int, stream: int) -> int:
    return (seed ^ ((stream * 0x9E3779B9 + 0x7F4A7C15) & 0xFFFFFFFF)) & 0xFFFFFFFF


def hash32(key: int, ctr: np.ndarray) -> np.ndarray:
    h = (ctr.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(key)) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & M32
    h ^= h >> np.uint64(16)
    return h


def thresh16(p: float) -> int:
    return 0 if p <= 0 else min(65535, int(p * 65536.0 + 0.5))


def _keep(key: int, elem_index: np.ndarray, th: int) -> np.ndarray:
    h = hash32(key, elem_index >> np.uint64(1))
    lane = np.where((elem_index & np.uint64(1)) == 0, h & np.uint64(0xFFFF), h >> np.uint64(16))
    return lane >= np.uint64(th)


# ---- attention dropout: bit-sliced Bernoulli(th15 / 32768) over Philox-2x32-4 words (attention.cu: keep_word32) ----
PHILOX_M = np.uint64(0xD256D193)
PHILOX_W = 0x9E3779B9


def thresh15(p: float) -> int:
    """attention.cu: th15 = (thresh16 + 1) >> 1 = round(p * 2^15) (p = 0.1 -> 3277)"""
    t = thresh16(p)
    th = (t + 1) >> 1
    if t and th == 0:
        th = 1
    return min(th, 32767)


def philox4(ctr: np.ndarray, key: int):
    """four rounds of Philox-2x32: (L, R) <- (hi(L * M) ^ k_r ^ R, lo(L * M)), k_r = key * 0x85EBCA6B + 0x1B873593 + r W"""
    L = ctr.astype(np.uint64) & M32
    R = np.full_like(L, key & 0xFFFFFFFF)
    k = (key * 0x85EBCA6B + 0x1B873593) & 0xFFFFFFFF
    for _ in range(4):
        prod = L * PHILOX_M
        L = (prod >> np.uint64(32)) ^ np.uint64(k) ^ R
        R = prod & M32
        k = (k + PHILOX_W) & 0xFFFFFFFF
    return L, R


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & M32


def attn_keep_words(rowid: np.ndarray, kw: np.ndarray, key: int, th15: int) -> np.ndarray:
    """keep word (uint32 in a uint64 array) of keys [32 kw, 32 kw + 32) of row `rowid` = (b*4+h)*1024 + q"""
    ctr0 = ((rowid.astype(np.uint64) * np.uint64(32) + kw.astype(np.uint64)) * np.uint64(8)) & M32
    w = []
    for c in range(6):
        L, R = philox4((ctr0 + np.uint64(c)) & M32, key)
        w += [L, R]
    w += [_rotl(w[0], 7), _rotl(w[1], 13), _rotl(w[2], 22)]
    r = np.zeros_like(ctr0)
    for i in range(14, -1, -1):
        r = (w[i] | r) if (th15 >> (14 - i)) & 1 else (w[i] & r)
    return (~r) & M32


def attn_keep_mask(B: int, seed: int, blk: int, p: float) -> torch.Tensor:
    """(B,4,1024,1024) bool keep mask of attention layer `blk`: key k of a 32-key group sits at bit 8 (k & 3) + (k % 32) // 4
    of the group's keep word (attention.cu)."""
    rowid = np.arange(B * 4 * 1024, dtype=np.uint64)[:, None]
    kw = np.arange(32, dtype=np.uint64)[None, :]
    words = attn_keep_words(rowid, kw, drop_key(seed, 2 * blk), thresh15(p))       # [B*4*1024, 32]
    kk = np.arange(32, dtype=np.uint64)
    sh = kk // np.uint64(4) + np.uint64(8) * (kk & np.uint64(3))
    m = (words[:, :, None] >> sh[None, None, :]) & np.uint64(1)                      # [rows, kw, kk]
    return torch.from_numpy(m.astype(bool).reshape(B, 4, 1024, 1024))


def attn_keep_scale(p: float) -> float:
    """attention.cu scales kept probabilities by 32768/(32768-th15) (exactly 1/(1-rate))"""
    return 32768.0 / (32768.0 - thresh15(p))


def attn_drop_rate(p: float) -> float:
    """the rate the attention kernels actually apply (3277/32768 = 0.100006 for p = 0.1)"""
    return thresh15(p) / 32768.0


def ffn_keep_mask(B: int, seed: int, blk: int, p: float) -> torch.Tensor:
    """(B,1024,128) bool: element (t,n) has index t*128+n (tc_gemm.cu epilogue, 16-bit lanes)."""
    idx = np.arange(B * 1024 * 128, dtype=np.uint64)
    return torch.from_numpy(_keep(drop_key(seed, 2 * blk + 1), idx, thresh16(p)).reshape(B, 1024, 128))


def keep_scale(p: float) -> float:
    """the FFN epilogue scales kept values by 65536/(65536-thresh16) (exactly 1/(1-p_effective))"""
    return 65536.0 / (65536.0 - thresh16(p))
