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


def thresh7(p: float) -> int:
    """attention.cu compares 7-bit lanes: th7 = (thresh16 + 256) >> 9 (p quantised to 1/128)"""
    t = thresh16(p)
    th = (t + 256) >> 9
    if t and th == 0:
        th = 1
    return min(th, 127)


def attn_keep_mask(B: int, seed: int, blk: int, p: float) -> torch.Tensor:
    """(B,4,1024,1024) bool: element (b,h,q,k) has index ((b*4+h)*1024+q)*1024 + k; the hash of (index >> 2) carries
    four 7-bit lanes (low 7 bits of byte k & 3); kept iff lane >= th7 (attention.cu)."""
    idx = np.arange(B * 4 * 1024 * 1024, dtype=np.uint64)
    h = hash32(drop_key(seed, 2 * blk), idx >> np.uint64(2))
    lane = (h >> (np.uint64(8) * (idx & np.uint64(3)))) & np.uint64(0x7F)
    return torch.from_numpy((lane >= np.uint64(thresh7(p))).reshape(B, 4, 1024, 1024))


def attn_keep_scale(p: float) -> float:
    """attention.cu scales kept probabilities by 128/(128-th7)"""
    return 128.0 / (128.0 - thresh7(p))


def attn_drop_rate(p: float) -> float:
    """the rate the attention kernels actually apply (13/128 for p = 0.1)"""
    return thresh7(p) / 128.0


def ffn_keep_mask(B: int, seed: int, blk: int, p: float) -> torch.Tensor:
    """(B,1024,128) bool: element (t,n) has index t*128+n (tc_gemm.cu epilogue, 16-bit lanes)."""
    idx = np.arange(B * 1024 * 128, dtype=np.uint64)
    return torch.from_numpy(_keep(drop_key(seed, 2 * blk + 1), idx, thresh16(p)).reshape(B, 1024, 128))


def keep_scale(p: float) -> float:
    """the FFN epilogue scales kept values by 65536/(65536-thresh16) (exactly 1/(1-p_effective))"""
    return 65536.0 / (65536.0 - thresh16(p))
