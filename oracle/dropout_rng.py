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


def thresh15(p: float) -> int:
    """attention.cu compares 15-bit lanes: th15 = (thresh16 + 1) >> 1"""
    return (thresh16(p) + 1) >> 1


def attn_keep_mask(B: int, seed: int, blk: int, p: float) -> torch.Tensor:
    """(B,4,1024,1024) bool: element (b,h,q,k) has index ((b*4+h)*1024+q)*1024 + k; its hash (counter index >> 1)
    carries two 15-bit lanes (bits 0-14 for even k, bits 16-30 for odd k); kept iff lane >= th15 (attention.cu)."""
    idx = np.arange(B * 4 * 1024 * 1024, dtype=np.uint64)
    h = hash32(drop_key(seed, 2 * blk), idx >> np.uint64(1))
    lane = np.where((idx & np.uint64(1)) == 0, h & np.uint64(0x7FFF), (h >> np.uint64(16)) & np.uint64(0x7FFF))
    return torch.from_numpy((lane >= np.uint64(thresh15(p))).reshape(B, 4, 1024, 1024))


def attn_keep_scale(p: float) -> float:
    """attention.cu scales kept probabilities by 32768/(32768-th15)"""
    return 32768.0 / (32768.0 - thresh15(p))


def ffn_keep_mask(B: int, seed: int, blk: int, p: float) -> torch.Tensor:
    """(B,1024,128) bool: element (t,n) has index t*128+n (tc_gemm.cu epilogue)."""
    idx = np.arange(B * 1024 * 128, dtype=np.uint64)
    return torch.from_numpy(_keep(drop_key(seed, 2 * blk + 1), idx, thresh16(p)).reshape(B, 1024, 128))


def keep_scale(p: float) -> float:
    """the kernels scale kept values by 65536/(65536-thresh16) (exactly 1/(1-p_effective))"""
    return 65536.0 / (65536.0 - thresh16(p))
