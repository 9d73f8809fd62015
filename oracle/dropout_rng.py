"""ORACLE support — numpy twin of the device dropout RNG (fudanocr_b200/csrc/common.cuh:
drop_rand2 / drop_key).  Lets the oracle reproduce the exact keep-masks the kernels draw, so parity
can be checked with dropout ON (test infrastructure only)."""
from __future__ import annotations

import numpy as np
import torch

M32 = np.uint64(0xFFFFFFFF)


def drop_key(seed: int, stream: int) -> int:
    return (seed ^ ((stream * 0x9E3779B9 + 0x7F4A7C15) & 0xFFFFFFFF)) & 0xFFFFFFFF


def rand2(key: int, ctr: np.ndarray):
    """(r0, r1): the 32-bit uniforms of elements 2*ctr and 2*ctr+1 (drop_rand2 in common.cuh)."""
    m = ((ctr.astype(np.uint64) ^ np.uint64(key)) & M32) * np.uint64(0xD2511F53)
    h = ((m >> np.uint64(32)) ^ m) & M32
    m = h * np.uint64(0xCD9E8D57)
    r0 = ((m >> np.uint64(32)) ^ m) & M32
    r1 = (r0 * np.uint64(0x85EBCA6B) + np.uint64(0xC2B2AE35)) & M32
    return r0, r1


def thresh32(p: float) -> int:
    return 0 if p <= 0 else min(4294967295, int(float(np.float32(p)) * 4294967296.0 + 0.5))


def _keep(key: int, elem_index: np.ndarray, th: int) -> np.ndarray:
    r0, r1 = rand2(key, elem_index >> np.uint64(1))
    return np.where((elem_index & np.uint64(1)) == 0, r0, r1) >= np.uint64(th)


def attn_keep_mask(B: int, seed: int, blk: int, p: float) -> torch.Tensor:
    """(B,4,1024,1024) bool: element (b,h,q,k) has index ((b*4+h)*1024+q)*1024 + k (attention.cu)."""
    idx = np.arange(B * 4 * 1024 * 1024, dtype=np.uint64)
    return torch.from_numpy(_keep(drop_key(seed, 2 * blk), idx, thresh32(p)).reshape(B, 4, 1024, 1024))


def ffn_keep_mask(B: int, seed: int, blk: int, p: float) -> torch.Tensor:
    """(B,1024,128) bool: element (t,n) has index t*128+n (tc_gemm.cu epilogue)."""
    idx = np.arange(B * 1024 * 128, dtype=np.uint64)
    return torch.from_numpy(_keep(drop_key(seed, 2 * blk + 1), idx, thresh32(p)).reshape(B, 1024, 128))


def keep_scale(p: float) -> float:
    """the kernels scale kept values by 2^32/(2^32-thresh32) (exactly 1/(1-p_effective))"""
    return 4294967296.0 / (4294967296.0 - thresh32(p))
