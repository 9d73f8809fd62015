"""ORACLE support — deterministic synthetic weights / inputs (test infrastructure only).

The reference ships no checkpoints and no fixtures, so parity is pinned on
seeded synthetic tensors.  numpy's legacy RandomState stream is stable across numpy versions,
hence the same (spec, seed) reproduces the same state dict in this container, on the GPU box and in
oracle/make_golden.py, which loads it into the *reference* nn.Modules.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict

import numpy as np
import torch

GOLDEN_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load_spec(name: str) -> Dict[str, list]:
    return json.loads((GOLDEN_DIR / f"{name}_spec.json").read_text())


def synth_state_dict(spec: Dict[str, list], seed: int = 1234, computed: Dict[str, torch.Tensor] = None):
    """Values are drawn per key in spec order: weights ~ N(0, 1/sqrt(fan_in)) (so activations stay
    O(1) through 5 SRBs), affine scales ~ 1 + 0.1 N, biases ~ 0.05 N, running_var ~ U(0.5, 1.5).
    `computed` supplies buffers that are functions of the architecture (the TPS matrices)."""
    rs = np.random.RandomState(seed)
    sd = {}
    for key, shape in spec.items():
        if computed and key in computed:
            sd[key] = computed[key].clone()
            continue
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros((), dtype=torch.long)
            continue
        n = int(np.prod(shape)) if shape else 1
        z = rs.standard_normal(n).astype(np.float32).reshape(shape)
        if key.endswith("stn_fc2.bias"):  # identity control points (stn_head.py:69-86) + small noise
            v = identity_ctrl_points().reshape(-1) + 0.02 * z
        elif key.endswith("stn_fc2.weight"):
            # the reference initialises this weight to ZERO (stn_head.py:85: identity warp) and Adam moves it by
            # ~1e-4 per step, so a lightly trained STN has |w| ~ 1e-3; non-zero so the gradient path is exercised
            v = 0.002 * z
        elif leaf == "running_var":
            v = (0.5 + rs.random_sample(n)).astype(np.float32).reshape(shape)
        elif leaf == "running_mean":
            v = 0.1 * z
        elif leaf == "a_2" or (leaf == "weight" and len(shape) == 1 and shape[0] > 1):  # norm scales
            v = 1.0 + 0.1 * z
        elif leaf == "weight" and len(shape) == 1:  # PReLU slope (single parameter)
            v = 0.25 + 0.05 * z
        elif leaf == "weight" and len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            v = z / np.sqrt(fan_in)
        elif leaf.startswith("weight_") or leaf.startswith("bias_"):  # GRU / LSTM
            v = 0.1 * z
        else:  # biases, b_2
            v = 0.05 * z
        sd[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def identity_ctrl_points(n: int = 20, margin: float = 0.01) -> np.ndarray:
    """stn_head.py:69-86 (init_stn): the bias that makes the TPS warp the identity."""
    k = n // 2
    xs = np.linspace(margin, 1.0 - margin, k)
    top = np.stack([xs, np.ones(k) * margin], axis=1)
    bot = np.stack([xs, np.ones(k) * (1 - margin)], axis=1)
    return np.concatenate([top, bot], axis=0).astype(np.float32)


def synth_images(B: int, seed: int = 1234):
    """LR (B,3,16,64) and HR (B,3,32,128) fp32 in [0,1] (ToTensor range, dataset/dataset.py:143-152).
    Smooth, text-crop-like content: HR is a random low-frequency field plus a few sharp "strokes" and mild
    noise; LR is its 2x2 box down-sample (the synthetic-LR recipe of dataset.py:240-254 uses bicubic).
    White noise would make the TPS resampling step chaotic under any perturbation of the control points."""
    rs = np.random.RandomState(seed + 17)
    yy, xx = np.meshgrid(np.linspace(0, 1, 32), np.linspace(0, 1, 128), indexing="ij")
    hr = np.zeros((B, 3, 32, 128), dtype=np.float64)
    for b in range(B):
        for c in range(3):
            f = np.zeros_like(yy)
            for _ in range(6):
                fx, fy = rs.uniform(0.5, 6.0), rs.uniform(0.3, 2.0)
                f += rs.uniform(0.2, 1.0) * np.cos(2 * np.pi * (fx * xx + fy * yy) + rs.uniform(0, 2 * np.pi))
            for _ in range(4):  # vertical / slanted strokes
                x0, wd, sl = rs.uniform(0.05, 0.95), rs.uniform(0.01, 0.03), rs.uniform(-0.2, 0.2)
                f += rs.uniform(1.0, 2.0) * np.exp(-((xx - x0 - sl * (yy - 0.5)) / wd) ** 2)
            f = (f - f.min()) / (f.max() - f.min() + 1e-9)
            hr[b, c] = np.clip(0.9 * f + 0.05 + 0.01 * rs.standard_normal(f.shape), 0.0, 1.0)
    lr = hr.reshape(B, 3, 16, 2, 64, 2).mean(axis=(3, 5))
    return torch.from_numpy(lr.astype(np.float32)), torch.from_numpy(hr.astype(np.float32))
