"""Weight table of the confusion-weighted cross entropy — host-side half of ``loss.weight_ce_loss`` of scene-text-telescope
(scene-text-telescope/loss/weight_ce_loss.py:10-33).  The loss itself (``weight_cross_entropy``, :36-45) runs inside
``focr_text_focus_loss`` (wce_kernel, csrc/focus.cu); calling it from Python is not a product path."""
from __future__ import annotations

import pickle

import numpy as np
import torch

standard_alphebet = "-0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"   # (sic) weight_ce_loss.py:5


def confuse_weight_table(data: np.ndarray) -> torch.Tensor:
    """62x62 confusion counts (digits, upper, lower) -> fp32 (37,37) weights: reciprocal counts (1 where the count is 0),
    an all-ones row/column for the end symbol, and for every lower-case column the max with its upper-case twin"""
    number, upper, lower = data[:10], data[10:36], data[36:]
    re = np.concatenate((np.ones((1, 62)), number, lower, upper), axis=0)
    re = np.concatenate((np.ones((63, 1)), re), axis=1)
    with np.errstate(divide="ignore"):
        re = 1 / re
    re[re == np.inf] = 1
    t = torch.tensor(re, dtype=torch.float32)
    low = "abcdefghijklmnopqrstuvwxyz"
    for i in range(63):
        for j in range(63):
            if i != j and standard_alphebet[j] in low:
                t[i][j] = max(t[i][j], t[i][j + 26])
    return t[:37, :37].contiguous()


def load_confuse_matrix(path: str = "./dataset/mydata/confuse.pkl") -> torch.Tensor:
    with open(path, "rb") as f:
        data = pickle.load(f)
    return confuse_weight_table(np.asarray(data, dtype=np.float64))


def weight_cross_entropy(pred, gt):
    raise RuntimeError("weight_cross_entropy is fused into focr_text_focus_loss (wce_kernel); use TextFocusLoss")
