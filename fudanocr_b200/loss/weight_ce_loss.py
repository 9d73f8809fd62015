"""``loss.weight_ce_loss`` of scene-text-telescope on the focr engine (scene-text-telescope/loss/weight_ce_loss.py): the
weight table built on the host (:10-33) and ``weight_cross_entropy(pred, gt)`` (:36-45) as a CUDA autograd function.  Inside
``TextFocusLoss`` the same arithmetic runs fused in ``focr_text_focus_loss`` (wce_kernel, csrc/focus.cu); this module-level
callable exists because the reference exports it.  No CPU fallback."""
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


weight_table = None   # the reference builds it at import from ./dataset/mydata/confuse.pkl (:35); here on first use


def set_weight_table(table: torch.Tensor) -> None:
    """install a (C, C) weight table (e.g. ``confuse_weight_table(counts)``) instead of reading confuse.pkl"""
    global weight_table
    weight_table = table.detach().float().contiguous()


class _WeightCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, table):
        from .. import _lib as L
        if not pred.is_cuda:
            raise L.FocrError("weight_cross_entropy runs on CUDA tensors only (no CPU fallback)")
        p = pred.detach().float().contiguous()
        g = gt.to(device=p.device, dtype=torch.long).contiguous()
        n, c = p.shape
        if table.shape != (c, c):
            raise ValueError(f"weight table {tuple(table.shape)} does not match {c} classes")
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        d_pred = torch.empty_like(p)
        status = torch.zeros(1, dtype=torch.int32, device=p.device)
        ws = torch.empty(L.lib.focr_weight_cross_entropy_workspace_bytes(n), dtype=torch.uint8, device=p.device)
        L.check(L.lib.focr_weight_cross_entropy(p.data_ptr(), g.data_ptr(), table.data_ptr(), loss.data_ptr(), d_pred.data_ptr(),
                                                status.data_ptr(), n, c, ws.data_ptr(), ws.numel(), L.cur_stream()),
                "weight_cross_entropy")
        if int(status.item()):   # the reference's weight_table[gt] raises IndexError on such labels
            raise IndexError("weight_cross_entropy: target index out of range")
        ctx.save_for_backward(d_pred)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_pred,) = ctx.saved_tensors
        return d_pred * g, None, None


def weight_cross_entropy(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """-mean_i log(w[gt_i, gt_i] e^{pred_i, gt_i} / sum_j w[gt_i, j] e^{pred_ij})  (weight_ce_loss.py:36-45)"""
    global weight_table
    if weight_table is None:
        set_weight_table(load_confuse_matrix())
    if weight_table.device != pred.device:
        weight_table = weight_table.to(pred.device)
    return _WeightCE.apply(pred, gt, weight_table)
