"""CTC loss on the focr engine (csrc/ctc.cu) - the "CTC forward-backward" half of the north star's
``loss/ (stroke-focus + CTC)``.  The reference repo never calls a CTC loss (its CRNN is a frozen evaluator,
scene-text-telescope/interfaces/super_resolution.py:143-158), so the surface mirrors the framework class a training script
for that CRNN would use: ``torch.nn.CTCLoss(blank, reduction, zero_infinity)`` called as
``criterion(log_probs_or_logits (T,B,C), targets, input_lengths, target_lengths)``.

One kernel launch computes log-softmax, the alpha/beta lattices, the per-sample negative log-likelihood AND the gradient
with respect to the input in the forward call (the backward of the autograd node only rescales it).  Because the gradient
formula is that of the logits (Graves eq. 16), the module may be fed raw CRNN logits or ``log_softmax`` outputs alike:
log-softmax is idempotent and its backward passes a zero-sum gradient through unchanged, which is exactly how
``F.ctc_loss`` behaves.  CUDA tensors only; no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Union

import torch

from .. import _lib as L

__all__ = ["CTCLoss", "ctc_loss"]

_RED = {"none": 0, "mean": 1, "sum": 2}


def _pad_targets(targets: torch.Tensor, target_lengths: torch.Tensor) -> torch.Tensor:
    """1-D concatenated targets (torch's second accepted form) -> (B, S_max) padded"""
    tl = target_lengths.tolist()
    smax = max(max(tl, default=0), 1)
    out = torch.zeros(len(tl), smax, dtype=torch.long)
    src = targets.cpu()
    o = 0
    for b, n in enumerate(tl):
        out[b, :n] = src[o:o + n]
        o += n
    return out


class _CTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, blank, reduction, zero_infinity):
        T, B, Cn = logits.shape
        dev = logits.device
        x = logits.detach().contiguous().float()
        S_max = targets.shape[1]
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(x)
        ws = torch.empty(L.lib.focr_ctc_loss_workspace_bytes(T, B, S_max), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib.focr_ctc_loss(x.data_ptr(), T, B, Cn, targets.data_ptr(), S_max, input_lengths.data_ptr(),
                                        target_lengths.data_ptr(), blank, reduction, int(zero_infinity), 1.0, nll.data_ptr(),
                                        loss.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(), L.cur_stream()), "ctc_loss")
        if L.status_checks():   # F.ctc_loss raises on out-of-range lengths / labels; the kernel reports them in a status word
            import ctypes as C
            st = C.c_int(0)
            L.check(L.lib.focr_ctc_loss_status(ws.data_ptr(), T, B, S_max, C.byref(st), L.cur_stream()), "ctc_loss_status")
            if st.value == 1:
                raise ValueError("ctc_loss: an input / target length is out of range")
            if st.value == 2:
                raise ValueError(f"ctc_loss: a target label lies outside [0, {Cn})")
        ctx.save_for_backward(grad)
        ctx.reduction = reduction
        ctx.in_dtype = logits.dtype
        return nll if reduction == 0 else loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        d = grad * (g.view(1, -1, 1) if ctx.reduction == 0 else g)
        return d.to(ctx.in_dtype), None, None, None, None, None, None


def ctc_loss(log_probs: torch.Tensor, targets: torch.Tensor, input_lengths: Union[torch.Tensor, Sequence[int]],
             target_lengths: Union[torch.Tensor, Sequence[int]], blank: int = 0, reduction: str = "mean",
             zero_infinity: bool = False) -> torch.Tensor:
    """same call as torch.nn.functional.ctc_loss; `log_probs` (T, B, C) may be raw logits (see module docstring)"""
    if not log_probs.is_cuda:
        raise L.FocrError("focr CTC loss runs on CUDA tensors only (no CPU fallback)")
    if log_probs.dim() != 3:
        raise ValueError(f"ctc_loss expects (T, B, C) scores, got {tuple(log_probs.shape)}")
    if reduction not in _RED:
        raise ValueError(f"reduction must be none / mean / sum, got {reduction!r}")
    dev = log_probs.device
    il = torch.as_tensor(input_lengths, dtype=torch.long)
    tl = torch.as_tensor(target_lengths, dtype=torch.long)
    B = log_probs.shape[1]
    if il.numel() != B or tl.numel() != B:
        raise ValueError("input_lengths / target_lengths must have one entry per batch element")
    if not il.is_cuda and (int(il.min()) < 0 or int(il.max()) > log_probs.shape[0]):
        raise ValueError(f"input_lengths must lie in [0, {log_probs.shape[0]}]")
    if not tl.is_cuda and int(tl.min()) < 0:
        raise ValueError("target_lengths must be non-negative")
    if targets.dim() == 1:
        targets = _pad_targets(targets, tl.cpu())
    elif targets.dim() != 2 or targets.shape[0] != B:
        raise ValueError(f"targets must be (B, S) or 1-D concatenated, got {tuple(targets.shape)}")
    if targets.shape[1] == 0:
        targets = torch.zeros(B, 1, dtype=torch.long)
    targets = targets.to(dev, torch.long).contiguous()
    return _CTCFn.apply(log_probs, targets, il.to(dev).contiguous(), tl.to(dev).contiguous(), int(blank), _RED[reduction],
                        bool(zero_infinity))


class CTCLoss(torch.nn.Module):
    """drop-in for torch.nn.CTCLoss(blank=0, reduction='mean', zero_infinity=False)"""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        if reduction not in _RED:
            raise ValueError(f"reduction must be none / mean / sum, got {reduction!r}")
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction, self.zero_infinity)
