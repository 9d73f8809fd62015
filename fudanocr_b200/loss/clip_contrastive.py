"""Contrastive head of the CCR-CLIP pre-training stage (image-ids-CTR/CCR-CLIP) on the focr engine.

Reference: ``CLIP.forward`` normalises both feature sets and returns ``logit_scale.exp()`` (model.py:209-222); the training loop
forms ``logits_per_image = logit_scale * image_features @ text_features.t()``, its transpose, and
``(CE(logits_per_image, gt) + CE(logits_per_text, gt)) / 2`` with ``gt[i] = ''.join(label).index(label[i])`` (main.py:98-110).
``clip_contrastive_loss`` is that arithmetic - normalisation included - as ONE fused call with gradients for both towers'
(un-normalised) outputs and the ``logit_scale`` parameter.  The towers (ResNet-50, 12-layer text transformer) are outside this
repository's scope; this module is what a data-parallel run adds to them:

* one process per GPU; every rank all-gathers the (B_local, D) features of both towers (NCCL all-gather over NVLink, 2 x B x D x 4
  bytes = 2 MB at the reference's B = 128, D = 2048), evaluates the global B x B problem redundantly and keeps the gradient rows of
  its own shard - no second exchange; the towers' parameter gradients are then SUMMED over ranks (``grad_reduce='sum'``) or, under a
  wrapper that averages them, pre-scaled by the world size (``grad_reduce='mean'``); ``logit_scale``'s gradient, identical on every
  rank, is scaled the other way so that either reduction yields the single-process value.

CUDA tensors only - there is no CPU path."""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist


def ground_truth_from_labels(labels: Sequence[str], device=None) -> torch.Tensor:
    """main.py:101-105: the target of sample i is the position of its label in the concatenation of the batch's labels (for the
    single-character labels of the font images: the first sample showing the same character)"""
    label_str = "".join(labels)
    gt = torch.tensor([label_str.index(lab) for lab in labels], dtype=torch.long)
    return gt if device is None else gt.to(device)


def _fused(image: torch.Tensor, text: torch.Tensor, logit_scale: torch.Tensor, gt: torch.Tensor, want_grad: bool):
    """-> (loss (1,), d_image, d_text, d_logit_scale (1,)) through focr_clip_contrastive_loss (module-level so tests can swap it)"""
    from .. import _lib as L
    if not image.is_cuda:
        raise L.FocrError("clip_contrastive_loss runs on CUDA tensors only (no CPU fallback)")
    B, D = image.shape
    dev = image.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d_img = torch.empty_like(image) if want_grad else None
    d_txt = torch.empty_like(text) if want_grad else None
    d_ls = torch.empty(1, dtype=torch.float32, device=dev) if want_grad else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(L.lib.focr_clip_contrastive_workspace_bytes(B, D), dtype=torch.uint8, device=dev)
    L.check(L.lib.focr_clip_contrastive_loss(image.data_ptr(), text.data_ptr(), logit_scale.data_ptr(), gt.data_ptr(), B, D,
                                             loss.data_ptr(), L.ptr(d_img), L.ptr(d_txt), L.ptr(d_ls), status.data_ptr(),
                                             ws.data_ptr(), ws.numel(), L.cur_stream()), "clip_contrastive_loss")
    if L.status_checks() and int(status.item()):
        raise IndexError("clip_contrastive_loss: target index outside the (global) batch")
    return loss, d_img, d_txt, d_ls


class _ClipContrastive(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, gt, group, grad_reduce):
        img = image_features.detach().float().contiguous()
        txt = text_features.detach().float().contiguous()
        ls = logit_scale.detach().float().reshape(1).contiguous()
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        b_local = img.shape[0]
        row0 = 0
        if world > 1:   # global batch = concatenation over ranks, in rank order (equal shards)
            g_img = torch.empty(world * b_local, img.shape[1], dtype=img.dtype, device=img.device)
            g_txt = torch.empty_like(g_img)
            dist.all_gather_into_tensor(g_img, img, group=group)
            dist.all_gather_into_tensor(g_txt, txt, group=group)
            row0 = dist.get_rank(group) * b_local
            img, txt = g_img, g_txt
        if gt.shape[0] != img.shape[0]:
            raise ValueError(f"ground truth has {gt.shape[0]} entries for a global batch of {img.shape[0]}")
        want = any(ctx.needs_input_grad[:3])
        loss, d_img, d_txt, d_ls = _fused(img, txt, ls, gt.to(device=img.device, dtype=torch.long).contiguous(), want)
        if want:
            f_scale = float(world) if grad_reduce == "mean" else 1.0      # towers: local rows, summed over ranks downstream
            s_scale = 1.0 if grad_reduce == "mean" else 1.0 / world       # logit_scale: the same value on every rank
            ctx.save_for_backward(d_img[row0:row0 + b_local] * f_scale, d_txt[row0:row0 + b_local] * f_scale, d_ls * s_scale)
        ctx.ls_shape = logit_scale.shape
        ctx.dtypes = (image_features.dtype, text_features.dtype, logit_scale.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        d_img, d_txt, d_ls = ctx.saved_tensors
        t0, t1, t2 = ctx.dtypes
        return (d_img * g).to(t0), (d_txt * g).to(t1), (d_ls * g).reshape(ctx.ls_shape).to(t2), None, None, None


def clip_contrastive_loss(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor,
                          ground_truth: torch.Tensor, process_group=None, grad_reduce: str = "sum") -> torch.Tensor:
    """image_features / text_features: (B_local, D) un-normalised tower outputs of this rank; logit_scale: the model's log-scale
    parameter (``CLIP.logit_scale``, model.py:179); ground_truth: int64 targets of the GLOBAL batch (``ground_truth_from_labels`` of
    the labels of all ranks in rank order).  Returns the scalar loss of main.py:106 (identical on every rank)."""
    if grad_reduce not in ("sum", "mean"):
        raise ValueError("grad_reduce must be 'sum' or 'mean'")
    if image_features.dim() != 2 or image_features.shape != text_features.shape:
        raise ValueError(f"features must be (B, D) and of equal shape, got {tuple(image_features.shape)} / {tuple(text_features.shape)}")
    return _ClipContrastive.apply(image_features, text_features, logit_scale, ground_truth, process_group, grad_reduce)


class ClipContrastiveLoss(torch.nn.Module):
    """module form: ``crit(image_features, text_features, model.logit_scale, labels_of_all_ranks)``"""

    def __init__(self, process_group=None, grad_reduce: str = "sum"):
        super().__init__()
        self.process_group, self.grad_reduce = process_group, grad_reduce

    def forward(self, image_features, text_features, logit_scale, labels: Optional[Sequence[str]] = None, ground_truth=None):
        if ground_truth is None:
            if labels is None:
                raise ValueError("pass the global batch's labels or its ground_truth tensor")
            ground_truth = ground_truth_from_labels(labels, image_features.device)
        return clip_contrastive_loss(image_features, text_features, logit_scale, ground_truth, self.process_group, self.grad_reduce)
