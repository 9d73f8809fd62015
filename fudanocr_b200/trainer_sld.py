"""Fused training step of the stroke-level-decomposition recogniser on the focr engine - the B200 form of
stroke-level-decomposition/train.py:63-77:

    model.train(); optimizer.zero_grad(); result = model(image, length, text_input)
    loss = CrossEntropyLoss(result['pred'], text_gt); loss.backward(); Adadelta(lr 1, rho 0.9).step()

Forward, packed cross entropy (value + logits gradient in one kernel), backward into ONE flat fp32 gradient buffer,
(multi-GPU: a single NCCL all-reduce of that buffer over NVLink - 287 MB for the 71.7 M parameters),
then one multi-tensor Adadelta launch.  Parameters keep living in the module's nn.Parameters (re-pointed into a flat buffer)
so ``state_dict`` and checkpoints are unchanged.  image-ids-CTR's optimiser differs only by ``weight_decay`` (train.py there)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .model import recog_ops as ops
from .model.transformer import Transformer

_CHUNK = 65536   # elements per optimizer CTA: ~1100 CTAs for 71.7 M parameters


class _FlatAdadeltaTrainer:
    """parameters, gradients and the two Adadelta states in flat fp32 buffers; autograd accumulates straight into the gradient
    buffer; `skip` names the parameters the reference forward never touches (their .grad stays None there and its optimiser
    skips them - which matters once weight decay is on)"""

    def __init__(self, model, skip, lr, rho, eps, weight_decay, process_group):
        self.model = model
        self.lr, self.rho, self.eps, self.weight_decay = lr, rho, eps, weight_decay
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.names = [k for k, _ in model.named_parameters() if not any(s in k for s in skip)]
        params = dict(model.named_parameters())
        plist = [params[k] for k in self.names]
        dev = plist[0].device
        ops.require_cuda(plist[0])
        offs, tot = [], 0
        for p in plist:
            offs.append(tot)
            tot += (p.numel() + 3) // 4 * 4
        self.flat_p = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.flat_sq = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.flat_acc = torch.zeros(tot, dtype=torch.float32, device=dev)
        recs = []
        for p, o in zip(plist, offs):
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + n].view_as(p)
            p.grad = self.flat_g[o:o + n].view_as(p)     # autograd accumulates in place into the flat buffer
            k = 0
            while k < n:
                ln = min(_CHUNK, n - k)
                recs.append([self.flat_p.data_ptr() + 4 * (o + k), self.flat_g.data_ptr() + 4 * (o + k),
                             self.flat_sq.data_ptr() + 4 * (o + k), self.flat_acc.data_ptr() + 4 * (o + k), ln])
                k += ln
        self.grad_views = [p.grad for p in plist]
        self.plist = plist
        self.chunks = torch.tensor(recs, dtype=torch.int64, device=dev)
        self.loss: Optional[torch.Tensor] = None
        if self.world > 1:   # start from rank 0's weights on every rank
            dist.broadcast(self.flat_p, src=0, group=self.pg)
            rank = dist.get_rank(self.pg)
            if hasattr(model, "_seed"):   # replicas draw independent dropout masks, as nn.DataParallel's do
                model._seed = (model._seed ^ (rank * 0x9E3779B1)) & 0x7FFFFFFF

    def _begin(self):
        self.model.train()
        self.flat_g.zero_()                                # optimizer.zero_grad()
        for p, g in zip(self.plist, self.grad_views):      # keep .grad pointing into the flat buffer
            if p.grad is not g:
                p.grad = g

    def _finish(self, loss, lr=None):
        loss.backward()
        if self.world > 1:
            dist.all_reduce(self.flat_g, group=self.pg)    # sum; the mean is taken inside the optimizer kernel
        ops.adadelta_step(self.chunks, self.chunks.shape[0], 1.0 / self.world, self.lr if lr is None else lr, self.rho, self.eps,
                          self.weight_decay)
        self.loss = loss.detach()
        return self.loss


class SLDTrainer(_FlatAdadeltaTrainer):
    def __init__(self, model: Transformer, lr: float = 1.0, rho: float = 0.9, eps: float = 1e-6, weight_decay: float = 0.0,
                 process_group=None):
        if not isinstance(model, Transformer):
            raise TypeError("SLDTrainer drives fudanocr_b200.model.transformer.Transformer")
        super().__init__(model, ("compress_attention_linear",), lr, rho, eps, weight_decay, process_group)

    def step(self, image: torch.Tensor, length: torch.Tensor, text_input: torch.Tensor, text_gt: torch.Tensor) -> torch.Tensor:
        """one optimisation step on device-resident tensors; returns this rank's (device) loss, nothing blocks the host"""
        self._begin()
        return self._finish(self.model.loss(image, length, text_input, text_gt))


class IDSTrainer(_FlatAdadeltaTrainer):
    """image-ids-CTR/train.py:28,63-90: Adadelta(lr, rho 0.9, weight_decay 1e-4) on loss_rec + 0.001 * loss_dis against frozen
    text features; `lr` may be passed per step (the reference drives it with CosineAnnealingWarmRestarts, train.py:29)"""

    def __init__(self, model, text_features: torch.Tensor, lr: float = 1.0, rho: float = 0.9, eps: float = 1e-6,
                 weight_decay: float = 1e-4, process_group=None):
        from .model.ids_transformer import Transformer as IDSTransformer
        if not isinstance(model, IDSTransformer):
            raise TypeError("IDSTrainer drives fudanocr_b200.model.ids_transformer.Transformer")
        super().__init__(model, ("compress_attention_linear", "encoder.layer4"), lr, rho, eps, weight_decay, process_group)
        self.text_features = text_features.float().contiguous()
        self.text_features_padded = model.pad_text_features(self.text_features)
        self.loss_rec = self.loss_dis = None

    def step(self, image, length, text_input, text_gt, lr: Optional[float] = None) -> torch.Tensor:
        self._begin()
        loss, rec, dis = self.model.loss(image, length, text_input, text_gt, self.text_features, self.text_features_padded)
        self.loss_rec, self.loss_dis = rec.detach(), dis.detach()
        return self._finish(loss, lr)
