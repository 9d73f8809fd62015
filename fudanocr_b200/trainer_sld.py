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

from . import _lib as L
from .model import recog_ops as ops
from .model.transformer import Transformer

_CHUNK = 65536   # elements per optimizer CTA: ~1100 CTAs for 71.7 M parameters
_T_BUCKET = 8    # graph replay: the text length is padded up to a multiple of this, so a training run needs <= 4 graphs


class _FlatAdadeltaTrainer:
    """parameters, gradients and the two Adadelta states in flat fp32 buffers; autograd accumulates straight into the gradient
    buffer; `skip` names the parameters the reference forward never touches (their .grad stays None there and its optimiser
    skips them - which matters once weight decay is on)"""

    def __init__(self, model, skip, lr, rho, eps, weight_decay, process_group, use_graph=True):
        self.model = model
        # The step - ~780 kernel launches issued from Python through autograd - is replayed as ONE CUDA graph per input shape
        # (forward, loss, backward, the gradient all-reduce, Adadelta): the first step with a new shape runs eagerly (it also
        # sets every kernel attribute), the second is captured, later ones are a copy into static buffers + one graph launch.
        # Seeds are frozen at capture time, so the graph advances the kernels' dropout epoch on the device (focr.h).
        self.use_graph = bool(use_graph)
        self._graphs = {}        # shape key -> (graph, static inputs, static outputs)
        self._seen = {}          # shape key -> eager steps done
        self._pool = None
        self.kernel_launches = 0   # launches of focr kernels issued (eager) or replayed (graph nodes) by step()
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
        # parameter data pointer -> its gradient view: inside a step the backward kernels write there directly (model/transformer.py)
        self._sinks = {p.data_ptr(): g for p, g in zip(plist, self.grad_views)}
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
        from .model.transformer import grad_sinks
        with grad_sinks(self._sinks):
            loss.backward()
        if self.world > 1:
            dist.all_reduce(self.flat_g, group=self.pg)    # sum; the mean is taken inside the optimizer kernel
        ops.adadelta_step(self.chunks, self.chunks.shape[0], 1.0 / self.world, self.lr if lr is None else lr, self.rho, self.eps,
                          self.weight_decay)
        self.loss = loss.detach()
        return self.loss

    # ---- graph replay -----------------------------------------------------------------------------------------------------------
    def _graphable(self, image):
        return (self.use_graph and image.is_cuda and not L.prof_enabled() and not L.status_checks()
                and not torch.cuda.is_current_stream_capturing())

    @staticmethod
    def _pad_text(length, text_input, text_gt):
        """static-shape form of the label tensors: T padded to a multiple of _T_BUCKET (positions t >= length[b] carry the start
        symbol 0, exactly like the columns the reference's converter leaves untouched; the decoder is causal and the loss only
        reads t < length[b], so the padded step computes the same loss and gradients), text_gt padded to B * T entries"""
        B, T = text_input.shape
        Tb = (T + _T_BUCKET - 1) // _T_BUCKET * _T_BUCKET
        return B, T, Tb

    def _run_graph(self, image, length, text_input, text_gt, body, lr=None):
        """body(image, length, text_input, text_gt) -> tuple of device scalars (total loss first); runs one eager step per new
        shape, then captures, then replays.  Returns the tuple of (static) outputs."""
        B, T, Tb = self._pad_text(length, text_input, text_gt)
        # what the captured launches bake in; a per-step learning rate (image-ids-CTR drives it with a per-EPOCH scheduler,
        # train.py:29,92) is a host number in the optimiser launch: one graph per value, the oldest dropped beyond eight
        key = (tuple(image.shape), image.dtype, Tb, getattr(self.model, "dropout_p", None), lr)
        ent = self._graphs.get(key)
        if ent is None:
            if self._seen.get(key, 0) < 1:            # first step of this shape: eager (also the kernels' lazy initialisation)
                self._seen[key] = 1
                self._begin()
                return None
            dev = image.device
            st_image = torch.empty_like(image, memory_format=torch.contiguous_format)
            st_len = torch.zeros(B, dtype=torch.long, device=dev)
            st_in = torch.zeros(B, Tb, dtype=torch.long, device=dev)
            st_gt = torch.zeros(B * Tb, dtype=torch.long, device=dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                st_image.copy_(image)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()    # graphs of different shapes replay one at a time: one pool
            st_len.copy_(length)
            st_in[:, :T].copy_(text_input)
            st_gt[:text_gt.numel()].copy_(text_gt)
            n0 = L.lib.focr_launch_count()
            try:
                with torch.cuda.graph(g, pool=self._pool):
                    L.check(L.lib.focr_recog_epoch_advance(L.cur_stream()), "recog_epoch_advance")
                    self._begin()
                    outs = body(st_image, st_len, st_in, st_gt)
            except Exception as ex:   # noqa: BLE001 - e.g. a collective the installed NCCL cannot capture: keep training eagerly
                import warnings
                warnings.warn(f"focr: the recogniser step could not be captured as a CUDA graph ({ex}); continuing with eager launches")
                self.use_graph = False
                torch.cuda.synchronize()
                self._begin()
                return None
            ent = (g, (st_image, st_len, st_in, st_gt), outs, int(L.lib.focr_launch_count() - n0))
            if len(self._graphs) >= 8:
                del self._graphs[next(iter(self._graphs))]
            self._graphs[key] = ent
        g, (st_image, st_len, st_in, st_gt), outs, n_nodes = ent
        self.kernel_launches += n_nodes
        st_image.copy_(image, non_blocking=True)
        st_len.copy_(length, non_blocking=True)
        if Tb != T:
            st_in.zero_()
        st_in[:, :T].copy_(text_input, non_blocking=True)
        st_gt[:text_gt.numel()].copy_(text_gt, non_blocking=True)
        g.replay()
        return outs


class SLDTrainer(_FlatAdadeltaTrainer):
    def __init__(self, model: Transformer, lr: float = 1.0, rho: float = 0.9, eps: float = 1e-6, weight_decay: float = 0.0,
                 process_group=None, use_graph: bool = True):
        if not isinstance(model, Transformer):
            raise TypeError("SLDTrainer drives fudanocr_b200.model.transformer.Transformer")
        super().__init__(model, ("compress_attention_linear",), lr, rho, eps, weight_decay, process_group, use_graph)

    def step(self, image: torch.Tensor, length: torch.Tensor, text_input: torch.Tensor, text_gt: torch.Tensor) -> torch.Tensor:
        """one optimisation step on device-resident tensors; returns this rank's (device) loss, nothing blocks the host"""
        if self._graphable(image):
            outs = self._run_graph(image, length, text_input, text_gt,
                                   lambda im, ln, ti, gt: (self._finish(self.model.loss(im, ln, ti, gt)),))
            if outs is not None:
                self.loss = outs[0]
                return self.loss
        else:
            self._begin()
        n0 = L.lib.focr_launch_count()
        out = self._finish(self.model.loss(image, length, text_input, text_gt))
        self.kernel_launches += int(L.lib.focr_launch_count() - n0)
        return out


class IDSTrainer(_FlatAdadeltaTrainer):
    """image-ids-CTR/train.py:28,63-90: Adadelta(lr, rho 0.9, weight_decay 1e-4) on loss_rec + 0.001 * loss_dis against frozen
    text features; `lr` may be passed per step (the reference drives it with CosineAnnealingWarmRestarts, train.py:29)"""

    def __init__(self, model, text_features: torch.Tensor, lr: float = 1.0, rho: float = 0.9, eps: float = 1e-6,
                 weight_decay: float = 1e-4, process_group=None, use_graph: bool = True):
        from .model.ids_transformer import Transformer as IDSTransformer
        if not isinstance(model, IDSTransformer):
            raise TypeError("IDSTrainer drives fudanocr_b200.model.ids_transformer.Transformer")
        super().__init__(model, ("compress_attention_linear", "encoder.layer4"), lr, rho, eps, weight_decay, process_group,
                         use_graph)
        self.text_features = text_features.float().contiguous()
        self.text_features_padded = model.pad_text_features(self.text_features)
        self.loss_rec = self.loss_dis = None

    def step(self, image, length, text_input, text_gt, lr: Optional[float] = None) -> torch.Tensor:
        """`lr`: this step's learning rate (None = the constructor's); constant within an epoch under the reference's scheduler"""
        def body(im, ln, ti, gt):
            loss, rec, dis = self.model.loss(im, ln, ti, gt, self.text_features, self.text_features_padded)
            return self._finish(loss, lr), rec.detach(), dis.detach()
        if self._graphable(image):
            outs = self._run_graph(image, length, text_input, text_gt, body, None if lr is None else float(lr))
            if outs is not None:
                self.loss, self.loss_rec, self.loss_dis = outs
                return self.loss
        else:
            self._begin()
        n0 = L.lib.focr_launch_count()
        self.loss, self.loss_rec, self.loss_dis = body(image, length, text_input, text_gt)
        self.kernel_launches += int(L.lib.focr_launch_count() - n0)
        return self.loss
