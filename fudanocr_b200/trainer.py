"""Fused training step for TBSRN on the focr engine — the B200 form of the reference step body
(scene-text-telescope/interfaces/super_resolution.py:69-84 with the MSE image_crit of
loss/text_focus_loss.py:86 and Adam from interfaces/base.py:194-198):

    sr = model(lr); loss = mse(sr, hr); (loss*100).backward(); clip_grad_norm_(0.25); Adam.step()

Everything runs as hand-written kernels on one stream with no host synchronisation: forward, MSE
gradient, backward into ONE flat fp32 gradient buffer, (multi-GPU: a single NCCL all-reduce of that buffer
over NVLink — the data-parallel exchange that replaces nn.DataParallel's per-step replicate/reduce,
interfaces/base.py:178-179), then the fused norm + clip + Adam kernel.  Parameters keep living in the
module's own nn.Parameters (re-pointed into one flat buffer), so ``state_dict`` / checkpoints are unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib as L
from .model.tbsrn import TBSRN, _SREngineModule

_CHUNK = 8192   # elements per optimizer CTA: ~600 CTAs for 3.2 M parameters (65536 left 50 CTAs streaming 90 MB: 0.34 ms)


class TBSRNTrainer:
    def __init__(self, model: TBSRN, lr: float = 1e-4, betas=(0.5, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 0.25, loss_scale: float = 100.0, process_group=None, criterion=None,
                 use_graph: bool = True):
        if not isinstance(model, _SREngineModule):
            raise TypeError("TBSRNTrainer drives the engine-backed SR models (fudanocr_b200.model.tbsrn.TBSRN / tsrn.TSRN)")
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.max_grad_norm, self.loss_scale = max_grad_norm, loss_scale
        self.pg = process_group
        # image_crit of the step body: None = the MSE term alone (text_focus off); otherwise a fudanocr_b200.loss
        # module with loss_and_grad() (StrokeFocusLoss: text-gestalt/interfaces/super_resolution.py:69, config 3)
        self.criterion = criterion
        self.losses: Optional[torch.Tensor] = None
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        tensors, _ = model._slots()
        self.slots = list(model._grad_slots)
        params = [tensors[i] for i in self.slots]
        dev = params[0].device
        sizes = [p.numel() for p in params]
        # 16-byte align every tensor inside the flat buffers
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot)
            tot += (n + 3) // 4 * 4
        self.flat_p = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(tot, dtype=torch.float32, device=dev)
        for p, o, n in zip(params, offs, sizes):
            self.flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + n].view_as(p)  # parameters now alias the flat buffer
        model._cache = None
        tensors, self.ptable = model._slots()
        gptr = [0] * len(tensors)
        recs = []
        for i, o, n in zip(self.slots, offs, sizes):
            gptr[i] = self.flat_g.data_ptr() + 4 * o
            k = 0
            while k < n:
                ln = min(_CHUNK, n - k)
                recs.append([self.flat_p.data_ptr() + 4 * (o + k), self.flat_g.data_ptr() + 4 * (o + k),
                             self.flat_m.data_ptr() + 4 * (o + k), self.flat_v.data_ptr() + 4 * (o + k), ln])
                k += ln
        self.gtable = (C.c_void_p * len(tensors))(*gptr)
        self.chunks = torch.tensor(recs, dtype=torch.int64, device=dev)
        self.step_count = torch.zeros((), dtype=torch.int64, device=dev)
        self.opt_state = torch.zeros(4, dtype=torch.float32, device=dev)   # grad norm, clip*gscale, lr_t, 1/sqrt(bc2)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.scratch = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
        self.d_sr: Optional[torch.Tensor] = None
        self.sr: Optional[torch.Tensor] = None
        self.kernel_launches = 0
        # The step is ~500 short launches: replaying it as a CUDA graph removes the launch gaps (measured 24.0 -> 22.8
        # ms/step at batch 256).  One graph: [forward + loss + backward -> gradient all-reduce -> clip + Adam] (see
        # _capture); the dropout seed lives in a device word the kernels read at run time, the batch in
        # static input buffers.  Captured lazily per batch size after one eager step (which also sets every kernel
        # attribute); a criterion with host-side label encoding keeps the eager path.
        self.use_graph = bool(use_graph)
        self._graphs = None
        self._graph_key = None
        self._eager_steps = 0
        self.seed_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.world > 1:  # start from identical weights on every rank (rank 0's)
            dist.broadcast(self.flat_p, src=0, group=self.pg)

    # ---- the three phases of a step (each enqueues kernels on the current stream, nothing else) ----------------
    def _fwd_loss_bwd(self, lr, hr, B, flags, p, seed, ws, labels, seed_dev=None):
        m, lib, st = self.model, L.lib, L.cur_stream()
        L.check(m._c_forward(self.ptable, lr, self.sr, B, flags, p, seed, ws, seed_dev), "sr_forward")
        if self.criterion is None:
            L.check(lib.focr_mse_loss_grad(self.sr.data_ptr(), hr.data_ptr(), self.d_sr.data_ptr(),
                                           self.loss.data_ptr(), self.sr.numel(), self.loss_scale, self.scratch.data_ptr(),
                                           self.scratch.numel(), st), "mse_loss_grad")
        else:
            self.losses = self.criterion.loss_and_grad(self.sr, hr, labels, self.loss_scale, self.d_sr)
            self.loss = self.losses[0:1]
        L.check(m._c_backward(self.ptable, self.gtable, lr, self.d_sr, B, flags, p, seed, ws), "sr_backward")

    def _optimizer(self):
        gscale = 1.0 / self.world  # mean over ranks == gradient of the global-batch mean loss
        L.check(L.lib.focr_adam_clip_step(self.chunks.data_ptr(), self.chunks.shape[0], gscale, self.max_grad_norm,
                                          self.lr, self.betas[0], self.betas[1], self.eps, self.step_count.data_ptr(),
                                          self.opt_state.data_ptr(), self.scratch.data_ptr(), self.scratch.numel(),
                                          L.cur_stream()), "adam_clip_step")

    def _capture(self, B, flags, p, ws, dev):
        self._lr_in = torch.empty(B, 3, 16, 64, dtype=torch.float32, device=dev)
        self._hr_in = torch.empty(B, 3, 32, 128, dtype=torch.float32, device=dev)
        use_seed_dev = self.seed_dev if (p > 0 and self.model._ARCH == "tbsrn") else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # kernels must have run once on a non-default stream before capture
            self._lr_in.copy_(self._last_lr)
            self._hr_in.copy_(self._last_hr)
        torch.cuda.current_stream().wait_stream(side)
        # ONE graph for the whole step: forward + loss + backward, the gradient all-reduce (NCCL collectives can be
        # captured: the exchange becomes a graph node between the last backward kernel and the optimizer instead of an
        # eager call wedged between two replays - at 32 crops per GPU that gap was ~13x the wire time), clip + Adam.
        # FOCR_GRAPH_NCCL=0, or a capture that NCCL refuses, falls back to two graphs with the collective between them.
        import os
        n0 = L.lib.focr_launch_count()
        one = self.world == 1 or os.environ.get("FOCR_GRAPH_NCCL", "1") != "0"
        if one:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._fwd_loss_bwd(self._lr_in, self._hr_in, B, flags, p, 0, ws, None, use_seed_dev)
                    if self.world > 1:
                        dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
                    self._optimizer()
                self._graph_launches = int(L.lib.focr_launch_count() - n0)   # kernel nodes replayed per step
                self._graphs = (g,)
                return
            except Exception as ex:   # noqa: BLE001 - keep training on the two-graph path
                if self.world == 1:
                    raise
                import warnings
                warnings.warn(f"focr: NCCL all-reduce could not be captured into the step graph ({ex}); using two graphs")
                torch.cuda.synchronize()
                n0 = L.lib.focr_launch_count()
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self._fwd_loss_bwd(self._lr_in, self._hr_in, B, flags, p, 0, ws, None, use_seed_dev)
        with torch.cuda.graph(g2):
            self._optimizer()
        self._graph_launches = int(L.lib.focr_launch_count() - n0)   # kernel nodes replayed per step
        self._graphs = (g1, g2)

    def step(self, images_lr: torch.Tensor, images_hr: torch.Tensor, seed: Optional[int] = None,
             labels=None) -> torch.Tensor:
        """One optimisation step on device-resident fp32 NCHW batches.  Returns the (device) loss tensor of this
        rank's shard (MSE, or the criterion's total loss; its parts are in ``self.losses``); nothing here blocks
        the host."""
        m = self.model
        B = images_lr.shape[0]
        dev = images_lr.device
        ws = m._workspace(B, dev)
        if self.sr is None or self.sr.shape[0] != B:
            self.sr = torch.empty(B, 3, 32, 128, dtype=torch.float32, device=dev)
            self.d_sr = torch.empty_like(self.sr)
            self._graphs = None
        flags = 1 | (2 if m.stn else 0)
        p = m.dropout_p
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if p > 0 else 0
        if self.world > 1 and p > 0:   # replicas draw independent dropout masks, as nn.DataParallel's do
            seed ^= (self.rank * 0x9E3779B1) & 0x7FFFFFFF
        key = (B, flags, p, ws.data_ptr())
        graphable = (self.use_graph and self.criterion is None and not L.prof_enabled() and images_lr.is_contiguous()
                     and images_hr.is_contiguous() and images_lr.dtype == torch.float32 and images_hr.dtype == torch.float32)
        if graphable and self._graph_key == key and self._eager_steps >= 1:
            if self._graphs is None:
                self._last_lr, self._last_hr = images_lr, images_hr
                self._capture(B, flags, p, ws, dev)
            self.seed_dev.fill_(seed & 0x7FFFFFFF)   # value travels as a kernel argument: no host buffer to race on
            self._lr_in.copy_(images_lr, non_blocking=True)
            self._hr_in.copy_(images_hr, non_blocking=True)
            self._graphs[0].replay()
            if len(self._graphs) == 2:
                if self.world > 1:
                    dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
                self._graphs[1].replay()
            self.kernel_launches += self._graph_launches
            return self.loss
        if self._graph_key != key:
            self._graph_key, self._graphs, self._eager_steps = key, None, 0
        self._eager_steps += 1
        n0 = L.lib.focr_launch_count()
        self._fwd_loss_bwd(images_lr, images_hr, B, flags, p, seed & 0x7FFFFFFF if p > 0 else seed, ws, labels)
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
        self._optimizer()
        self.kernel_launches += int(L.lib.focr_launch_count() - n0)
        return self.loss

    @property
    def grad_norm(self) -> torch.Tensor:
        return self.opt_state[0]
