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

_CHUNK = 65536


class TBSRNTrainer:
    def __init__(self, model: TBSRN, lr: float = 1e-4, betas=(0.5, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 0.25, loss_scale: float = 100.0, process_group=None, criterion=None):
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
        if self.world > 1:  # start from identical weights on every rank (rank 0's)
            dist.broadcast(self.flat_p, src=0, group=self.pg)

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
        flags = 1 | (2 if m.stn else 0)
        p = m.dropout_p
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if p > 0 else 0
        st = L.cur_stream()
        lib = L.lib
        L.check(m._c_forward(self.ptable, images_lr, self.sr, B, flags, p, seed, ws), "sr_forward")
        if self.criterion is None:
            L.check(lib.focr_mse_loss_grad(self.sr.data_ptr(), images_hr.data_ptr(), self.d_sr.data_ptr(),
                                           self.loss.data_ptr(), self.sr.numel(), self.loss_scale, self.scratch.data_ptr(),
                                           self.scratch.numel(), st), "mse_loss_grad")
        else:
            self.losses = self.criterion.loss_and_grad(self.sr, images_hr, labels, self.loss_scale, self.d_sr)
            self.loss = self.losses[0:1]
        L.check(m._c_backward(self.ptable, self.gtable, images_lr, self.d_sr, B, flags, p, seed, ws), "sr_backward")
        gscale = 1.0
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
            gscale = 1.0 / self.world  # mean over ranks == gradient of the global-batch mean loss
        L.check(lib.focr_adam_clip_step(self.chunks.data_ptr(), self.chunks.shape[0], gscale, self.max_grad_norm,
                                        self.lr, self.betas[0], self.betas[1], self.eps, self.step_count.data_ptr(),
                                        self.opt_state.data_ptr(), self.scratch.data_ptr(), self.scratch.numel(), st),
                "adam_clip_step")
        return self.loss

    @property
    def grad_norm(self) -> torch.Tensor:
        return self.opt_state[0]
