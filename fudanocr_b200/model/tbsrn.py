"""TBSRN on the focr sm_100a engine — drop-in for ``model.tbsrn.TBSRN`` of the reference
(scene-text-telescope/model/tbsrn.py:166-226; text-gestalt/model/tbsrn.py is the same network).

Same constructor signature, same ``state_dict`` keys (346 entries, including the members the
reference builds but never calls: ``conv``/``bn``, every SRB's ``gru1``/``gru2`` and
``compress_attention_linear``), same ``forward((B,3,16,64)) -> (B,3,32,128)``, same train/eval
semantics (STN rectification only in ``train()``, tbsrn.py:215).  The submodules below are only
*parameter containers*: nothing in this file computes with torch ops.  ``forward`` hands raw device
pointers to ``focr_tbsrn_forward`` / ``focr_tbsrn_backward`` (include/focr.h) through one
``torch.autograd.Function``; there is no PyTorch fallback path.
"""
from __future__ import annotations

import ctypes as C
import weakref
import math
from typing import Dict, List

import numpy as np
import torch
from torch import nn

from .. import _lib as L

__all__ = ["TBSRN"]


# ---------------------------------------------------------------------------------------------
# parameter containers (names mirror the reference so checkpoints interchange)
# ---------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    """A module whose forward must never run: its parameters are consumed by the CUDA engine."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is a parameter container of the focr engine; "
                           "call the enclosing TBSRN instead")


class _StdLayerNorm(_Holder):  # tbsrn.py:23-36
    def __init__(self, features: int):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(features))
        self.b_2 = nn.Parameter(torch.zeros(features))


class _MultiHead(_Holder):  # tbsrn.py:95-107
    def __init__(self, h: int, d_model: int, dropout: float):
        super().__init__()
        self.linears = nn.ModuleList([nn.Linear(d_model, d_model) for _ in range(4)])
        self.dropout = nn.Dropout(p=dropout)
        self.compress_attention_linear = nn.Linear(h, 1)  # dead in the reference as well


class _FeedForward(_Holder):  # tbsrn.py:153-163
    def __init__(self, d_model: int, d_ff: int, dropout: float = 0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_ff)
        self.w_2 = nn.Linear(d_ff, d_model)
        self.dropout = nn.Dropout(dropout)


class _FeatureEnhancer(_Holder):  # tbsrn.py:63-74
    def __init__(self):
        super().__init__()
        self.multihead = _MultiHead(4, 128, 0.1)
        self.mul_layernorm1 = _StdLayerNorm(128)
        self.pff = _FeedForward(128, 128)
        self.mul_layernorm3 = _StdLayerNorm(128)
        self.linear = nn.Linear(128, 64)


class _GruBlock(_Holder):  # tbsrn.py:288-295 (constructed, never called by TBSRN)
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, kernel_size=1, padding=0)
        self.gru = nn.GRU(cout, cout // 2, bidirectional=True, batch_first=True)


class _SRB(_Holder):  # RecurrentResidualBlock, tbsrn.py:229-244
    def __init__(self, ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, ch, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(ch)
        self.gru1 = _GruBlock(ch, ch)
        self.conv2 = nn.Conv2d(ch, ch, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(ch)
        self.gru2 = _GruBlock(ch, ch)
        self.feature_enhancer = _FeatureEnhancer()
        for p in self.parameters():  # tbsrn.py:242-244
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class _Upsample(_Holder):  # UpsampleBLock, tbsrn.py:261-268
    def __init__(self, ch: int, up: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch * up ** 2, kernel_size=3, padding=1)


class _Seq(nn.Sequential):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the focr engine")


def _identity_ctrl_points(n: int, margin: float = 0.01) -> torch.Tensor:
    k = n // 2
    xs = np.linspace(margin, 1.0 - margin, k)
    pts = np.concatenate([np.stack([xs, np.full(k, margin)], 1), np.stack([xs, np.full(k, 1 - margin)], 1)], 0)
    return torch.tensor(pts.astype(np.float32)).reshape(-1)


class _STNHead(_Holder):  # model/stn_head.py:25-86
    def __init__(self, in_planes: int, num_ctrlpoints: int):
        super().__init__()
        chans = [(in_planes, 32), (32, 64), (64, 128), (128, 256), (256, 256), (256, 256)]
        layers: List[nn.Module] = []
        for i, (ci, co) in enumerate(chans):
            layers.append(_Seq(nn.Conv2d(ci, co, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(co),
                               nn.ReLU(inplace=True)))
            if i < 4:
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            elif i == 4:
                layers.append(nn.MaxPool2d(kernel_size=(1, 2), stride=(1, 2)))
        self.stn_convnet = _Seq(*layers)
        self.stn_fc1 = _Seq(nn.Linear(2 * 256, 512), nn.BatchNorm1d(512), nn.ReLU(inplace=True))
        self.stn_fc2 = nn.Linear(512, num_ctrlpoints * 2)
        for m in list(self.stn_convnet.modules()) + list(self.stn_fc1.modules()):  # stn_head.py:55-67
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan))
                m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.001)
                m.bias.data.zero_()
        self.stn_fc2.weight.data.zero_()  # stn_head.py:85-86: start from the identity warp
        self.stn_fc2.bias.data = _identity_ctrl_points(num_ctrlpoints)


class _TPS(_Holder):  # model/tps_spatial_transformer.py:54-95 (buffers only)
    def __init__(self, out_hw, n_ctrl: int, margins):
        super().__init__()
        h, w = out_hw
        k = n_ctrl // 2
        xs = np.linspace(margins[0], 1.0 - margins[0], k)
        tcp = torch.tensor(np.concatenate([np.stack([xs, np.full(k, margins[1])], 1),
                                           np.stack([xs, np.full(k, 1.0 - margins[1])], 1)], 0), dtype=torch.float32)

        def rbf(a, b):
            d = a.view(-1, 1, 2) - b.view(1, -1, 2)
            r2 = (d * d).sum(-1)
            out = 0.5 * r2 * torch.log(r2)
            return torch.where(torch.isnan(out), torch.zeros_like(out), out)

        fk = torch.zeros(n_ctrl + 3, n_ctrl + 3)
        fk[:n_ctrl, :n_ctrl] = rbf(tcp, tcp)
        fk[:n_ctrl, -3] = 1
        fk[-3, :n_ctrl] = 1
        fk[:n_ctrl, -2:] = tcp
        fk[-2:, :n_ctrl] = tcp.t()
        yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                                indexing="ij")
        coord = torch.stack([xx.reshape(-1) / (w - 1), yy.reshape(-1) / (h - 1)], 1)
        self.register_buffer("inverse_kernel", torch.inverse(fk).contiguous())  # torch.inverse returns column-major
        self.register_buffer("padding_matrix", torch.zeros(3, 2))
        self.register_buffer("target_coordinate_repr",
                             torch.cat([rbf(coord, tcp), torch.ones(h * w, 1), coord], 1).contiguous())
        self.register_buffer("target_control_points", tcp)


# ---------------------------------------------------------------------------------------------
# autograd bridge
# ---------------------------------------------------------------------------------------------
class _FwdState(dict):
    """state of one engine forward (a dict that can be weakly referenced)"""


class _EngineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, model, *params):
        st = model._launch_forward(x, pending=True)
        ctx.model, ctx.st = model, st
        ctx.set_materialize_grads(False)
        return st["sr"]

    @staticmethod
    def backward(ctx, d_sr):
        model, st = ctx.model, ctx.st
        if d_sr is None:
            return (None, None) + (None,) * len(model._grad_slots)
        grads = model._launch_backward(st, d_sr)
        return (None, None) + tuple(grads)


class _SREngineModule(nn.Module):
    """Shared plumbing of the engine-backed SR networks (TBSRN, TSRN): slot table, workspace, launches."""
    _ARCH = "tbsrn"

    def _engine_init(self, srb_nums: int):
        lib = L.lib
        if self._ARCH == "tsrn":
            n = lib.focr_tsrn_num_slots(srb_nums)
            self._slot_names = [lib.focr_tsrn_slot_name(srb_nums, i).decode() for i in range(n)]
        else:
            n = lib.focr_tbsrn_num_slots(srb_nums)
            self._slot_names = [lib.focr_tbsrn_slot_name(srb_nums, i).decode() for i in range(n)]
        self._cache = None   # (slot tensors, pointer table) — invalidated by _apply (.to / .cuda / .float)
        self._ws: Dict[int, torch.Tensor] = {}
        self._ws_gen = 0
        self._ws_stamp: Dict[int, int] = {}   # arena address -> generation of the forward whose activations it holds
        self._ws_owner: Dict[int, "weakref.ref"] = {}   # arena address -> forward state whose backward is still pending

    def _ws_bytes(self, B: int) -> int:
        if self._ARCH == "tsrn":
            return L.lib.focr_tsrn_workspace_bytes(B, self.srb_nums)
        return L.lib.focr_tbsrn_workspace_bytes(B, self.srb_nums)

    def _c_forward(self, table, x, sr, B, flags, p, seed, ws, seed_dev=None):
        if self._ARCH == "tsrn":
            return L.lib.focr_tsrn_forward(table, x.data_ptr(), sr.data_ptr(), B, self.srb_nums, flags, ws.data_ptr(),
                                           ws.numel(), L.cur_stream())
        if seed_dev is not None:   # seed read on the device at run time (CUDA-graph replay)
            return L.lib.focr_tbsrn_forward_devseed(table, x.data_ptr(), sr.data_ptr(), B, self.srb_nums, flags, p,
                                                    seed_dev.data_ptr(), ws.data_ptr(), ws.numel(), L.cur_stream())
        return L.lib.focr_tbsrn_forward(table, x.data_ptr(), sr.data_ptr(), B, self.srb_nums, flags, p, seed,
                                        ws.data_ptr(), ws.numel(), L.cur_stream())

    def _c_backward(self, table, gtable, x, d_sr, B, flags, p, seed, ws):
        if self._ARCH == "tsrn":
            return L.lib.focr_tsrn_backward(table, gtable, x.data_ptr(), d_sr.data_ptr(), B, self.srb_nums, flags,
                                            ws.data_ptr(), ws.numel(), L.cur_stream())
        return L.lib.focr_tbsrn_backward(table, gtable, x.data_ptr(), d_sr.data_ptr(), B, self.srb_nums, flags, p, seed,
                                         ws.data_ptr(), ws.numel(), L.cur_stream())

    # -- plumbing --------------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._cache = None
        self._ws = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._cache = None
        return super().load_state_dict(*a, **k)

    @property
    def dropout_p(self) -> float:
        ps = {m.p for m in self.modules() if isinstance(m, nn.Dropout)}
        if len(ps) > 1:
            raise RuntimeError(f"focr TBSRN needs one dropout rate for all SRBs, found {sorted(ps)}")
        return float(ps.pop()) if ps else 0.0

    def _slots(self):
        """slot tensors in engine order; STN slots get placeholders when the model has no STN"""
        if self._cache is not None:
            return self._cache
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        tensors = []
        for name in self._slot_names:
            t = sd.get(name)
            if t is None:
                if not (name.startswith("stn_head.") or name.startswith("tps.")):
                    raise KeyError(name)
                tensors.append(None)
                continue
            if not t.is_contiguous() and not isinstance(t, nn.Parameter):
                # buffers (e.g. a loaded column-major TPS matrix) are re-laid-out in place
                mod, leaf = self, name
                while "." in leaf:
                    head, leaf = leaf.split(".", 1)
                    mod = getattr(mod, head)
                t = t.contiguous()
                setattr(mod, leaf, t)
            if not t.is_cuda or not t.is_contiguous() or t.dtype not in (torch.float32, torch.int64):
                raise L.FocrError(f"parameter {name}: the focr engine needs contiguous fp32 CUDA tensors "
                                  f"(got {t.dtype} on {t.device}); there is no CPU path")
            tensors.append(t)
        table = (C.c_void_p * len(tensors))(*[0 if t is None else t.data_ptr() for t in tensors])
        # every nn.Parameter slot, whatever its requires_grad flag says right now: the reference's eval() switches the flag
        # off on all parameters and the train loop switches it back on (interfaces/super_resolution.py:166-170, :61-62) -
        # a cache keyed on the flag would be empty forever if the first forward happened while frozen
        grad_slots = [i for i, t in enumerate(tensors)
                      if t is not None and isinstance(t, nn.Parameter)
                      and (self.stn or not self._slot_names[i].startswith("stn_head."))]
        self._cache = (tensors, table)
        self._grad_slots = grad_slots
        return self._cache

    def _workspace(self, B: int, device) -> torch.Tensor:
        ws = self._ws.get(B)
        if ws is None or ws.device != device:
            nbytes = self._ws_bytes(B)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws = {B: ws}  # keep one batch size resident
        return ws

    def _forward_workspace(self, B: int, device):
        """workspace for one forward.  The arena holds the activations the backward reads, so a forward whose backward is
        still pending OWNS its arena: a second forward of the same batch size before that backward (two micro-batches summed
        into one loss, a validation forward between forward and backward) gets a fresh arena instead of overwriting it.
        Ownership is a weak reference to the forward's state (held by its autograd node): a forward whose graph was dropped
        without a backward releases its arena by itself.  Returns (arena, generation stamp)."""
        ws = self._workspace(B, device)
        owner = self._ws_owner.get(ws.data_ptr())
        if owner is not None and owner() is not None:
            ws = torch.empty(ws.numel(), dtype=torch.uint8, device=device)   # not cached: dies with its autograd node
        self._ws_gen += 1
        self._ws_stamp[ws.data_ptr()] = self._ws_gen
        return ws, self._ws_gen

    def _launch_forward(self, x: torch.Tensor, pending: bool = False) -> dict:
        if not x.is_cuda:
            raise L.FocrError(f"focr {type(self).__name__} runs on CUDA (sm_100a) only; move the input to the GPU")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 16, 64):
            raise ValueError(f"{type(self).__name__} expects (B,3,16,64) LR crops, got {tuple(x.shape)}")
        tensors, table = self._slots()
        x = x.detach().contiguous().float()
        B = x.shape[0]
        ws, gen = self._forward_workspace(B, x.device)
        sr = torch.empty(B, 3, 32, 128, dtype=torch.float32, device=x.device)
        training = bool(self.training)
        flags = (1 if training else 0) | (2 if self.stn else 0)
        p = self.dropout_p if training else 0.0
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if p > 0 else 0
        with torch.cuda.device(x.device):
            L.check(self._c_forward(table, x, sr, B, flags, p, seed, ws), f"focr_{self._ARCH}_forward")
        st = _FwdState(x=x, sr=sr, ws=ws, gen=gen, B=B, flags=flags, p=p, seed=seed)
        if pending:
            self._ws_owner[ws.data_ptr()] = weakref.ref(st)
        else:
            self._ws_owner.pop(ws.data_ptr(), None)
        return st

    def _launch_backward(self, st: dict, d_sr: torch.Tensor):
        if self._ws_stamp.get(st["ws"].data_ptr()) != st["gen"]:
            raise L.FocrError("the activations of this forward have been overwritten by a later forward of the same batch size "
                              "(backward called twice on one forward?); run forward again before backward")
        self._ws_owner.pop(st["ws"].data_ptr(), None)   # the arena is free for the next forward
        tensors, table = self._slots()
        sizes = [tensors[i].numel() for i in self._grad_slots]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=d_sr.device)
        views = list(flat.split(sizes))
        gptr = [0] * len(tensors)
        for i, v in zip(self._grad_slots, views):
            gptr[i] = v.data_ptr()
        gtable = (C.c_void_p * len(tensors))(*gptr)
        d_sr = d_sr.contiguous().float()
        with torch.cuda.device(d_sr.device):
            L.check(self._c_backward(table, gtable, st["x"], d_sr, st["B"], st["flags"], st["p"], st["seed"], st["ws"]),
                    f"focr_{self._ARCH}_backward")
        return [v.view_as(tensors[i]) for i, v in zip(self._grad_slots, views)]

    # -- public API ------------------------------------------------------------------------------
    def forward(self, x):
        tensors, _ = self._slots()
        if self.training and torch.is_grad_enabled():
            return _EngineFn.apply(x, self, *[tensors[i] for i in self._grad_slots])
        return self._launch_forward(x)["sr"]


class TBSRN(_SREngineModule):
    def __init__(self, scale_factor=2, width=128, height=32, STN=True, srb_nums=5, mask=False, hidden_units=32,
                 input_channel=3):
        super().__init__()
        if mask or input_channel != 3:
            raise NotImplementedError("focr TBSRN: the 4-channel (mask=True) variant is not built; "
                                      "reference default is mask=False (interfaces/base.py:141-142)")
        if scale_factor != 2 or hidden_units != 32 or (width, height) != (128, 32):
            # the reference hard-wires the 16x64 positional encoding and 64 channels (tbsrn.py:83)
            raise NotImplementedError("focr TBSRN supports scale_factor=2, hidden_units=32, 128x32 only")
        ch = 2 * hidden_units
        self.conv = nn.Conv2d(input_channel, 3, 3, 1, 1)  # dead members kept for state_dict parity
        self.bn = nn.BatchNorm2d(3)
        self.relu = nn.ReLU()
        self.block1 = _Seq(nn.Conv2d(3, ch, kernel_size=9, padding=4), nn.PReLU())
        self.srb_nums = srb_nums
        for i in range(srb_nums):
            setattr(self, f"block{i + 2}", _SRB(ch))
        setattr(self, f"block{srb_nums + 2}", _Seq(nn.Conv2d(ch, ch, kernel_size=3, padding=1), nn.BatchNorm2d(ch)))
        setattr(self, f"block{srb_nums + 3}", _Seq(_Upsample(ch, 2), nn.Conv2d(ch, 3, kernel_size=9, padding=4)))
        self.tps_inputsize = [height // scale_factor, width // scale_factor]
        self.stn = STN
        if self.stn:
            self.tps = _TPS(tuple(self.tps_inputsize), 20, (0.05, 0.05))
            self.stn_head = _STNHead(3, 20)
        self._engine_init(srb_nums)

