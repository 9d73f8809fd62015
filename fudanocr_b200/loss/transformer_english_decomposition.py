"""Parameter container of the frozen stroke recogniser — drop-in for ``loss.transformer_english_decomposition.Transformer``
(text-gestalt/loss/transformer_english_decomposition.py:343-362) and, with ``variant='stt'``, for
``loss.transformer.Transformer`` (scene-text-telescope/loss/transformer.py:348-362).

Same ``state_dict`` keys and shapes as the reference, so ``pretrain_transformer_stroke_decomposition.pth`` /
``pretrain_transformer.pth`` load unchanged (the ``module.`` prefix of their DataParallel wrapper is stripped by the loss
modules).  Nothing here computes: ``fudanocr_b200.loss.stroke_focus_loss`` hands the tensors to the CUDA engine
(``focr_strokenet_prepare`` / ``focr_focus_loss``, include/focr.h).  There is no PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
from torch import nn

from .. import _lib as L

TG_ALPHABET = "0123456789"                                   # transformer_english_decomposition.py:8
STT_ALPHABET = "-0123456789abcdefghijklmnopqrstuvwxyz"        # scene-text-telescope/loss/transformer.py:8


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the focr engine; use StrokeFocusLoss / TextFocusLoss")


def _conv_shapes() -> List[Tuple[int, int]]:
    """(cin, cout) of the encoder's 31 convolutions in state_dict order (ResNet(1, BasicBlock, [1,2,5,3]), :70-128)"""
    out = [(1, 64), (64, 128)]
    nblk, cin, cout = [1, 2, 5, 3], [128, 256, 256, 512], [256, 256, 512, 512]
    for li in range(4):
        for bi in range(nblk[li]):
            ci = cin[li] if bi == 0 else cout[li]
            out += [(ci, cout[li]), (cout[li], cout[li])]
            if bi == 0 and ci != cout[li]:
                out.append((ci, cout[li]))
        out.append((cout[li], cout[li]) if li < 3 else (512, 1024))
    return out


def _positional_encoding(d_model: int = 512, max_len: int = 5000) -> torch.Tensor:
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len).unsqueeze(1).float()
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def slot_shapes(n_class: int) -> List[Tuple[int, ...]]:
    shp: List[Tuple[int, ...]] = [(n_class, 512), (1, 5000, 512)]
    for ci, co in _conv_shapes():
        shp += [(co, ci, 3, 3), (co,), (co,), (co,), (co,), (co,), ()]
    for _ in range(2):
        shp += [(1024, 1024), (1024,)] * 4 + [(1, 16), (1,), (1024,), (1024,)]
    shp += [(2048, 1024), (2048,), (1024, 2048), (1024,), (1024,), (1024,), (n_class, 1024), (n_class,)]
    return shp


class Transformer(_Holder):
    def __init__(self, variant: str = "tg"):
        super().__init__()
        self.variant = 0 if variant == "tg" else 1
        self.n_class = len(TG_ALPHABET if self.variant == 0 else STT_ALPHABET)
        n = L.lib.focr_strokenet_num_slots()
        names = [L.lib.focr_strokenet_slot_name(self.variant, i).decode() for i in range(n)]
        shapes = slot_shapes(self.n_class)
        assert len(shapes) == n, (len(shapes), n)
        self._names = names
        for name, shape in zip(names, shapes):
            mod: nn.Module = self
            parts = name.split(".")
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Holder())
                mod = mod._modules[p]
            leaf = parts[-1]
            if leaf == "num_batches_tracked":
                mod.register_buffer(leaf, torch.zeros((), dtype=torch.long))
            elif leaf == "pe":
                mod.register_buffer(leaf, _positional_encoding())
            elif leaf == "running_mean":
                mod.register_buffer(leaf, torch.zeros(shape))
            elif leaf == "running_var":
                mod.register_buffer(leaf, torch.ones(shape))
            else:
                t = torch.empty(shape)
                if len(shape) > 1:
                    nn.init.xavier_uniform_(t)          # Transformer.__init__ (:358-360)
                elif leaf == "a_2" or (leaf == "weight" and (parts[-2].startswith("bn") or parts[-2].endswith("_bn")
                                                             or parts[-3:-1] == ["downsample", "1"])):
                    t.fill_(1.0)                        # LayerNorm gain / BatchNorm weight
                else:
                    t.zero_()
                mod.register_parameter(leaf, nn.Parameter(t))

    def slot_tensors(self) -> List[torch.Tensor]:
        sd = dict(self.state_dict(keep_vars=True))
        return [sd[n] for n in self._names]
