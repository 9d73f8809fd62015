"""CRNN evaluator on the focr sm_100a engine — drop-in for ``model.crnn.CRNN`` of the reference
(scene-text-telescope/model/crnn/crnn.py:25-80; text-gestalt's copy is byte-identical).

Same constructor, same 49 ``state_dict`` keys (``cnn.conv0.weight`` … ``rnn.1.embedding.bias``) so ``crnn.pth``
loads unchanged (interfaces/base.py:316), same ``forward((B,1,32,100)) -> (26,B,37)``.  The reference only ever
uses this network frozen and in ``eval()`` (base.py:309-317, super_resolution.py:166-171); the engine therefore
implements inference only and refuses ``train()`` mode.  No PyTorch fallback: the submodules hold parameters only.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from ... import _lib as L


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the focr engine")


class BidirectionalLSTM(_Holder):  # crnn.py:6-12
    def __init__(self, nIn, nHidden, nOut):
        super().__init__()
        self.rnn = nn.LSTM(nIn, nHidden, bidirectional=True)
        self.embedding = nn.Linear(nHidden * 2, nOut)


class _Seq(nn.Sequential):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the focr engine")


class CRNN(nn.Module):
    def __init__(self, imgH, nc, nclass, nh, n_rnn=2, leakyRelu=False):
        super().__init__()
        if (imgH, nc, nclass, nh, n_rnn, leakyRelu) != (32, 1, 37, 256, 2, False):
            raise NotImplementedError("focr CRNN implements the configuration the reference instantiates: "
                                      "CRNN(32, 1, 37, 256) (interfaces/base.py:310)")
        ks, ps, nm = [3, 3, 3, 3, 3, 3, 2], [1, 1, 1, 1, 1, 1, 0], [64, 128, 256, 256, 512, 512, 512]
        cnn = _Seq()
        for i in range(7):  # module names / order as crnn.py:36-63
            cnn.add_module(f"conv{i}", nn.Conv2d(nc if i == 0 else nm[i - 1], nm[i], ks[i], 1, ps[i]))
            if i in (2, 4, 6):
                cnn.add_module(f"batchnorm{i}", nn.BatchNorm2d(nm[i]))
            cnn.add_module(f"relu{i}", nn.ReLU(True))
            if i in (0, 1):
                cnn.add_module(f"pooling{i}", nn.MaxPool2d(2, 2))
            elif i in (3, 5):
                cnn.add_module(f"pooling{2 if i == 3 else 3}", nn.MaxPool2d((2, 2), (2, 1), (0, 1)))
        self.cnn = cnn
        self.rnn = _Seq(BidirectionalLSTM(512, nh, nh), BidirectionalLSTM(nh, nh, nclass))
        self._table = None
        self._ws = {}

    def _apply(self, fn, *a, **k):
        self._table, self._ws = None, {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._table = None
        return super().load_state_dict(*a, **k)

    def _params(self):
        if self._table is None:
            sd = self.state_dict(keep_vars=True)
            ts = list(sd.values())
            assert len(ts) == L.lib.focr_crnn_num_slots()
            for k, t in sd.items():
                if not t.is_cuda or not t.is_contiguous():
                    raise L.FocrError(f"CRNN parameter {k}: contiguous CUDA tensors required (no CPU path)")
            self._table = (ts, (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts]))
        return self._table[1]

    def _run(self, x: torch.Tensor, is_gray: bool) -> torch.Tensor:
        if self.training:
            raise L.FocrError("focr CRNN is inference-only (the reference keeps it frozen in eval mode); call .eval()")
        if not x.is_cuda:
            raise L.FocrError("focr CRNN runs on CUDA (sm_100a) only")
        want = (1, 32, 100) if is_gray else (3, 32, 128)
        if tuple(x.shape[1:]) != want:
            raise ValueError(f"expected (B,{want[0]},{want[1]},{want[2]}), got {tuple(x.shape)}")
        x = x.detach().contiguous().float()
        B = x.shape[0]
        ws = self._ws.get(B)
        if ws is None:
            ws = torch.empty(L.lib.focr_crnn_workspace_bytes(B), dtype=torch.uint8, device=x.device)
            self._ws = {B: ws}
        out = torch.empty(26, B, 37, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.lib.focr_crnn_forward(self._params(), x.data_ptr(), 1 if is_gray else 0, out.data_ptr(), B,
                                            ws.data_ptr(), ws.numel(), L.cur_stream()), "focr_crnn_forward")
        return out

    def forward(self, input):
        """(B,1,32,100) gray -> (26,B,37) logits, as the reference's CRNN.forward (crnn.py:70-80)."""
        return self._run(input, True)

    def forward_rgb(self, images_sr):
        """(B,3,32,128) SR output -> logits with parse_crnn_data (base.py:319-325) fused in."""
        return self._run(images_sr[:, :3], False)
