"""TSRN on the focr sm_100a engine — drop-in for ``model.tsrn.TSRN`` of the reference
(scene-text-telescope/model/tsrn.py:18-74; text-gestalt/model/tsrn.py is identical).

Same constructor signature and ``state_dict`` keys (239 entries with STN, e.g. ``block2.gru1.gru.weight_hh_l0_reverse``),
same forward contract.  The sequence residual block is conv/BN/mish/conv/BN followed by a vertical and a
horizontal bidirectional GRU (tsrn.py:77-98); here the GRU input projections run on the tcgen05 GEMM engine and the
recurrences in one persistent-warp kernel per block (csrc/gru.cu).  Parameter containers only: no torch compute.
"""
from __future__ import annotations

from torch import nn

from .tbsrn import _Holder, _Seq, _SREngineModule, _STNHead, _TPS, _Upsample

__all__ = ["TSRN"]


class _GruBlock(_Holder):  # tsrn.py:128-133
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, kernel_size=1, padding=0)
        self.gru = nn.GRU(cout, cout // 2, bidirectional=True, batch_first=True)


class _SRB(_Holder):  # RecurrentResidualBlock, tsrn.py:77-87 (module order = state_dict order)
    def __init__(self, ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, ch, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(ch)
        self.gru1 = _GruBlock(ch, ch)
        self.conv2 = nn.Conv2d(ch, ch, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(ch)
        self.gru2 = _GruBlock(ch, ch)


class TSRN(_SREngineModule):
    _ARCH = "tsrn"

    def __init__(self, scale_factor=2, width=128, height=32, STN=False, srb_nums=5, mask=False, hidden_units=32):
        super().__init__()
        if mask:
            raise NotImplementedError("focr TSRN: the 4-channel (mask=True) variant is not built "
                                      "(reference default mask=False, interfaces/base.py:141-142)")
        if scale_factor != 2 or hidden_units != 32 or (width, height) != (128, 32):
            raise NotImplementedError("focr TSRN supports scale_factor=2, hidden_units=32, 128x32 only")
        ch = 2 * hidden_units
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
