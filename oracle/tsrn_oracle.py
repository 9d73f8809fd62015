"""ORACLE — test infrastructure only.  Torch-fp32 restatement of the reference's TSRN training path
(scene-text-telescope/model/tsrn.py; text-gestalt's copy is identical), pinned to the real module by
oracle/make_golden_tsrn.py -> tests/golden/tsrn_b4.pt.  Trunk pieces shared with TBSRN come from tbsrn_oracle."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import tbsrn_oracle as O

Tensor = torch.Tensor


def _gru_dir(x: Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> Tensor:
    """nn.GRU single direction, batch_first, gate order r, z, n:  x (N,T,In) -> (N,T,H)."""
    N, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(N, H)
    xp = F.linear(x, w_ih, b_ih)  # (N,T,3H)
    out = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        hp = F.linear(h, w_hh, b_hh)
        r = torch.sigmoid(xp[:, t, :H] + hp[:, :H])
        z = torch.sigmoid(xp[:, t, H:2 * H] + hp[:, H:2 * H])
        n = torch.tanh(xp[:, t, 2 * H:] + r * hp[:, 2 * H:])
        h = (1 - z) * n + z * h
        out[t] = h
    return torch.stack(out, 1)


def gru_block(sd, pre: str, x: Tensor) -> Tensor:
    """GruBlock.forward, tsrn.py:135-145: conv1x1, then a BiGRU along the LAST spatial axis of x (b, c, d2, d3)."""
    x = F.conv2d(x, sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"])
    x = x.permute(0, 2, 3, 1).contiguous()
    b = x.shape
    seq = x.view(b[0] * b[1], b[2], b[3])
    p = pre + ".gru."
    fwd = _gru_dir(seq, sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"], sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"], False)
    bwd = _gru_dir(seq, sd[p + "weight_ih_l0_reverse"], sd[p + "weight_hh_l0_reverse"], sd[p + "bias_ih_l0_reverse"],
                   sd[p + "bias_hh_l0_reverse"], True)
    y = torch.cat([fwd, bwd], 2).view(b[0], b[1], b[2], b[3])
    return y.permute(0, 3, 1, 2).contiguous()


def srb(sd, pre, x, new_stats, training, taps=None):
    """RecurrentResidualBlock.forward, tsrn.py:89-98."""
    r = F.conv2d(x, sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], padding=1)
    r = O.mish(O.batch_norm_train(r, sd, pre + ".bn1", new_stats, training))
    r = F.conv2d(r, sd[pre + ".conv2.weight"], sd[pre + ".conv2.bias"], padding=1)
    r = O.batch_norm_train(r, sd, pre + ".bn2", new_stats, training)
    r1 = gru_block(sd, pre + ".gru1", r.transpose(-1, -2).contiguous()).transpose(-1, -2).contiguous()
    out = gru_block(sd, pre + ".gru2", x + r1).contiguous()
    if taps is not None:
        taps.update({pre + ".r0": r.detach(), pre + ".o1": r1.detach(), pre + ".out": out.detach()})
    return out


def tsrn_forward(sd: Dict[str, Tensor], x: Tensor, training: bool = True, stn: bool = True, srb_nums: int = 5,
                 new_stats: Optional[dict] = None, taps: Optional[dict] = None) -> Tensor:
    """TSRN.forward, tsrn.py:61-74."""
    if stn and training:
        ctrl = O.stn_head(sd, x, new_stats, training, taps=taps)
        x = O.tps_transform(sd, x, ctrl)
    b1 = F.prelu(F.conv2d(x, sd["block1.0.weight"], sd["block1.0.bias"], padding=4), sd["block1.1.weight"])
    cur = b1
    for i in range(srb_nums):
        cur = srb(sd, f"block{i + 2}", cur, new_stats, training, taps)
    k = srb_nums + 2
    cur = O.batch_norm_train(F.conv2d(cur, sd[f"block{k}.0.weight"], sd[f"block{k}.0.bias"], padding=1), sd,
                             f"block{k}.1", new_stats, training)
    k = srb_nums + 3
    u = O.mish(F.pixel_shuffle(F.conv2d(b1 + cur, sd[f"block{k}.0.conv.weight"], sd[f"block{k}.0.conv.bias"], padding=1), 2))
    out = F.conv2d(u, sd[f"block{k}.1.weight"], sd[f"block{k}.1.bias"], padding=4)
    if taps is not None:
        taps.update({"b1": b1.detach(), "opre": out.detach()})
    return torch.tanh(out)


def train_step(sd, lr_img, hr_img, opt_state, stn=True, srb_nums=5, taps=None):
    """step body of TextSR.train with the MSE criterion (interfaces/super_resolution.py:60-84) on TSRN"""
    float_keys = [k for k, v in sd.items() if v.is_floating_point() and not O.is_buffer(k)]
    leaf = {k: (sd[k].detach().clone().requires_grad_(True) if k in float_keys else sd[k]) for k in sd}
    new_stats: dict = {}
    sr = tsrn_forward(leaf, lr_img, training=True, stn=stn, srb_nums=srb_nums, new_stats=new_stats, taps=taps)
    mse = F.mse_loss(sr, hr_img)
    (mse * 100).backward()
    grads = {k: leaf[k].grad for k in float_keys if leaf[k].grad is not None}
    clipped, gnorm = O.clip_grad_norm(grads)
    new_params = O.adam_step({k: sd[k] for k in float_keys}, clipped, opt_state)
    new_sd = dict(sd)
    new_sd.update(new_params)
    new_sd.update(new_stats)
    return new_sd, {"sr": sr.detach(), "mse": mse.detach(), "grad_norm": gnorm, "grads": grads}
