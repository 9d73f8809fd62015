"""Golden vectors for TSRN from the UNMODIFIED reference module (scene-text-telescope/model/tsrn.py) on CPU."""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("FOCR_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "scene-text-telescope"))
from oracle import synth, tbsrn_oracle as O, tsrn_oracle as TS  # noqa: E402


def main():
    ipy = types.ModuleType("IPython")
    ipy.embed = lambda *a, **k: None
    sys.modules.setdefault("IPython", ipy)
    from model import tsrn as ref
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    model = ref.TSRN(scale_factor=2, width=128, height=32, STN=True, srb_nums=5, mask=False, hidden_units=32)
    spec = {k: list(v.shape) for k, v in model.state_dict().items()}
    gd = synth.GOLDEN_DIR
    (gd / "tsrn_spec.json").write_text(json.dumps(spec, indent=0))
    sd = synth.synth_state_dict(spec, seed=2468, computed=O.tps_buffers())
    model.load_state_dict(sd)
    B = 4
    lr, hr = synth.synth_images(B, seed=1234)
    out = {}
    model.eval()
    with torch.no_grad():
        out["eval_sr"] = model(lr).clone()
        assert torch.allclose(TS.tsrn_forward(sd, lr, training=False), out["eval_sr"], atol=2e-5, rtol=1e-4)
    model.train()
    model.load_state_dict(sd)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999))
    sr = model(lr)
    loss = torch.nn.functional.mse_loss(sr, hr)
    opt.zero_grad()
    (loss * 100).backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    assert len(grads) == len(list(model.parameters()))  # TSRN has no dead parameters
    gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 0.25)
    opt.step()
    new_sd = {k: v.clone() for k, v in model.state_dict().items()}
    o_sd, info = TS.train_step(sd, lr, hr, {})
    assert torch.allclose(info["sr"], sr.detach(), atol=2e-5, rtol=1e-4)
    assert abs(info["grad_norm"].item() - gnorm.item()) < 1e-4 * gnorm.item()
    for k, g in grads.items():
        assert torch.allclose(info["grads"][k], g, atol=1e-5 + 1e-4 * g.abs().max().item(), rtol=1e-3), k
    gmax = max(g.norm().item() for g in grads.values())
    for k in new_sd:
        if not new_sd[k].is_floating_point():
            continue
        if k in grads and grads[k].norm().item() < 1e-5 * gmax:
            # numerically-zero gradient (conv bias in front of a batch-stat BN): Adam turns rounding noise into +-lr
            assert (o_sd[k] - new_sd[k]).abs().max() <= 2.5e-4, k
            continue
        # Adam's first step is lr*g/(|g|+eps): elements whose gradient is ~0 flip sign under 1e-4-relative gradient
        # noise (the oracle GRU is a loop, the reference a fused op), so require agreement on 99.5 % of the elements
        # (the Adam restatement itself is pinned bit-for-bit by the TBSRN golden)
        bad = ((o_sd[k] - new_sd[k]).abs() > 2e-6 + 1e-4 * new_sd[k].abs()).float().mean().item()
        assert bad < 5e-3, (k, bad)
    out["train_sr"] = sr.detach().clone()
    out["train_mse"] = loss.detach().clone()
    out["grad_norm"] = gnorm.detach().clone()
    out["grad_l2"] = {k: g.norm().item() for k, g in grads.items()}
    keep = ["block2.gru1.gru.weight_hh_l0", "block2.gru1.gru.weight_ih_l0_reverse", "block2.gru2.gru.bias_hh_l0",
            "block2.gru2.conv1.weight", "block6.gru1.gru.bias_ih_l0_reverse", "block6.gru2.gru.weight_hh_l0_reverse",
            "block2.conv1.weight", "block1.0.weight"]
    out["grads"] = {k: grads[k] for k in keep}
    torch.save(out, gd / "tsrn_b4.pt")
    h = hashlib.sha256((gd / "tsrn_b4.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "tsrn_b4.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  tsrn_b4.pt"]) + "\n")
    print("tsrn golden: mse", loss.item(), "gnorm", gnorm.item(), os.path.getsize(gd / "tsrn_b4.pt"), "bytes")


if __name__ == "__main__":
    main()
