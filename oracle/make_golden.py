"""Generate tests/golden/* by running the UNMODIFIED reference modules (imported from
/root/reference) on CPU.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Shims: a stub IPython module, and `.cuda()` neutralised because
model/tbsrn.py:83 hard-codes `.cuda()` inside forward.  Nothing else is patched.
"""
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

from oracle import synth, tbsrn_oracle as O  # noqa: E402


def import_reference_stt():
    ipy = types.ModuleType("IPython")
    ipy.embed = lambda *a, **k: None
    sys.modules.setdefault("IPython", ipy)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    stt = str(REF / "scene-text-telescope")
    if stt not in sys.path:
        sys.path.insert(0, stt)
    import importlib
    return importlib.import_module("model.tbsrn")


def main():
    torch.manual_seed(1234)
    torch.set_num_threads(os.cpu_count())
    tb = import_reference_stt()
    model = tb.TBSRN(scale_factor=2, width=128, height=32, STN=True, srb_nums=5, mask=False, hidden_units=32)
    ref_sd = model.state_dict()
    spec = {k: list(v.shape) for k, v in ref_sd.items()}
    gd = synth.GOLDEN_DIR
    gd.mkdir(parents=True, exist_ok=True)
    (gd / "tbsrn_spec.json").write_text(json.dumps(spec, indent=0))

    # the oracle's TPS buffers must equal the reference's registered buffers
    tbuf = O.tps_buffers()
    for k, v in tbuf.items():
        assert torch.allclose(v, ref_sd[k], atol=1e-6, rtol=1e-5), k
    # positional encoding restatement
    assert torch.equal(O.positionalencoding2d(64, 16, 64), tb.positionalencoding2d(64, 16, 64))

    sd = synth.synth_state_dict(spec, seed=1234, computed=tbuf)
    model.load_state_dict(sd)
    out = {}
    B = 4
    lr, hr = synth.synth_images(B, seed=1234)

    # ---- eval forward (STN off in eval, BN running stats, dropout off) ----------------------
    model.eval()
    with torch.no_grad():
        out["eval_sr"] = model(lr).clone()

    # ---- train-mode forward/backward with dropout disabled (p=0) but BN in train mode --------
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.load_state_dict(sd)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999))
    sr = model(lr)
    loss = torch.nn.functional.mse_loss(sr, hr)
    opt.zero_grad()
    (loss * 100).backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    no_grad = sorted(k for k, p in model.named_parameters() if p.grad is None)
    gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 0.25)
    opt.step()
    new_sd = {k: v.clone() for k, v in model.state_dict().items()}

    out["train_sr"] = sr.detach().clone()
    out["train_mse"] = loss.detach().clone()
    out["grad_norm"] = gnorm.detach().clone()
    out["grad_l2"] = {k: g.norm().item() for k, g in grads.items()}
    out["no_grad_params"] = no_grad
    keep = ["block1.0.weight", "block1.1.weight", "block2.conv1.weight", "block2.bn1.weight",
            "block2.feature_enhancer.multihead.linears.0.weight", "block2.feature_enhancer.mul_layernorm1.a_2",
            "block2.feature_enhancer.pff.w_1.weight", "block2.feature_enhancer.linear.weight",
            "block6.conv2.weight", "block6.feature_enhancer.multihead.linears.3.bias",
            "block7.0.weight", "block7.1.bias", "block8.0.conv.weight", "block8.1.weight", "block8.1.bias",
            "stn_head.stn_convnet.0.0.weight", "stn_head.stn_convnet.2.0.weight", "stn_head.stn_fc1.0.bias",
            "stn_head.stn_fc2.weight", "stn_head.stn_fc2.bias"]
    out["grads"] = {k: grads[k] for k in keep}
    out["new_param_sum"] = {k: (new_sd[k].double().sum().item(), new_sd[k].double().abs().sum().item())
                            for k in new_sd if new_sd[k].is_floating_point()}
    out["new_running"] = {k: new_sd[k] for k in new_sd if "running_" in k and ("block2" in k or "block7" in k
                                                                                or "stn_convnet.0." in k
                                                                                or "stn_fc1" in k)}
    # ---- validate the restatement against the reference right here ---------------------------
    with torch.no_grad():
        o_eval = O.tbsrn_forward(sd, lr, training=False)
    assert torch.allclose(o_eval, out["eval_sr"], atol=2e-5, rtol=1e-4), (o_eval - out["eval_sr"]).abs().max()
    st = {}
    o_sd, info = O.train_step(sd, lr, hr, st, masks=None)
    assert torch.allclose(info["sr"], out["train_sr"], atol=2e-5, rtol=1e-4)
    assert abs(info["grad_norm"].item() - gnorm.item()) < 1e-4 * gnorm.item()
    for k, g in grads.items():
        assert torch.allclose(info["grads"][k], g, atol=1e-5 + 1e-4 * g.abs().max().item(), rtol=1e-3), k
    assert sorted(set(k for k, _ in model.named_parameters()) - set(info["grads"])) == no_grad
    for k in new_sd:
        if new_sd[k].is_floating_point():
            assert torch.allclose(o_sd[k], new_sd[k], atol=2e-6, rtol=1e-4), k
    torch.save(out, gd / "tbsrn_b4.pt")
    h = hashlib.sha256((gd / "tbsrn_b4.pt").read_bytes()).hexdigest()
    (gd / "SHA256SUMS").write_text(f"{h}  tbsrn_b4.pt\n")
    print("golden written:", gd / "tbsrn_b4.pt", os.path.getsize(gd / "tbsrn_b4.pt"), "bytes; mse", loss.item(),
          "gnorm", gnorm.item())


if __name__ == "__main__":
    main()
