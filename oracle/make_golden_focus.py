"""Golden vectors for the stroke-focus loss (text-gestalt), produced by the UNMODIFIED reference classes
(loss/transformer_english_decomposition.py Transformer, loss/stroke_focus_loss.py StrokeFocusLoss.forward /
label_stroke_encoder) on CPU.  Run in the build container only (needs /root/reference).

Shims: torch .cuda() neutralised; StrokeFocusLoss.__init__ is bypassed because it opens git-ignored
assets (english_decomposition.txt, pretrain_transformer_stroke_decomposition.pth) - the instance is assembled by hand
from the same members, with synthetic weights and a synthetic decomposition table; forward() runs unmodified."""
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
sys.path.insert(0, str(REF / "text-gestalt"))

from oracle import synth, focus_oracle as FO  # noqa: E402


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")
    from loss import transformer_english_decomposition as reft
    from loss import stroke_focus_loss as refl
    torch.manual_seed(0)
    model = reft.Transformer().eval()
    spec = {k: list(v.shape) for k, v in model.state_dict().items()}
    gd = synth.GOLDEN_DIR
    (gd / "focus_spec.json").write_text(json.dumps(spec, indent=0))
    sd = FO.synth_recogniser_state_dict(spec)
    assert torch.equal(sd["pe.pe"], model.state_dict()["pe.pe"])
    model.load_state_dict(sd)
    # A trained recogniser's BatchNorm running statistics match its activations; random ones do not, and a deep ReLU
    # stack with mismatched statistics collapses to content-independent features (uniform attention maps, nothing to
    # test).  Calibrate them the way training would: one train-mode pass of the encoder over a calibration batch with
    # cumulative averaging (momentum=None), then freeze.  The resulting buffers are stored in the fixture so the tests
    # rebuild exactly this state dict from (spec, seed) + fixture.
    _, cal = synth.synth_images(8, seed=11)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
            m.momentum = None
    model.train()
    with torch.no_grad():
        model.encoder(FO.to_gray_tensor(cal))
    model.eval()
    bn_stats = {k: v.clone() for k, v in model.state_dict().items()
                if k.endswith("running_mean") or k.endswith("running_var")}
    sd.update(bn_stats)

    B = 2
    lr, hr = synth.synth_images(B, seed=5)
    # an early-training SR output: the blurry bilinear up-sample of the LR crop plus a little noise
    sr = torch.nn.functional.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False)
    sr = (sr + 0.02 * torch.randn(hr.shape, generator=torch.Generator().manual_seed(3))).clamp(0, 1)
    labels = ["ab3", "Hello"]
    dic = FO.synth_decomposition()

    loss_mod = object.__new__(refl.StrokeFocusLoss)           # bypass __init__ (asset files), keep forward()
    torch.nn.Module.__init__(loss_mod)
    loss_mod.args = types.SimpleNamespace(text_focus=True, stroke_lambda=50)
    loss_mod.mse_loss = torch.nn.MSELoss()
    loss_mod.ce_loss = torch.nn.CrossEntropyLoss()
    loss_mod.l1_loss = torch.nn.L1Loss()
    loss_mod.english_stroke_alphabet = "0123456789"
    loss_mod.english_stroke_dict = {c: i for i, c in enumerate("0123456789")}
    loss_mod.dic = dic
    loss_mod.transformer = model

    # ---- label encoding: reference vs restatement
    r_len, r_inp, r_gt = loss_mod.label_stroke_encoder(labels)
    o_len, o_inp, o_gt = FO.label_stroke_encoder(labels, dic)
    assert torch.equal(r_len, o_len) and torch.equal(r_inp, o_inp) and torch.equal(r_gt, o_gt)

    # ---- Transformer.forward: reference vs restatement (probs, attention map, correct list)
    with torch.no_grad():
        r_probs, r_map, r_corr = model(FO.to_gray_tensor(hr), r_len, r_inp, test=False)
        o_probs, o_map, o_corr = FO.transformer_forward(sd, FO.to_gray_tensor(hr), o_len, o_inp)
    assert torch.allclose(r_probs, o_probs, atol=2e-5, rtol=1e-4), (r_probs - o_probs).abs().max()
    assert torch.allclose(r_map, o_map, atol=1e-6, rtol=1e-4), (r_map - o_map).abs().max()
    assert r_corr == o_corr

    # ---- the loss and its gradient w.r.t. the SR image
    sr_r = sr.clone().requires_grad_(True)
    loss_r, mse_r, att_r, rec_r = loss_mod(sr_r, hr, labels)
    (loss_r * 100).backward()                                  # interfaces/super_resolution.py: loss_im = loss * 100
    sr_o = sr.clone().requires_grad_(True)
    loss_o, mse_o, att_o, info = FO.stroke_focus_loss(sd, sr_o, hr, labels, dic, 50.0)
    (loss_o * 100).backward()
    assert rec_r == -1
    assert abs(loss_r.item() - loss_o.item()) < 1e-6 * abs(loss_r.item()) + 1e-8
    assert abs(att_r.item() - att_o.item()) < 1e-5 * abs(att_r.item()) + 1e-10
    assert torch.allclose(sr_r.grad, sr_o.grad, atol=1e-7, rtol=1e-3), (sr_r.grad - sr_o.grad).abs().max()
    # gradient of the attention term alone (what the CUDA path adds to the MSE gradient)
    sr_a = sr.clone().requires_grad_(True)
    _, _, att_a, _ = FO.stroke_focus_loss(sd, sr_a, hr, labels, dic, 50.0)
    (att_a * 50.0 * 100).backward()

    # the BatchNorm-folded form of the restatement is the same function
    with torch.no_grad():
        m_fold = FO.attention_map(sd, FO.to_gray_tensor(hr), r_inp, FO.Numerics(fold=True))
    assert torch.allclose(m_fold, r_map, atol=1e-6, rtol=1e-3), (m_fold - r_map).abs().max()

    out = {"hr": hr, "sr": sr, "labels": labels, "length": r_len, "text_input": r_inp, "loss": loss_r.detach(),
           "mse": mse_r.detach(), "attention_loss": att_r.detach(), "map_hr": r_map, "map_sr": info["map_sr"].detach(),
           "probs_hr": r_probs, "correct_hr": r_corr, "d_sr_total_x100": sr_r.grad, "d_sr_attn_x100": sr_a.grad, "bn_stats": bn_stats}
    torch.save(out, gd / "focus_b2.pt")
    h = hashlib.sha256((gd / "focus_b2.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "focus_b2.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  focus_b2.pt"]) + "\n")
    print("focus golden: loss", loss_r.item(), "mse", mse_r.item(), "attention", att_r.item(), "T", r_inp.shape[1],
          "|d_sr attn|", sr_a.grad.norm().item(), "|d_sr total|", sr_r.grad.norm().item())


if __name__ == "__main__":
    main()
