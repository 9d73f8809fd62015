"""Golden vectors for the text-focus loss (scene-text-telescope), produced by the UNMODIFIED reference modules
(loss/transformer.py Transformer, loss/weight_ce_loss.py, loss/text_focus_loss.py TextFocusLoss.forward / label_encoder) on CPU.
Run in the build container only (needs /root/reference).

Shims: torch .cuda() neutralised; weight_ce_loss.py opens ./dataset/mydata/confuse.pkl at import, so the
script chdir's into a temp dir holding a synthetic 62x62 pickle; TextFocusLoss.__init__ is bypassed (it loads the git-ignored
pretrain_transformer.pth) and the instance is assembled from the same members with synthetic weights; forward() runs unmodified."""
from __future__ import annotations

import hashlib
import json
import os
import pickle
import sys
import tempfile
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("FOCR_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "scene-text-telescope"))

from oracle import synth, focus_oracle as FO, textfocus_oracle as TF  # noqa: E402


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    counts = TF.synth_confuse_counts()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "dataset", "mydata"))
    with open(os.path.join(tmp, "dataset", "mydata", "confuse.pkl"), "wb") as f:
        pickle.dump(counts, f)
    os.chdir(tmp)
    from loss import transformer as reft
    from loss import weight_ce_loss as refw
    from loss import text_focus_loss as refl
    table_o = TF.confuse_weight_table(counts)
    assert torch.allclose(refw.weight_table, table_o), (refw.weight_table - table_o).abs().max()

    torch.manual_seed(0)
    model = reft.Transformer().eval()
    spec = {k: list(v.shape) for k, v in model.state_dict().items()}
    gd = synth.GOLDEN_DIR
    (gd / "textfocus_spec.json").write_text(json.dumps(spec, indent=0))
    sd = FO.synth_recogniser_state_dict(spec, seed=778)
    model.load_state_dict(sd)
    # calibrated BatchNorm statistics, as in make_golden_focus.py
    _, cal = synth.synth_images(8, seed=11)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
            m.momentum = None
    model.train()
    with torch.no_grad():
        model.encoder(FO.to_gray_tensor(cal))
    model.eval()
    bn_stats = {k: v.clone() for k, v in model.state_dict().items() if k.endswith("running_mean") or k.endswith("running_var")}
    sd.update(bn_stats)

    B = 3
    lr, hr = synth.synth_images(B, seed=6)
    sr = torch.nn.functional.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False)
    sr = (sr + 0.02 * torch.randn(hr.shape, generator=torch.Generator().manual_seed(4))).clamp(0, 1)
    labels = ["B200", "text-Focus!", "a"]

    loss_mod = object.__new__(refl.TextFocusLoss)
    torch.nn.Module.__init__(loss_mod)
    loss_mod.args = types.SimpleNamespace(text_focus=True)
    loss_mod.mse_loss = torch.nn.MSELoss()
    loss_mod.ce_loss = torch.nn.CrossEntropyLoss()
    loss_mod.l1_loss = torch.nn.L1Loss()
    loss_mod.english_alphabet = TF.LABEL_ALPHABET
    loss_mod.english_dict = {c: i for i, c in enumerate(TF.LABEL_ALPHABET)}
    loss_mod.transformer = model

    filt = [refl.str_filt(s, "lower") + "-" for s in labels]
    assert filt == [TF.str_filt(s, "lower") + "-" for s in labels]
    r_len, r_inp, r_gt = loss_mod.label_encoder(filt)
    o_len, o_inp, o_gt = TF.label_encoder(filt)
    assert torch.equal(r_len, o_len) and torch.equal(r_inp, o_inp) and torch.equal(r_gt, o_gt)

    sr_r = sr.clone().requires_grad_(True)
    loss_r, mse_r, att_r, rec_r = loss_mod(sr_r, hr, labels)
    (loss_r * 100).backward()
    sr_o = sr.clone().requires_grad_(True)
    loss_o, mse_o, att_o, rec_o, info = TF.text_focus_loss(sd, sr_o, hr, labels, table_o)
    (loss_o * 100).backward()
    assert abs(loss_r.item() - loss_o.item()) < 1e-5 * abs(loss_r.item())
    assert abs(att_r.item() - att_o.item()) < 1e-4 * abs(att_r.item())
    assert abs(rec_r.item() - rec_o.item()) < 1e-5 * abs(rec_r.item())
    rel = ((sr_r.grad - sr_o.grad).norm() / sr_r.grad.norm()).item()
    assert rel < 2e-2, rel            # chaotic recogniser + sign(): fp32 summation order alone moves the gradient by ~1e-3..1e-2
    # recognition term alone
    sr_c = sr.clone().requires_grad_(True)
    _, _, _, rec_c, _ = TF.text_focus_loss(sd, sr_c, hr, labels, table_o)
    (rec_c * 0.0005 * 100).backward()

    out = {"hr": hr, "sr": sr, "labels": labels, "length": r_len, "text_input": r_inp, "text_gt": r_gt,
           "confuse_counts": torch.tensor(counts), "weight_table": refw.weight_table.clone(),
           "loss": loss_r.detach(), "mse": mse_r.detach(), "attention_loss": att_r.detach(), "recognition_loss": rec_r.detach(),
           "map_hr": info["map_hr"].detach(), "map_sr": info["map_sr"].detach(), "sr_pred": info["sr_pred"].detach(),
           "d_sr_total_x100": sr_r.grad, "d_sr_ce_x100": sr_c.grad, "bn_stats": bn_stats}
    torch.save(out, gd / "textfocus_b3.pt")
    h = hashlib.sha256((gd / "textfocus_b3.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "textfocus_b3.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  textfocus_b3.pt"]) + "\n")
    print("textfocus golden: loss", loss_r.item(), "mse", mse_r.item(), "attention", att_r.item(), "recognition", rec_r.item(),
          "T", r_inp.shape[1], "|d total|", sr_r.grad.norm().item(), "|d ce|", sr_c.grad.norm().item(), "oracle-vs-ref grad rel", rel)


if __name__ == "__main__":
    main()
