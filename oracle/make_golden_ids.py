"""Golden values of the image-ids-CTR recogniser recorded from the UNMODIFIED reference module
(/root/reference/image-ids-CTR/model/transformer.py) and the step body of train.py:63-80 on synthetic weights / stand-in text
features, and the check that oracle/ids_oracle.py reproduces them.  Shims: stub lmdb, neutralised .cuda(), chdir into the
subproject (util.py opens ./data/*.txt at import); dropout probabilities set to 0 (parity mode)."""
import hashlib
import json
import os
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = "/root/reference/image-ids-CTR"
sys.path.insert(0, str(ROOT))
from oracle import ids_oracle as IO, sld_oracle as SO, synth  # noqa: E402


def load_reference():
    for name in ("lmdb", "Levenshtein", "IPython"):
        sys.modules.setdefault(name, types.ModuleType(name))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    try:
        from model.transformer import Transformer
        import util
    finally:
        os.chdir(cwd)
    return Transformer, util


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    Transformer, util = load_reference()
    model = Transformer()
    assert model.word_n_class == IO.N_CLASS
    gd = synth.GOLDEN_DIR
    spec = {k: list(v.shape) for k, v in model.state_dict().items() if k != "pe.pe"}
    (gd / "ids_spec.json").write_text(json.dumps(spec))
    sd = synth.synth_state_dict(spec, 4321)
    model.load_state_dict(sd, strict=False)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    B = 4
    image, labels = IO.synth_batch(B)
    length, text_input, text_gt = IO.converter(labels)
    # the reference's converter on the corresponding character strings gives the same tensors (util.py:101-127)
    strings = ["".join(util.alphabet_character[i] for i in s[:-1]) + "#" for s in labels]
    l2, ti2, tg2, _ = util.converter(strings)
    assert torch.equal(l2, length) and torch.equal(ti2, text_input) and torch.equal(tg2, text_gt)
    text_features = IO.synth_text_features()

    model.train()
    result = model(image, length, text_input)                       # train.py:71-80, verbatim arithmetic
    reg = torch.cat([text_features[item].unsqueeze(0) for item in text_gt], dim=0)
    text_pred = result["pred"]
    text_pred = text_pred / text_pred.norm(dim=1, keepdim=True)
    final_res = text_pred @ text_features.t()
    loss_rec = torch.nn.CrossEntropyLoss()(final_res, text_gt)
    loss_dis = -torch.nn.MSELoss()(text_pred, reg)
    loss = loss_rec + 0.001 * loss_dis
    model.zero_grad()
    loss.backward()
    ref_grads = {k: p.grad for k, p in model.named_parameters()}
    ref_sd_after = {k: v.clone() for k, v in model.state_dict().items()}

    osd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    stats = {}
    o_loss, o_rec, o_dis, o_pred, o_map, o_conv = IO.loss_fn(osd, image, length, text_input, text_gt, text_features, stats)
    o_loss.backward()
    assert torch.allclose(o_loss, loss, rtol=1e-5) and torch.allclose(o_rec, loss_rec, rtol=1e-5) and torch.allclose(o_dis, loss_dis, rtol=1e-5)
    assert torch.allclose(SO.pack(o_pred, length), result["pred"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(o_map, result["map"], rtol=1e-4, atol=1e-6) and torch.allclose(o_conv, result["conv"], rtol=1e-4, atol=1e-5)
    worst = 0.0
    for k, g in ref_grads.items():
        if g is None:
            assert osd[k].grad is None, k
            continue
        e = ((osd[k].grad - g).norm() / (g.norm() + 1e-12)).item()
        worst = max(worst, e)
        assert e < 2e-3, (k, e)
    for k, v in stats.items():
        assert torch.allclose(v, ref_sd_after[k], rtol=1e-4, atol=1e-6), k
    model.eval()
    with torch.no_grad():
        ev = model(image, length, text_input, test=True)
        e_pred, e_map, _ = IO.forward(ref_sd_after, image, text_input, train=False)
    assert torch.allclose(e_pred, ev["pred"], rtol=1e-4, atol=1e-5)
    # one optimiser step with the reference's settings (train.py:28)
    opt = torch.optim.Adadelta(model.parameters(), lr=1.0, rho=0.9, weight_decay=1e-4)
    opt.step()
    small = [k for k, g in ref_grads.items() if g is not None and g.numel() <= 2048]
    upd = {}
    for k in small:
        p2, _, _ = SO.adadelta_update(sd[k], ref_grads[k], torch.zeros_like(sd[k]), torch.zeros_like(sd[k]), wd=1e-4)
        assert torch.allclose(p2, dict(model.named_parameters())[k].detach(), rtol=1e-5, atol=1e-7), k
        upd[k] = dict(model.named_parameters())[k].detach().clone()
    golden = {
        "B": B, "labels": labels, "length": length, "text_input": text_input, "text_gt": text_gt,
        "image_checksum": float(image.double().sum()), "text_features_checksum": float(text_features.double().sum()),
        "loss": loss.detach(), "loss_rec": loss_rec.detach(), "loss_dis": loss_dis.detach(),
        "pred_sample": result["pred"].detach()[:, ::8].clone(), "map": result["map"].detach(),
        "conv_sample": result["conv"].detach()[:, ::16].clone(), "eval_pred_sample": ev["pred"][:, :, ::8].clone(),
        "grad_norms": {k: (g.norm() if g is not None else None) for k, g in ref_grads.items()},
        "grads_small": {k: ref_grads[k].clone() for k in small},
        "grad_samples": {k: g.reshape(-1)[::max(g.numel() // 4096, 1)][:4096].clone() for k, g in ref_grads.items()
                         if g is not None and g.numel() > 2048},
        "running_after": {k: v for k, v in ref_sd_after.items() if "running" in k and v.numel() <= 256 and "layer4" not in k},
        "adadelta_small": upd, "worst_oracle_vs_reference_grad_err": worst,
    }
    torch.save(golden, gd / "ids_b4.pt")
    h = hashlib.sha256((gd / "ids_b4.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "ids_b4.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  ids_b4.pt"]) + "\n")
    print(f"ids golden: loss {float(loss.detach()):.6f} (rec {float(loss_rec.detach()):.6f}, dis {float(loss_dis.detach()):.6f}), "
          f"worst oracle-vs-reference gradient error {worst:.2e}, {(gd / 'ids_b4.pt').stat().st_size / 1e6:.1f} MB")


def main_b32():
    """the same reference module and train.py:71-80 arithmetic at batch 32: loss terms, every gradient norm, 256-element samples"""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    Transformer, util = load_reference()
    model = Transformer()
    gd = synth.GOLDEN_DIR
    sd = synth.synth_state_dict(synth.load_spec("ids"), 4321)
    model.load_state_dict(sd, strict=False)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    B = 32
    image, labels = IO.synth_batch(B)
    length, text_input, text_gt = IO.converter(labels)
    text_features = IO.synth_text_features()
    model.train()
    result = model(image, length, text_input)
    reg = torch.cat([text_features[item].unsqueeze(0) for item in text_gt], dim=0)
    text_pred = result["pred"]
    text_pred = text_pred / text_pred.norm(dim=1, keepdim=True)
    final_res = text_pred @ text_features.t()
    loss_rec = torch.nn.CrossEntropyLoss()(final_res, text_gt)
    loss_dis = -torch.nn.MSELoss()(text_pred, reg)
    loss = loss_rec + 0.001 * loss_dis
    model.zero_grad()
    loss.backward()
    ref_grads = {k: p.grad for k, p in model.named_parameters()}
    golden = {
        "B": B, "labels": labels, "length": length, "text_input": text_input, "text_gt": text_gt,
        "image_checksum": float(image.double().sum()), "loss": loss.detach(), "loss_rec": loss_rec.detach(),
        "loss_dis": loss_dis.detach(),
        "grad_norms": {k: (g.norm() if g is not None else None) for k, g in ref_grads.items()},
        "grad_samples": {k: g.reshape(-1)[::max(g.numel() // 256, 1)][:256].clone() for k, g in ref_grads.items() if g is not None},
    }
    torch.save(golden, gd / "ids_b32.pt")
    h = hashlib.sha256((gd / "ids_b32.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "ids_b32.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  ids_b32.pt"]) + "\n")
    print(f"ids b32 golden: loss {float(loss.detach()):.6f}, {(gd / 'ids_b32.pt').stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    if "--b32" in sys.argv:
        main_b32()
    else:
        main()
