"""Golden values of the stroke-level-decomposition recogniser recorded from the UNMODIFIED reference module
(/root/reference/stroke-level-decomposition/model/transformer.py) on synthetic weights, and the check that
oracle/sld_oracle.py reproduces it: logits, packed prediction, attention map, encoder features, loss, every parameter
gradient, running statistics, one Adadelta step.  Shims: stub lmdb / Levenshtein / IPython, neutralised .cuda(), chdir into the
subproject (util.py opens ./data/*.txt at import).  Dropout probabilities are set to 0 (parity mode)."""
import hashlib
import json
import os
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = "/root/reference/stroke-level-decomposition"
sys.path.insert(0, str(ROOT))
from oracle import sld_oracle as SO, synth  # noqa: E402


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
    model = Transformer("stroke")
    gd = synth.GOLDEN_DIR
    spec = {k: list(v.shape) for k, v in model.state_dict().items() if k != "pe.pe"}
    (gd / "sld_spec.json").write_text(json.dumps(spec))
    sd = synth.synth_state_dict(spec, 1234)
    model.load_state_dict(sd, strict=False)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    B = 3
    image, strings = SO.synth_batch(B)
    length, text_input, text_gt = SO.converter_stroke(strings)
    # the reference's own converter builds the same tensors (util.py:90-116) - checked through its stroke branch
    util.character_to_strokelist.update({f"c{i}": s[:-1] for i, s in enumerate(strings)})
    l2, ti2, tg2, _ = util.converter("stroke", [[f"c{i}"] for i in range(B)])
    assert torch.equal(l2, length) and torch.equal(ti2, text_input) and torch.equal(tg2, text_gt)

    model.train()
    out = model(image, length, text_input)
    loss = torch.nn.CrossEntropyLoss()(out["pred"], text_gt)
    model.zero_grad()
    loss.backward()
    ref_grads = {k: p.grad for k, p in model.named_parameters()}
    ref_sd_after = {k: v.clone() for k, v in model.state_dict().items()}

    # oracle on the same inputs
    osd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    stats = {}
    o_loss, o_logits, o_map, o_conv = SO.loss_fn(osd, image, length, text_input, text_gt, None, stats)
    o_loss.backward()
    assert torch.allclose(o_loss, loss, rtol=1e-5), (float(o_loss), float(loss))
    assert torch.allclose(SO.pack(o_logits, length), out["pred"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(o_map, out["map"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(o_conv, out["conv"], rtol=1e-4, atol=1e-5)
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
    # eval-mode forward with the test=True contract
    model.eval()
    with torch.no_grad():
        ev = model(image, length, text_input, test=True)
        sd_eval = {k: v for k, v in ref_sd_after.items()}
        e_logits, e_map, e_conv = SO.forward(sd_eval, image, text_input, train=False)
    assert torch.allclose(e_logits, ev["pred"], rtol=1e-4, atol=1e-5) and torch.allclose(e_map, ev["map"], rtol=1e-4, atol=1e-6)
    # one Adadelta step with the reference's optimiser settings (train.py:32-36)
    opt = torch.optim.Adadelta(model.parameters(), lr=1.0, rho=0.9)
    opt.step()
    upd = {}
    small = [k for k, g in ref_grads.items() if g is not None and g.numel() <= 2048]
    for k in small:
        p2, _, _ = SO.adadelta_update(sd[k], ref_grads[k], torch.zeros_like(sd[k]), torch.zeros_like(sd[k]))
        assert torch.allclose(p2, dict(model.named_parameters())[k].detach(), rtol=1e-5, atol=1e-7), k
        upd[k] = dict(model.named_parameters())[k].detach().clone()

    golden = {
        "B": B, "strings": strings, "length": length, "text_input": text_input, "text_gt": text_gt,
        "image_checksum": float(image.double().sum()),
        "loss": loss.detach(), "pred": out["pred"].detach(), "map": out["map"].detach(),
        "conv_sample": out["conv"].detach()[:, ::16, ::2, ::2].clone(), "conv_norm": out["conv"].detach().norm(),
        "eval_pred": ev["pred"], "eval_map": ev["map"],
        "grad_norms": {k: (g.norm() if g is not None else None) for k, g in ref_grads.items()},
        "grads_small": {k: ref_grads[k].clone() for k in small},
        "grad_samples": {k: g.reshape(-1)[::max(g.numel() // 4096, 1)][:4096].clone() for k, g in ref_grads.items()
                         if g is not None and g.numel() > 2048},
        "running_after": {k: v for k, v in ref_sd_after.items() if "running" in k and v.numel() <= 256},
        "adadelta_small": upd,
        "worst_oracle_vs_reference_grad_err": worst,
    }
    torch.save(golden, gd / "sld_b3.pt")
    h = hashlib.sha256((gd / "sld_b3.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "sld_b3.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  sld_b3.pt"]) + "\n")
    print(f"sld golden: loss {float(loss):.6f}, {len(ref_grads)} params, worst oracle-vs-reference gradient error {worst:.2e}, "
          f"{(gd / 'sld_b3.pt').stat().st_size / 1e6:.1f} MB")


def main_b32():
    """the same reference module at batch 32 (BatchNorm over 8192 positions per channel: the conditioning the GPU whole-step
    test needs); records the loss, every gradient norm and 256-element samples of every gradient"""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    Transformer, util = load_reference()
    model = Transformer("stroke")
    gd = synth.GOLDEN_DIR
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    model.load_state_dict(sd, strict=False)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    B = 32
    image, strings = SO.synth_batch(B)
    length, text_input, text_gt = SO.converter_stroke(strings)
    model.train()
    out = model(image, length, text_input)
    loss = torch.nn.CrossEntropyLoss()(out["pred"], text_gt)
    model.zero_grad()
    loss.backward()
    ref_grads = {k: p.grad for k, p in model.named_parameters()}
    golden = {
        "B": B, "strings": strings, "length": length, "text_input": text_input, "text_gt": text_gt,
        "image_checksum": float(image.double().sum()), "loss": loss.detach(), "pred_norm": out["pred"].detach().norm(),
        "conv_norm": out["conv"].detach().norm(),
        "grad_norms": {k: (g.norm() if g is not None else None) for k, g in ref_grads.items()},
        "grad_samples": {k: g.reshape(-1)[::max(g.numel() // 256, 1)][:256].clone() for k, g in ref_grads.items() if g is not None},
    }
    torch.save(golden, gd / "sld_b32.pt")
    h = hashlib.sha256((gd / "sld_b32.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "sld_b32.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  sld_b32.pt"]) + "\n")
    print(f"sld b32 golden: loss {float(loss):.6f}, {(gd / 'sld_b32.pt').stat().st_size / 1e6:.2f} MB")


def wide_batch(B):
    """(B, 3, 32, 320) crops: ten synthetic squares side by side (the same construction as tests/test_gpu_sld.py::_setup_wide)"""
    parts = [SO.synth_batch(B, seed=1234 + 17 * i) for i in range(10)]
    return torch.cat([p[0] for p in parts], dim=3), parts[0][1]


def main_w320():
    """the UNMODIFIED reference module on 32 x 320 crops (BASELINE configs[3]; its ResNet is fully convolutional,
    model/transformer.py:126-164, so the 16 x 160 map simply becomes 2 560 image tokens): loss, output norms, every gradient norm
    and gradient samples at batch 2 - the pin of oracle/sld_oracle.py at the shape the GPU tests of that configuration use"""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    Transformer, util = load_reference()
    model = Transformer("stroke")
    gd = synth.GOLDEN_DIR
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    model.load_state_dict(sd, strict=False)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    B = 2
    image, strings = wide_batch(B)
    length, text_input, text_gt = SO.converter_stroke(strings)
    model.train()
    out = model(image, length, text_input)
    assert out["conv"].shape == (B, 1024, 16, 160) and out["map"].shape[-1] == 2560
    loss = torch.nn.CrossEntropyLoss()(out["pred"], text_gt)
    model.zero_grad()
    loss.backward()
    ref_grads = {k: p.grad for k, p in model.named_parameters()}
    golden = {
        "B": B, "strings": strings, "length": length, "text_input": text_input, "text_gt": text_gt,
        "image_checksum": float(image.double().sum()), "loss": loss.detach(), "pred": out["pred"].detach(),
        "map_sample": out["map"].detach()[:, :, :, ::64].clone(), "conv_norm": out["conv"].detach().norm(),
        "grad_norms": {k: (g.norm() if g is not None else None) for k, g in ref_grads.items()},
        "grad_samples": {k: g.reshape(-1)[::max(g.numel() // 256, 1)][:256].clone() for k, g in ref_grads.items() if g is not None},
    }
    torch.save(golden, gd / "sld_w320_b2.pt")
    h = hashlib.sha256((gd / "sld_w320_b2.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "sld_w320_b2.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  sld_w320_b2.pt"]) + "\n")
    print(f"sld 32x320 golden: loss {float(loss):.6f}, {(gd / 'sld_w320_b2.pt').stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    if "--b32" in sys.argv:
        main_b32()
    elif "--w320" in sys.argv:
        main_w320()
    else:
        main()
