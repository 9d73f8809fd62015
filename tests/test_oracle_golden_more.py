"""CPU: the CRNN and TSRN oracle restatements vs golden vectors recorded from the real reference modules."""
import hashlib

import torch

from oracle import crnn_oracle as C
from oracle import synth, tbsrn_oracle as O, tsrn_oracle as TS


def _load(name):
    path = synth.GOLDEN_DIR / name
    sums = dict(reversed(ln.split()) for ln in (synth.GOLDEN_DIR / "SHA256SUMS").read_text().splitlines() if ln.strip())
    assert hashlib.sha256(path.read_bytes()).hexdigest() == sums[name]
    return torch.load(path, weights_only=False)


def test_crnn_oracle_vs_reference_golden():
    g = _load("crnn_b2.pt")
    sd = synth.synth_state_dict(synth.load_spec("crnn"), seed=4321)
    assert len(sd) == 49
    gray = C.parse_crnn_data(g["sr"])
    assert torch.equal(gray, g["gray"])
    with torch.no_grad():
        logits = C.crnn_forward(sd, gray)
    assert torch.allclose(logits, g["logits"], atol=2e-5, rtol=1e-4)
    assert torch.equal(C.greedy_path(logits), g["path"])
    assert C.get_crnn_pred(logits.permute(1, 0, 2)) == g["strings"]
    assert ["".join(C.ALPHABET[i] for i in C.ctc_collapse(p.tolist())) for p in g["path"]] == g["strings"]


def test_ctc_collapse_cases():
    assert C.ctc_collapse([0, 0, 0]) == []
    assert C.ctc_collapse([3, 3, 3]) == [3]
    assert C.ctc_collapse([3, 0, 3]) == [3, 3]
    assert C.ctc_collapse([1, 1, 0, 0, 2, 2, 1]) == [1, 2, 1]


def test_tsrn_oracle_vs_reference_golden():
    g = _load("tsrn_b4.pt")
    spec = synth.load_spec("tsrn")
    assert len(spec) == 239
    sd = synth.synth_state_dict(spec, seed=2468, computed=O.tps_buffers())
    lr, hr = synth.synth_images(4)
    with torch.no_grad():
        assert torch.allclose(TS.tsrn_forward(sd, lr, training=False), g["eval_sr"], atol=2e-5, rtol=1e-4)
    _, info = TS.train_step(sd, lr, hr, {})
    assert torch.allclose(info["sr"], g["train_sr"], atol=2e-5, rtol=1e-4)
    assert abs(info["grad_norm"].item() - g["grad_norm"].item()) < 1e-4 * g["grad_norm"].item()
    for k, v in g["grads"].items():
        assert torch.allclose(info["grads"][k], v, atol=1e-5 + 1e-4 * v.abs().max().item(), rtol=1e-3), k
    gmax = max(g["grad_l2"].values())
    for k, n in g["grad_l2"].items():
        if n < 1e-5 * gmax:  # numerically-zero gradients (conv bias before a batch-stat BN) are rounding noise
            assert info["grads"][k].norm().item() < 1e-4 * gmax, k
        else:
            assert abs(info["grads"][k].norm().item() - n) <= 1e-3 * n + 1e-6, k


def focus_state_dict(g):
    """(spec, seed 777) synthetic recogniser + the calibrated BatchNorm statistics stored in the fixture"""
    from oracle import focus_oracle as FO
    sd = FO.synth_recogniser_state_dict(synth.load_spec("focus"), g["bn_stats"])
    return sd


def test_focus_oracle_vs_reference_golden():
    """stroke-focus loss (text-gestalt): restatement vs outputs of the unmodified reference classes"""
    from oracle import focus_oracle as FO
    g = _load("focus_b2.pt")
    sd = focus_state_dict(g)
    assert len(sd) == 244
    dic = FO.synth_decomposition()
    ln, inp, _ = FO.label_stroke_encoder(g["labels"], dic)
    assert torch.equal(ln, g["length"]) and torch.equal(inp, g["text_input"])
    sr = g["sr"].clone().requires_grad_(True)
    loss, mse, att, info = FO.stroke_focus_loss(sd, sr, g["hr"], g["labels"], dic, 50.0)
    (loss * 100).backward()
    assert torch.allclose(info["map_hr"], g["map_hr"], atol=1e-6, rtol=1e-3)
    assert torch.allclose(info["map_sr"], g["map_sr"], atol=1e-6, rtol=1e-3)
    assert abs(att.item() - g["attention_loss"].item()) < 1e-4 * g["attention_loss"].item()
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * g["loss"].item()
    d = g["d_sr_total_x100"]
    assert ((sr.grad - d).norm() / d.norm()).item() < 2e-3


def test_textfocus_oracle_vs_reference_golden():
    """text-focus loss (scene-text-telescope): restatement vs outputs of the unmodified reference modules"""
    from oracle import focus_oracle as FO, textfocus_oracle as TF
    g = _load("textfocus_b3.pt")
    sd = FO.synth_recogniser_state_dict(synth.load_spec("textfocus"), g["bn_stats"], seed=778)
    table = TF.confuse_weight_table(g["confuse_counts"].numpy())
    assert torch.equal(table, g["weight_table"])
    sr = g["sr"].clone().requires_grad_(True)
    loss, mse, att, rec, info = TF.text_focus_loss(sd, sr, g["hr"], g["labels"], table)
    (loss * 100).backward()
    assert torch.equal(info["text_input"], g["text_input"]) and torch.equal(info["text_gt"], g["text_gt"])
    assert torch.allclose(info["sr_pred"], g["sr_pred"], atol=1e-4, rtol=1e-3)
    assert abs(rec.item() - g["recognition_loss"].item()) < 1e-4 * g["recognition_loss"].item()
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * g["loss"].item()
    d = g["d_sr_total_x100"]
    assert ((sr.grad - d).norm() / d.norm()).item() < 2e-2
    # the log-sum-exp form the CUDA kernel uses is the same function
    lw = torch.log(table[g["text_gt"]])
    lse = torch.logsumexp(g["sr_pred"] + lw, 1) - (g["sr_pred"] + lw).gather(1, g["text_gt"][:, None])[:, 0]
    assert abs(lse.mean().item() - g["recognition_loss"].item()) < 1e-5 * g["recognition_loss"].item()


def test_metrics_oracle_vs_reference_golden():
    """PSNR / SSIM restatement vs values recorded from the unmodified utils/ssim_psnr.py"""
    from oracle import metrics_oracle as MO
    g = _load("metrics.pt")
    for key, rec in g.items():
        sr, hr = MO.synth_pair(int(key[1:]), rec["seed"])
        assert abs(float(sr.double().sum() + hr.double().sum()) - rec["checksum"]) < 1e-6 * abs(rec["checksum"])
        assert torch.allclose(MO.calculate_psnr(sr, hr), rec["psnr"], rtol=1e-5)
        assert torch.allclose(MO.ssim(sr, hr), rec["ssim"], rtol=1e-5)
        assert torch.allclose(MO.ssim(sr, hr, 11, False), rec["ssim_per_image"], rtol=1e-5)
    assert MO.calculate_psnr(hr, hr) == float("inf")
