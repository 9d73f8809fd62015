"""CPU: the CRNN and TSRN oracle restatements vs golden vectors recorded from the real reference modules."""
import os
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


def test_resize_oracle_vs_pillow_golden():
    """numpy restatement of Pillow's 8-bit bicubic resampler (resizeNormalize, dataset.py:136-152) vs outputs recorded
    from PIL itself (oracle/make_golden_resize.py) - bit-exact uint8; plus PIL live where importable"""
    import numpy as np
    from oracle import resize_oracle as R
    g = np.load(synth.GOLDEN_DIR / "resize.npz")
    crops = R.synth_crops(24, seed=7)
    assert len(g.files) == 2 * len(crops)
    for size in ((128, 32), (64, 16)):
        for i, c in enumerate(crops):
            assert np.array_equal(R.resize_bicubic_u8(c, size), g[f"{size[0]}x{size[1]}_{i}"]), (i, c.shape, size)
    t = R.resize_normalize(crops[0], (128, 32))
    assert t.shape == (3, 32, 128) and t.dtype == np.float32 and 0.0 <= t.min() and t.max() <= 1.0
    # size-preserving resize is the identity (Pillow skips both passes)
    assert np.array_equal(R.resize_bicubic_u8(crops[-2], (128, 32)), crops[-2])
    # coefficient rows are normalised: each sums to 2^22 within the rounding of its taps
    for in_size, out_size in ((200, 64), (9, 16), (33, 32), (1, 16)):
        bounds, kk = R.precompute_coeffs(in_size, out_size)
        assert (np.abs(kk.sum(1) - (1 << R.PRECISION_BITS)) <= kk.shape[1]).all()
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= in_size).all()
    try:
        from PIL import Image
    except ImportError:
        return
    rs = np.random.RandomState(3)
    for h, w in ((1, 1), (2, 300), (7, 13), (64, 256), (100, 31)):
        c = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        for size in ((128, 32), (64, 16)):
            assert np.array_equal(R.resize_bicubic_u8(c, size), np.asarray(Image.fromarray(c, "RGB").resize(size, Image.BICUBIC)))


def test_ctc_oracle_vs_torch_golden():
    """float64 CTC forward-backward restatement vs values recorded from torch.nn.functional.ctc_loss + autograd
    (oracle/make_golden_ctc.py), plus closed-form cases"""
    import numpy as np
    from oracle import ctc_oracle as CO
    g = np.load(synth.GOLDEN_DIR / "ctc.npz")
    cases = {"crnn_b8": (26, 8, 37, 12, 1, True), "full_b4": (26, 4, 37, 10, 2, False), "wide_b3": (40, 3, 97, 19, 3, True),
             "tiny_b5": (3, 5, 5, 2, 4, True)}
    for name, (T, B, C, S, seed, ragged) in cases.items():
        logits, targets, il, tl = CO.synth_case(T, B, C, S, seed, ragged)
        for red in ("mean", "sum", "none"):
            loss, _, grad = CO.ctc_loss(logits, targets, il, tl, 0, red)
            assert np.allclose(loss, g[f"{name}/{red}/loss"], rtol=1e-10)
            assert np.allclose(grad, g[f"{name}/{red}/grad"], rtol=1e-8, atol=1e-12)
    # T = 1, one label: nll = -log softmax(label)
    x = np.array([[[0.3, -1.2, 2.0]]], np.float32)
    loss, nll, grad = CO.ctc_loss(x, np.array([[2]]), [1], [1], 0, "sum")
    lp = CO.log_softmax(x)[0, 0]
    assert abs(nll[0] + lp[2]) < 1e-12
    assert np.allclose(grad[0, 0], np.exp(lp) - np.array([0, 0, 1.0]))
    # empty target: every frame must be blank
    x = np.random.RandomState(0).randn(5, 1, 4).astype(np.float32)
    loss, nll, _ = CO.ctc_loss(x, np.zeros((1, 1), np.int64), [5], [0], 0, "sum")
    assert abs(nll[0] + CO.log_softmax(x)[:, 0, 0].sum()) < 1e-12
    # a doubled character needs a blank in between: "aa" in 2 frames is infeasible, in 3 frames has exactly one path
    x = np.random.RandomState(1).randn(3, 1, 3).astype(np.float32)
    _, nll2, _ = CO.ctc_loss(x[:2], np.array([[1, 1]]), [2], [2], 0, "sum")
    assert np.isinf(nll2[0])
    _, nll3, _ = CO.ctc_loss(x, np.array([[1, 1]]), [3], [2], 0, "sum")
    lp = CO.log_softmax(x)[:, 0]
    assert abs(nll3[0] + (lp[0, 1] + lp[1, 0] + lp[2, 1])) < 1e-12


def test_clip_oracle_matches_the_reference_loop():
    """oracle/clip_oracle.py against the reference's OWN code: CLIP.forward (image-ids-CTR/CCR-CLIP/model.py:209-222, the real class
    with a small text tower) produces the normalised features and logit_scale.exp(); lines 99-107 of main.py - read from the
    checkout and executed verbatim, minus .cuda() - produce the loss.  Skipped where the reference checkout is absent."""
    import importlib.util
    import sys
    import textwrap
    ref = "/root/reference/image-ids-CTR/CCR-CLIP"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    from oracle import clip_oracle as CO
    sys.path.insert(0, ref)
    try:
        spec = importlib.util.spec_from_file_location("ccr_clip_model_ref", os.path.join(ref, "model.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        torch.manual_seed(0)
        model = mod.CLIP(embed_dim=2048, context_length=30, vocab_size=40, transformer_width=64, transformer_heads=2,
                         transformer_layers=1).eval()
    finally:
        sys.path.remove(ref)
        sys.modules.pop("resnet50", None)
    B = 6
    image = torch.rand(B, 3, 32, 32)
    text = torch.randint(1, 39, (B, 30))
    text[:, -1] = 39                                          # the end token carries the arg-max (model.py:205)
    label = ["a", "b", "a", "c", "b", "d"]
    with torch.no_grad():
        image_features, text_features, logit_scale = model(image, text)
        raw_i, raw_t = model.encode_image(image), model.encode_text(text)
    src = open(os.path.join(ref, "main.py")).read().split("\n")
    lines = textwrap.dedent("\n".join(src[98:107])).replace(".cuda()", "")       # main.py:99-107
    assert lines.lstrip().startswith("logits_per_image") and "total_loss" in lines
    env = dict(torch=torch, image=image, label=label, image_features=image_features, text_features=text_features,
               logit_scale=logit_scale.reshape(1),           # nn.DataParallel gathers the 0-dim scale into a vector (main.py:98 indexes it)
               loss_img=torch.nn.CrossEntropyLoss(), loss_txt=torch.nn.CrossEntropyLoss())
    exec(lines, env)
    gt = CO.ground_truth(label)
    assert torch.equal(gt, env["ground_truth"]) and gt.tolist() == [0, 1, 0, 3, 1, 5]
    loss, logits = CO.contrastive_loss(raw_i, raw_t, model.logit_scale.detach(), gt)
    assert torch.allclose(logits, env["logits_per_image"], rtol=1e-5, atol=1e-6)
    assert abs(float(loss) - float(env["total_loss"])) < 1e-6 * abs(float(env["total_loss"])) + 1e-7
