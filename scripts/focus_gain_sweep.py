"""How the end-to-end focus-loss gradient error (engine / stock autocast vs the fp32 oracle) depends on the conditioning of the
frozen recogniser: residual-branch gain of the synthetic weights (BatchNorm statistics re-calibrated for each), query/key scale.
Usage: python scripts/focus_gain_sweep.py"""
import os
import re
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import focus_oracle as FO, synth

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dic = FO.synth_decomposition()


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def build(gain, qk):
    sd = synth.synth_state_dict(synth.load_spec("focus"), seed=777, computed={"pe.pe": FO.positional_encoding(512, 5000)})
    for k in sd:
        if re.search(r"layer\d\.\d+\.bn2\.weight$", k):
            sd[k] = sd[k] * gain
        if re.search(r"decoder\.multihead\.linears\.[01]\.(weight|bias)$", k):
            sd[k] = sd[k] * qk
    sd = {k: v.to(DEV) for k, v in sd.items()}
    # calibrate the running statistics on a batch (what make_golden_focus.py does with the real module)
    orig = FO._conv_bn_pre

    def calib(sd_, conv, bn, x, nm, fp32_weights=False):
        y = F.conv2d(x, sd_[conv + ".weight"], sd_[conv + ".bias"], stride=1, padding=1)
        sd_[bn + ".running_mean"] = y.mean((0, 2, 3)).detach()
        sd_[bn + ".running_var"] = y.var((0, 2, 3), unbiased=True).detach()
        return F.batch_norm(y, None, None, sd_[bn + ".weight"], sd_[bn + ".bias"], training=True, eps=1e-5)
    FO._conv_bn_pre = calib
    try:
        with torch.no_grad():
            _, cal = synth.synth_images(64, seed=11)
            FO.resnet_encoder(sd, FO.to_gray_tensor(cal.to(DEV)), nm=FO.Numerics())
    finally:
        FO._conv_bn_pre = orig
    return sd


def measure(tag, sd, sr_kind="bilinear"):
    from fudanocr_b200.loss.stroke_focus_loss import StrokeFocusLoss
    crit = StrokeFocusLoss(types.SimpleNamespace(text_focus=True, stroke_lambda=50), decomposition=dic,
                           transformer_state_dict={k: v.detach().cpu() for k, v in sd.items()}).to(DEV)
    lr, hr = synth.synth_images(6, seed=23)
    sr = F.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False).clamp(0, 1).to(DEV)
    if sr_kind == "other":       # an unrelated crop: the two attention maps are far apart (early training)
        sr = synth.synth_images(6, seed=71)[1].to(DEV)
    elif sr_kind == "flat":
        sr = torch.full_like(sr, 0.5)
    hr = hr.to(DEV)
    labels = ["a", "focus", "B200", "stroke9", "xyzzy", "Q"]

    def oracle(autocast):
        s = sr.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss, mse, att, info = FO.stroke_focus_loss(sd, s, hr, labels, dic, 50.0, nm=FO.Numerics(fold=True))
        (loss.float() * 100).backward()
        return att.detach().float(), info["map_sr"].detach().float(), s.grad
    att_o, ms_o, g_o = oracle(False)
    att_b, ms_b, g_b = oracle(True)
    _, text_input, _ = crit.label_stroke_encoder(labels, DEV)
    losses, d_sr, mh, ms = crit._run(sr, hr, text_input, 50.0, 100.0, maps=True)
    torch.cuda.synchronize()
    mse_g = 200.0 * (sr - hr) / sr.numel()
    with torch.no_grad():
        mh_o = FO.attention_map(sd, FO.to_gray_tensor(hr), text_input, FO.Numerics(fold=True))
        flips = float(((ms - mh).sign() != (ms_o - mh_o).sign()).float().mean())
        near = float(((ms_o - mh_o).abs() < 2e-2 * ms_o.abs()).float().mean())
    # the fp32 oracle differentiated with the ENGINE's L1 sign pattern: everything else (activations, ReLU masks, pooling
    # arg-max, softmax) is the oracle's own fp32 evaluation
    s2 = sr.clone().requires_grad_(True)
    seed = (ms - mh).sign()
    ms2 = FO.attention_map(sd, FO.to_gray_tensor(s2), text_input, FO.Numerics(fold=True))
    (100.0 * (F.mse_loss(s2, hr) + 50.0 * (seed * (ms2 - mh_o)).mean())).backward()
    print(f"d_sr vs fp32 oracle with the engine's L1 signs: {rel(d_sr, s2.grad):.3e}; ", end="")
    print(f"[{sr_kind}] sign flips {flips:.3f}, |map_sr - map_hr| < 2% of map: {near:.3f}; ", end="")
    print(f"{tag}: d_sr engine {rel(d_sr, g_o):.3e} stock {rel(g_b, g_o):.3e} | attention part: engine {rel(d_sr - mse_g, g_o - mse_g):.3e} "
          f"stock {rel(g_b - mse_g, g_o - mse_g):.3e} | |att grad|/|mse grad| {float((g_o - mse_g).norm() / mse_g.norm()):.2f} | map_sr engine "
          f"{rel(ms, ms_o):.3e} stock {rel(ms_b, ms_o):.3e} | att {float(losses[2]):.5f} fp32 {float(att_o):.5f} | map max {float(ms_o.max()):.3f}", flush=True)


sd0 = build(0.25, 1.0)
for kind in ("bilinear", "other", "flat"):
    measure("gain 0.25", sd0, kind)
