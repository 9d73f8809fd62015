"""Text-focus loss (scene-text-telescope TextFocusLoss: MSE + 10 L1(attention maps) + 5e-4 weighted CE) on the focr engine vs
the oracle and the golden fixture recorded from the unmodified reference modules.  GPU only.  Same two yardsticks as
test_gpu_focus.py: SHARP teacher-forced checks (decoder tail layer by layer, the whole input-gradient chain) and CALIBRATED
end-to-end checks against fp32 with stock autocast(bf16) as the yardstick."""
import ctypes as C
import json
import os
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
REPORT = {}


def _dump():
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(REPORT, open("gpurun_out/textfocus_parity.json", "w"), indent=1)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def env():
    from oracle import synth, focus_oracle as FO, textfocus_oracle as TF
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.loss.text_focus_loss import TextFocusLoss
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.load(synth.GOLDEN_DIR / "textfocus_b3.pt", weights_only=False)
    sd = FO.synth_recogniser_state_dict(synth.load_spec("textfocus"), g["bn_stats"], seed=778)
    crit = TextFocusLoss(types.SimpleNamespace(text_focus=True), confuse_counts=g["confuse_counts"].numpy(),
                         transformer_state_dict=sd).to(DEV)
    return dict(FO=FO, TF=TF, L=L, g=g, sd={k: v.to(DEV) for k, v in sd.items()}, crit=crit, synth=synth,
                table=g["weight_table"].to(DEV))


def _enc(env, labels):
    crit = env["crit"]
    from fudanocr_b200.loss.text_focus_loss import str_filt
    return crit.label_encoder([str_filt(s, "lower") + "-" for s in labels], DEV)


def test_golden_fixture_calibrated(env):
    g, crit = env["g"], env["crit"]
    sr, hr = g["sr"].to(DEV), g["hr"].to(DEV)
    losses, d_sr, mh, ms, pred = crit._run_text(sr, hr, _enc(env, g["labels"]), 100.0, outputs=True)
    torch.cuda.synchronize()
    loss, mse, att, rec = [float(x) for x in losses.cpu()]
    rep = dict(map_sr_rel=_rel(ms.cpu(), g["map_sr"]), sr_pred_rel=_rel(pred.cpu(), g["sr_pred"]), loss=loss,
               loss_ref=float(g["loss"]), attention=att, attention_ref=float(g["attention_loss"]), recognition=rec,
               recognition_ref=float(g["recognition_loss"]), d_sr_rel=_rel(d_sr.cpu(), g["d_sr_total_x100"]))
    REPORT["golden"] = rep
    _dump()
    assert rep["map_sr_rel"] < 0.1 and rep["sr_pred_rel"] < 0.1, rep
    assert abs(mse - float(g["mse"])) < 1e-5 * float(g["mse"]) + 1e-9
    assert abs(att - rep["attention_ref"]) < 2e-2 * rep["attention_ref"], rep
    assert abs(rec - rep["recognition_ref"]) < 2e-2 * rep["recognition_ref"], rep
    assert abs(loss - rep["loss_ref"]) < 2e-2 * rep["loss_ref"], rep


def _ws(env, B, T, name, shape, nhwc=False):
    crit, L = env["crit"], env["L"]
    off, n, eb = C.c_longlong(), C.c_longlong(), C.c_int()
    L.check(L.lib.focr_focus_loss_ws_tensor(B, T, name.encode(), C.byref(off), C.byref(n), C.byref(eb)))
    raw = crit._ws[off.value: off.value + n.value * eb.value].view(torch.bfloat16 if eb.value == 2 else torch.float32)
    numel = 1
    for d in shape:
        numel *= d
    t = raw[:numel].view(*shape).float()
    return t.permute(0, 3, 1, 2).contiguous() if nhwc else t.clone()


def _engine_acts(env, B, T):
    FO = env["FO"]
    pre = "encoder.cnn"
    acts = {f"{pre}.conv1": _ws(env, B, T, "a1", (B, 32, 128, 64), True), f"{pre}.conv2": _ws(env, B, T, "a2", (B, 16, 64, 128), True)}
    ci = 2
    for li, (nblk, (cin, cout)) in enumerate(zip(FO.LAYERS, FO.PLANES), start=1):
        for bi in range(nblk):
            blk = f"{pre}.layer{li}.{bi}"
            acts[blk + ".conv1"] = _ws(env, B, T, f"act{ci}", (B, 8, 32, cout), True)
            acts[blk + ".conv2"] = _ws(env, B, T, f"act{ci + 1}", (B, 8, 32, cout), True)
            ci += 3 if (bi == 0 and cin != cout) else 2
        name = f"{pre}.layer{li}_conv" if li < 4 else f"{pre}.layer4_conv2"
        acts[name] = _ws(env, B, T, f"act{ci}", (B, 8, 32, cout if li < 4 else 1024), True)
        ci += 1
    d = "decoder"
    acts[f"{d}.multihead.linears.0"] = _ws(env, B, T, "Q", (B, T, 1024))
    acts[f"{d}.multihead.linears.1"] = _ws(env, B, T, "K", (B, 256, 1024))
    acts[f"{d}.multihead.linears.2"] = _ws(env, B, T, "V", (B, 256, 1024))
    acts[f"{d}.multihead.map"] = _ws(env, B, T, "map_sr", (B, 16, T, 256))
    acts[f"{d}.multihead.ctx"] = _ws(env, B, T, "ctx", (B, T, 1024))
    acts[f"{d}.x2"] = _ws(env, B, T, "x2", (B, T, 1024))
    acts[f"{d}.r2"] = _ws(env, B, T, "r2", (B, T, 1024))
    acts[f"{d}.pff.w_1"] = _ws(env, B, T, "hff", (B, T, 2048))
    acts[f"{d}.x3"] = _ws(env, B, T, "x3", (B, T, 1024))
    acts[f"{d}.r3"] = _ws(env, B, T, "r3", (B, T, 1024))
    return acts, _ws(env, B, T, "map_hr", (B, 16, T, 256))


def _case(env, B):
    if B == 3:
        g = env["g"]
        return g["sr"].to(DEV), g["hr"].to(DEV), g["labels"]
    lr, hr = env["synth"].synth_images(B, seed=31)
    sr = F.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False).clamp(0, 1).to(DEV)
    return sr, hr.to(DEV), ["Hello", "w0rld", "x", "textzoom", "B-200", "super", "res"][:B]


def test_teacher_forced_decoder_tail(env):
    """every stage of the SR-branch decoder tail against the oracle stage applied to the engine's own input"""
    FO, TF, sd0 = env["FO"], env["TF"], env["sd"]
    sd = TF._rename(sd0)
    sr, hr, labels = _case(env, 3)
    enc = _enc(env, labels)
    losses, d_sr, mh, ms, pred = env["crit"]._run_text(sr, hr, enc, 100.0, outputs=True)
    torch.cuda.synchronize()
    B, T = enc[1].shape
    a, _ = _engine_acts(env, B, T)
    d, nm = "decoder", FO.Numerics(fold=True)
    lin = lambda name, x: FO._lin(sd, name, x, nm)
    rep = {}
    with torch.no_grad():
        feat = a["encoder.cnn.layer4_conv2"]
        tokens = feat.view(B, 1024, 256).permute(0, 2, 1)
        rep["K"] = _rel(a[f"{d}.multihead.linears.1"], lin(f"{d}.multihead.linears.1", tokens))
        rep["V"] = _rel(a[f"{d}.multihead.linears.2"], lin(f"{d}.multihead.linears.2", tokens))
        q = a[f"{d}.multihead.linears.0"].view(B, T, 16, 64).transpose(1, 2)
        k = a[f"{d}.multihead.linears.1"].view(B, 256, 16, 64).transpose(1, 2)
        v = a[f"{d}.multihead.linears.2"].view(B, 256, 16, 64).transpose(1, 2)
        p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
        rep["map"] = _rel(a[f"{d}.multihead.map"], p)
        ctx = (a[f"{d}.multihead.map"] @ v).transpose(1, 2).reshape(B, T, 1024)
        rep["ctx"] = _rel(a[f"{d}.multihead.ctx"], ctx)
        query = _ws(env, B, T, "query", (B, T, 1024))
        rep["x2"] = _rel(a[f"{d}.x2"], query + lin(f"{d}.multihead.linears.3", a[f"{d}.multihead.ctx"]))
        rep["r2"] = _rel(a[f"{d}.r2"], FO.layer_norm_std(a[f"{d}.x2"], sd[f"{d}.mul_layernorm2.a_2"], sd[f"{d}.mul_layernorm2.b_2"]))
        rep["hff"] = _rel(a[f"{d}.pff.w_1"], F.relu(lin(f"{d}.pff.w_1", a[f"{d}.r2"])))
        rep["x3"] = _rel(a[f"{d}.x3"], a[f"{d}.r2"] + lin(f"{d}.pff.w_2", a[f"{d}.pff.w_1"]))
        rep["r3"] = _rel(a[f"{d}.r3"], FO.layer_norm_std(a[f"{d}.x3"], sd[f"{d}.mul_layernorm3.a_2"], sd[f"{d}.mul_layernorm3.b_2"]))
        logits = lin("generator_word_with_upperword.proj", a[f"{d}.r3"])
        packed = torch.cat([logits[i, :int(n)] for i, n in enumerate(enc[0].tolist())], 0)
        rep["logits"] = _rel(pred, packed)
        rec = TF.weight_cross_entropy(pred, enc[2], env["table"])
        rep["wce_value_rel"] = abs(float(losses[3]) - float(rec)) / float(rec)
    REPORT["teacher_forced_tail"] = rep
    _dump()
    bad = {k: v for k, v in rep.items() if not v < 1e-2}
    assert not bad, rep


@pytest.mark.parametrize("B", [3, 7])
def test_sharp_input_gradient_teacher_forced(env, B):
    """the whole input-gradient chain (weighted CE -> generator -> LN3 -> FFN -> LN2 -> out/value/key projections -> softmax ->
    encoder -> gray) vs autograd of the oracle linearised at the engine's own activations"""
    FO, TF, crit = env["FO"], env["TF"], env["crit"]
    sr, hr, labels = _case(env, B)
    enc = _enc(env, labels)
    losses, d_sr = crit._run_text(sr, hr, enc, 100.0)
    torch.cuda.synchronize()
    T = enc[1].shape[1]
    acts, map_hr = _engine_acts(env, B, T)
    nm = FO.Numerics.teacher_forced(acts)
    out = {}
    for part, (la, lc) in {"total": (10.0, 0.0005), "ce_only": (0.0, 0.0005)}.items():
        x = sr.clone().requires_grad_(True)
        loss, mse, att, rec, info = TF.text_focus_loss(env["sd"], x, hr, labels, env["table"], la, lc, nm=nm, map_hr=map_hr)
        ((loss - mse) * 100.0).backward()
        out[part] = x.grad
    d_attn_ce = d_sr - 200.0 * (sr - hr) / sr.numel()
    rep = dict(d_sr_rel=_rel(d_attn_ce, out["total"]), ce_share=float(out["ce_only"].norm() / out["total"].norm()))
    # isolate the recognition-term chain: run the engine with lambda_attn = 0
    crit.lambda_attn = 0.0
    try:
        _, d_ce = crit._run_text(sr, hr, enc, 100.0)
    finally:
        crit.lambda_attn = 10.0
    torch.cuda.synchronize()
    rep["d_sr_ce_only_rel"] = _rel(d_ce - 200.0 * (sr - hr) / sr.numel(), out["ce_only"])
    REPORT[f"teacher_bwd_B{B}"] = rep
    _dump()
    assert rep["d_sr_rel"] < 2e-2 and rep["d_sr_ce_only_rel"] < 3e-2, rep


@pytest.mark.parametrize("B", [3])
def test_calibrated_vs_fp32_oracle(env, B):
    TF, FO, crit = env["TF"], env["FO"], env["crit"]
    sr, hr, labels = _case(env, B)

    def oracle(autocast):
        x = sr.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss, mse, att, rec, info = TF.text_focus_loss(env["sd"], x, hr, labels, env["table"], nm=FO.Numerics(fold=True))
        (loss.float() * 100).backward()
        return float(att), float(rec), info["sr_pred"].detach().float(), x.grad

    att_o, rec_o, pred_o, g_o = oracle(False)
    att_b, rec_b, pred_b, g_b = oracle(True)
    losses, d_sr, mh, ms, pred = crit._run_text(sr, hr, _enc(env, labels), 100.0, outputs=True)
    torch.cuda.synchronize()
    rep = dict(sr_pred=_rel(pred, pred_o), sr_pred_stock_bf16=_rel(pred_b, pred_o), rec=float(losses[3]), rec_ref=rec_o,
               rec_stock_bf16=rec_b, d_sr=_rel(d_sr, g_o), d_sr_stock_bf16=_rel(g_b, g_o))
    REPORT[f"oracle_B{B}"] = rep
    _dump()
    assert rep["sr_pred"] < max(1e-2, 1.25 * rep["sr_pred_stock_bf16"]), rep
    assert abs(rep["rec"] - rec_o) / rec_o < max(1e-2, 1.25 * abs(rec_b - rec_o) / rec_o), rep
    assert rep["d_sr"] < max(1e-2, 1.25 * rep["d_sr_stock_bf16"]), rep


def test_autograd_surface(env):
    g, crit = env["g"], env["crit"]
    sr = g["sr"].to(DEV).requires_grad_(True)
    hr = g["hr"].to(DEV)
    loss, mse, att, rec = crit(sr, hr, g["labels"])
    assert loss.requires_grad and not rec.requires_grad and not att.requires_grad
    (loss * 100).backward()
    d = torch.empty_like(hr)
    crit.loss_and_grad(sr.detach(), hr, g["labels"], 100.0, d)
    assert _rel(sr.grad, d) < 2e-2
