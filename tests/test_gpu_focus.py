"""Stroke-focus loss (text-gestalt StrokeFocusLoss + its frozen recogniser) on the focr engine vs the oracle and the golden
fixture recorded from the unmodified reference classes.  GPU only.

Two yardsticks (the engine computes in bf16 with fp32 accumulation; north_star: 1e-2 relative for bf16):
 * SHARP (teacher-forced): the randomly-initialised 30-layer recogniser is chaotic - a 2^-9 relative perturbation moves its
   output by ~10 % and the L1 term's sign() and the ReLU masks turn that into O(1) gradient differences - so kernels are
   checked where chaos cannot enter: every forward layer against the oracle layer applied to the engine's OWN input
   activation (<= 1e-2, measured 2-3e-3), and the whole input-gradient chain against autograd of the oracle linearised at
   the engine's own activations (<= 2e-2).  Plus the oracle with the engine's bf16 rounding points (maps <= 2e-2).
 * CALIBRATED: the plain fp32 oracle / the golden fixture recorded from the reference.  A bf16 evaluation of a chaotic
   network cannot track fp32 to 1e-2 end to end; as in test_gpu_tbsrn.py the engine must be at least as close to fp32 as
   stock PyTorch autocast(bf16) of the same restatement (x1.25), and under a fixed cap."""
import json
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
REPORT = {}


def _dump():
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(REPORT, open("gpurun_out/focus_parity.json", "w"), indent=1)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def env():
    from oracle import synth, focus_oracle as FO
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.loss.stroke_focus_loss import StrokeFocusLoss
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.load(synth.GOLDEN_DIR / "focus_b2.pt", weights_only=False)
    sd = FO.synth_recogniser_state_dict(synth.load_spec("focus"), g["bn_stats"])
    dic = FO.synth_decomposition()
    crit = StrokeFocusLoss(types.SimpleNamespace(text_focus=True, stroke_lambda=50), decomposition=dic,
                           transformer_state_dict=sd).to(DEV)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    return dict(FO=FO, L=L, g=g, sd=sd_dev, dic=dic, crit=crit, synth=synth)


def test_golden_fixture_calibrated(env):
    """engine vs the fixture recorded from the unmodified reference classes (fp32)"""
    g, crit = env["g"], env["crit"]
    sr, hr = g["sr"].to(DEV), g["hr"].to(DEV)
    losses, d_sr, mh, ms = crit._run(sr, hr, g["text_input"].to(DEV), 50.0, 100.0, maps=True)
    torch.cuda.synchronize()
    r_hr, r_sr = _rel(mh.cpu(), g["map_hr"]), _rel(ms.cpu(), g["map_sr"])
    loss, mse, att = [float(x) for x in losses.cpu()]
    r_g = _rel(d_sr.cpu(), g["d_sr_total_x100"])
    REPORT["golden"] = dict(map_hr_rel=r_hr, map_sr_rel=r_sr, loss=loss, loss_ref=float(g["loss"]), mse=mse,
                            attention=att, attention_ref=float(g["attention_loss"]), d_sr_rel=r_g)
    _dump()
    assert r_hr < 0.1 and r_sr < 0.1, (r_hr, r_sr)
    assert abs(mse - float(g["mse"])) < 1e-5 * float(g["mse"]) + 1e-9
    assert abs(att - float(g["attention_loss"])) < 2e-2 * float(g["attention_loss"]), (att, float(g["attention_loss"]))
    assert abs(loss - float(g["loss"])) < 2e-2 * float(g["loss"])


def _oracle(env, sr, hr, labels, mode="fp32"):
    FO = env["FO"]
    sr = sr.clone().requires_grad_(True)
    nm = FO.Numerics.bf16_emulation() if mode == "emu" else FO.Numerics(fold=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast")):
        loss, mse, att, info = FO.stroke_focus_loss(env["sd"], sr, hr, labels, env["dic"], 50.0, nm=nm)
    (loss.float() * 100).backward()
    return loss.detach().float(), att.detach().float(), info["map_hr"].detach().float(), info["map_sr"].detach().float(), sr.grad


def _case(env, B):
    synth = env["synth"]
    if B == 2:
        return env["g"]["sr"].to(DEV), env["g"]["hr"].to(DEV), env["g"]["labels"]
    lr, hr = synth.synth_images(B, seed=23)
    sr = torch.nn.functional.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False).clamp(0, 1).to(DEV)
    return sr, hr.to(DEV), ["a", "focus", "B200", "stroke9", "xyzzy", "Q"][:B]


def _engine_acts(env, B, T):
    """the SR-branch activations the engine left in its workspace, as NCHW / token fp32 tensors keyed like the oracle"""
    import ctypes as C
    crit, L, FO = env["crit"], env["L"], env["FO"]

    def ws(name, shape, nhwc=True):
        off, n, eb = C.c_longlong(), C.c_longlong(), C.c_int()
        L.check(L.lib.focr_focus_loss_ws_tensor(B, T, name.encode(), C.byref(off), C.byref(n), C.byref(eb)))
        raw = crit._ws[off.value: off.value + n.value * eb.value].view(torch.bfloat16 if eb.value == 2 else torch.float32)
        t = raw[: int(torch.tensor(shape).prod())].view(*shape).float()
        return t.permute(0, 3, 1, 2).contiguous() if nhwc else t.clone()

    pre = "encoder.cnn"
    acts = {f"{pre}.conv1": ws("a1", (B, 32, 128, 64)), f"{pre}.conv2": ws("a2", (B, 16, 64, 128))}
    ci = 2
    for li, (nblk, (cin, cout)) in enumerate(zip(FO.LAYERS, FO.PLANES), start=1):
        for bi in range(nblk):
            blk = f"{pre}.layer{li}.{bi}"
            acts[blk + ".conv1"] = ws(f"act{ci}", (B, 8, 32, cout))
            acts[blk + ".conv2"] = ws(f"act{ci + 1}", (B, 8, 32, cout))
            ci += 3 if (bi == 0 and cin != cout) else 2
        name = f"{pre}.layer{li}_conv" if li < 4 else f"{pre}.layer4_conv2"
        acts[name] = ws(f"act{ci}", (B, 8, 32, cout if li < 4 else 1024))
        ci += 1
    acts["decoder.multihead.linears.1"] = ws("K", (B, 256, 1024), nhwc=False)
    acts["decoder.multihead.linears.0"] = ws("Q", (B, T, 1024), nhwc=False)
    acts["decoder.multihead.map"] = ws("map_sr", (B, 16, T, 256), nhwc=False)
    return acts, ws("map_hr", (B, 16, T, 256), nhwc=False)


@pytest.mark.parametrize("B", [2, 6])
def test_sharp_input_gradient_teacher_forced(env, B):
    """SHARP backward check: autograd of the oracle linearised at the engine's own activations (same ReLU masks, same
    pooling argmax, same sign pattern) vs the engine's input-gradient chain"""
    crit, FO = env["crit"], env["FO"]
    sr, hr, labels = _case(env, B)
    _, text_input, _ = crit.label_stroke_encoder(labels, DEV)
    losses, d_sr = crit._run(sr, hr, text_input, 50.0, 100.0)
    torch.cuda.synchronize()
    T = text_input.shape[1]
    acts, map_hr = _engine_acts(env, B, T)
    nm = FO.Numerics.teacher_forced(acts)
    x = sr.clone().requires_grad_(True)
    map_sr = FO.attention_map(env["sd"], FO.to_gray_tensor(x), text_input, nm)
    assert torch.allclose(map_sr.detach(), acts["decoder.multihead.map"], rtol=1e-5, atol=1e-9)
    att = torch.nn.functional.l1_loss(map_hr, map_sr)
    (att * 50.0 * 100.0).backward()
    d_attn = d_sr - 200.0 * (sr - hr) / sr.numel()
    rep = dict(d_sr_attn_rel=_rel(d_attn, x.grad), norm=float(x.grad.norm()), att=float(att.detach()), att_engine=float(losses[2]))
    REPORT[f"teacher_bwd_B{B}"] = rep
    _dump()
    assert abs(rep["att"] - rep["att_engine"]) < 1e-4 * rep["att"], rep
    assert rep["d_sr_attn_rel"] < 2e-2, rep


@pytest.mark.parametrize("B", [2, 6])
def test_vs_bf16_emulated_oracle(env, B):
    """maps / loss / gradient against the oracle evaluated with the engine's rounding points.  Even this twin diverges
    from the engine through bf16 rounding-boundary flips amplified by the chaotic random network, so the maps are held to
    2e-2 and the gradient is only reported (its sharp check is the teacher-forced test above)"""
    crit = env["crit"]
    sr, hr, labels = _case(env, B)
    loss_e, att_e, mh_e, ms_e, g_e = _oracle(env, sr, hr, labels, "emu")
    _, text_input, _ = crit.label_stroke_encoder(labels, DEV)
    losses, d_sr, mh, ms = crit._run(sr, hr, text_input, 50.0, 100.0, maps=True)
    torch.cuda.synchronize()
    rep = dict(map_hr=_rel(mh, mh_e), map_sr=_rel(ms, ms_e), att=float(losses[2]), att_emu=float(att_e),
               loss=float(losses[0]), loss_emu=float(loss_e), d_sr=_rel(d_sr, g_e))
    REPORT[f"emulated_B{B}"] = rep
    _dump()
    assert rep["map_hr"] < 2e-2 and rep["map_sr"] < 2e-2, rep
    assert abs(rep["att"] - rep["att_emu"]) < 1e-2 * rep["att_emu"], rep
    assert abs(rep["loss"] - rep["loss_emu"]) < 1e-2 * rep["loss_emu"], rep


@pytest.mark.parametrize("B", [2, 6])
def test_calibrated_vs_fp32_oracle(env, B):
    """against the fp32 oracle on the GPU; stock autocast(bf16) of the oracle is the yardstick"""
    crit = env["crit"]
    sr, hr, labels = _case(env, B)
    loss_o, att_o, mh_o, ms_o, g_o = _oracle(env, sr, hr, labels, "fp32")
    loss_b, att_b, mh_b, ms_b, g_b = _oracle(env, sr, hr, labels, "autocast")
    _, text_input, _ = crit.label_stroke_encoder(labels, DEV)
    losses, d_sr, mh, ms = crit._run(sr, hr, text_input, 50.0, 100.0, maps=True)
    torch.cuda.synchronize()
    rep = dict(map_hr=_rel(mh, mh_o), map_sr=_rel(ms, ms_o), map_sr_stock_bf16=_rel(ms_b, ms_o),
               att=float(losses[2]), att_ref=float(att_o), att_stock_bf16=float(att_b),
               d_sr=_rel(d_sr, g_o), d_sr_stock_bf16=_rel(g_b, g_o))
    REPORT[f"oracle_B{B}"] = rep
    _dump()
    assert rep["map_sr"] < max(1e-2, 1.25 * rep["map_sr_stock_bf16"]), rep
    att_err = abs(rep["att"] - rep["att_ref"]) / rep["att_ref"]
    att_err_b = abs(rep["att_stock_bf16"] - rep["att_ref"]) / rep["att_ref"]
    assert att_err < max(1e-2, 1.25 * att_err_b), rep
    assert rep["d_sr"] < max(1e-2, 1.25 * rep["d_sr_stock_bf16"]), rep


def test_teacher_forced_layers(env):
    """every encoder convolution checked in isolation: the oracle layer applied to the ENGINE's own input activation"""
    import ctypes as C
    import torch.nn.functional as F
    g, crit, L, FO, sd = env["g"], env["crit"], env["L"], env["FO"], env["sd"]
    sr, hr = g["sr"].to(DEV), g["hr"].to(DEV)
    text_input = g["text_input"].to(DEV)
    crit._run(sr, hr, text_input, 50.0, 100.0)
    torch.cuda.synchronize()
    B, T = text_input.shape

    def ws(name, shape):
        off, n, eb = C.c_longlong(), C.c_longlong(), C.c_int()
        L.check(L.lib.focr_focus_loss_ws_tensor(B, T, name.encode(), C.byref(off), C.byref(n), C.byref(eb)))
        raw = crit._ws[off.value: off.value + n.value * eb.value].view(torch.bfloat16)
        return raw.view(*shape).permute(0, 3, 1, 2).float()

    nm = FO.Numerics(fold=True)
    pre = "encoder.cnn"
    worst = {}
    with torch.no_grad():
        a1 = ws("a1", (B, 32, 128, 64))
        worst["conv1"] = _rel(a1, FO._conv_bn(sd, f"{pre}.conv1", f"{pre}.bn1", FO.to_gray_tensor(sr), True, nm))
        p2 = ws("p2", (B, 8, 32, 128))
        a2_ref = FO._conv_bn(sd, f"{pre}.conv2", f"{pre}.bn2", F.max_pool2d(a1, 2, 2), True, nm)
        worst["conv2+pool"] = _rel(p2, F.max_pool2d(a2_ref, 2, 2))
        x, ci = p2, 2
        for li, (nblk, (cin, cout)) in enumerate(zip(FO.LAYERS, FO.PLANES), start=1):
            for bi in range(nblk):
                blk = f"{pre}.layer{li}.{bi}"
                has_down = bi == 0 and cin != cout
                h1 = ws(f"act{ci}", (B, 8, 32, cout))
                worst[f"{blk}.conv1"] = _rel(h1, FO._conv_bn(sd, blk + ".conv1", blk + ".bn1", x, True, nm))
                out = ws(f"act{ci + 1}", (B, 8, 32, cout))
                res = FO._conv_bn(sd, blk + ".downsample.0", blk + ".downsample.1", x, False, nm) if has_down else x
                ref = F.relu(FO._conv_bn_pre(sd, blk + ".conv2", blk + ".bn2", h1, nm) + res)
                worst[f"{blk}.conv2"] = _rel(out, ref)
                x, ci = out, ci + (3 if has_down else 2)
            name = f"{pre}.layer{li}_conv" if li < 4 else f"{pre}.layer4_conv2"
            bn = f"{pre}.layer{li}_bn" if li < 4 else f"{pre}.layer4_conv2_bn"
            y = ws(f"act{ci}", (B, 8, 32, cout if li < 4 else 1024))
            worst[name] = _rel(y, FO._conv_bn(sd, name, bn, x, True, nm))
            x, ci = y, ci + 1
    REPORT["teacher_forced"] = worst
    _dump()
    bad = {k: v for k, v in worst.items() if not v < 1e-2}
    assert not bad, bad


def test_autograd_surface_matches_fused_call(env):
    """forward() returns the reference 4-tuple; loss.backward() delivers the same gradient as the fused entry"""
    g, crit = env["g"], env["crit"]
    sr = g["sr"].to(DEV).requires_grad_(True)
    hr = g["hr"].to(DEV)
    loss, mse, att, rec = crit(sr, hr, g["labels"])
    assert rec == -1 and loss.requires_grad and not mse.requires_grad
    (loss * 100).backward()
    d = torch.empty_like(hr)
    crit.loss_and_grad(sr.detach(), hr, g["labels"], 100.0, d)
    # gscale = 100 inside the chain vs x100 afterwards: the bf16 gradient buffers round differently (100 is not 2^k)
    assert _rel(sr.grad, d) < 2e-2
    d2 = torch.empty_like(hr)
    crit.loss_and_grad(sr.detach(), hr, g["labels"], 100.0, d2)
    assert torch.equal(d, d2)   # the engine is deterministic
    crit.args.text_focus = False
    try:
        sr2 = g["sr"].to(DEV).requires_grad_(True)
        l2, m2, a2, r2 = crit(sr2, hr, g["labels"])
        assert a2 == -1 and r2 == -1
        l2.backward()
        ref = 2 * (sr2.detach() - hr) / hr.numel()
        assert _rel(sr2.grad, ref) < 1e-5
    finally:
        crit.args.text_focus = True


def test_encoder_features_vs_oracle(env):
    """the recogniser's encoder output (B,1024,8,32) after the SR-branch forward, read back from the workspace"""
    import ctypes as C
    g, crit, L, FO = env["g"], env["crit"], env["L"], env["FO"]
    sr, hr = g["sr"].to(DEV), g["hr"].to(DEV)
    text_input = g["text_input"].to(DEV)
    crit._run(sr, hr, text_input, 50.0, 100.0)
    torch.cuda.synchronize()
    B, T = text_input.shape
    off, n, eb = C.c_longlong(), C.c_longlong(), C.c_int()
    L.check(L.lib.focr_focus_loss_ws_tensor(B, T, b"feat", C.byref(off), C.byref(n), C.byref(eb)))
    feat = crit._ws[off.value: off.value + n.value * eb.value].view(torch.bfloat16).view(B, 8, 32, 1024).permute(0, 3, 1, 2).float()
    with torch.no_grad():
        ref = FO.resnet_encoder(env["sd"], FO.to_gray_tensor(sr))
    r = _rel(feat, ref)
    L.check(L.lib.focr_focus_loss_ws_tensor(B, T, b"Q", C.byref(off), C.byref(n), C.byref(eb)))
    Q = crit._ws[off.value: off.value + n.value * eb.value].view(torch.bfloat16).view(-1, 1024)[: B * T].float()
    with torch.no_grad():
        q_ref = FO.decoder_query(env["sd"], FO.text_embedding(env["sd"], text_input))
        q_ref = torch.nn.functional.linear(q_ref, env["sd"]["decoder.multihead.linears.0.weight"],
                                           env["sd"]["decoder.multihead.linears.0.bias"]).reshape(B * T, 1024)
    rq = _rel(Q, q_ref)
    REPORT["features"] = dict(feat_rel=r, q_rel=rq)
    _dump()
    assert r < 0.15 and rq < 1e-2, (r, rq)   # feat: chaotic end-to-end (see module docstring); sharp checks above
