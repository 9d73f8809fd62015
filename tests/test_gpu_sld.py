"""Stroke-level-decomposition recogniser on the engine (fudanocr_b200/model/transformer.py, trainer_sld.py)
vs the oracle restatement (pinned to the unmodified reference module by tests/golden/sld_b3.pt) on the GPU.

The 40-layer train-mode-BN encoder is ill-conditioned: the oracle itself moves by 1.2e-2 (encoder gradients, relative L2)
between fp32 and fp64 (measured, tests/test_sld_assembly.py).  bf16 parity is therefore judged the way the TBSRN tests do:
(a) the decoder and generator teacher-forced on the oracle's encoder features against fp32 within bf16 accuracy; (b) the whole
step, tensor by tensor, against the error STOCK PyTorch autocast(bf16) makes on the same restatement, measured in the same
test (report: gpurun_out/sld_parity.json, copy under profiles/); (c) the fused trainer against the unchanged reference loop
run on the drop-in module.  Kernel-level parity is in tests/test_gpu_recog_ops.py, assembly parity in test_sld_assembly.py."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def _setup():
    from oracle import sld_oracle as SO, synth
    from fudanocr_b200.model.transformer import Transformer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.load(synth.GOLDEN_DIR / "sld_b3.pt", weights_only=False)
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    model = Transformer("stroke")
    model.load_state_dict(sd, strict=False)
    model = model.to(DEV)
    image, strings = SO.synth_batch(g["B"])
    assert strings == g["strings"]
    return SO, g, sd, model, image.to(DEV), g["length"].to(DEV), g["text_input"].to(DEV), g["text_gt"].to(DEV)


def test_sld_decoder_on_reference_features_and_eval_forward():
    """teacher-forced: the engine's decoder fed the ORACLE's encoder features (bf16-rounded) must match the fp32 decoder"""
    SO, g, sd, model, image, length, text_input, text_gt = _setup()
    dsd = {k: v.to(DEV) for k, v in sd.items()}
    model.eval()
    with torch.no_grad():
        o_logits, o_map, o_conv = SO.forward(dsd, image, text_input, train=False)
        feat = o_conv.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        o_logits_tf, o_map_tf, _ = SO.forward(dsd, image, text_input, train=False, conv_feature=feat.float().permute(0, 3, 1, 2))
        out = model(None, length, text_input, conv_feature=feat.permute(0, 3, 1, 2), test=True)
        assert _rel(out["pred"], o_logits_tf) < 2e-2, _rel(out["pred"], o_logits_tf)
        assert (out["map"] - o_map_tf).abs().max().item() < 2e-2 * o_map_tf.max().item()
        # whole eval forward (running statistics): encoder through 40 bf16 convs
        full = model(image, length, text_input, test=True)
        e_conv, e_pred = _rel(full["conv"], o_conv), _rel(full["pred"], o_logits)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            a_logits, _, a_conv = SO.forward(dsd, image, text_input, train=False)
        s_conv, s_pred = _rel(a_conv, o_conv), _rel(a_logits, o_logits)
        assert e_conv < max(1.5 * s_conv, 2e-2) and e_pred < max(1.5 * s_pred, 2e-2), (e_conv, s_conv, e_pred, s_pred)
        assert torch.equal(full["pred"].argmax(2), o_logits.argmax(2)) or e_pred < s_pred * 1.5


def test_sld_train_step_vs_oracle_calibrated_against_stock_bf16():
    SO, g, sd, model, image, length, text_input, text_gt = _setup()
    from fudanocr_b200.trainer_sld import SLDTrainer
    model.train()
    model.dropout_p = 0.0
    # fp32 oracle and stock autocast(bf16) of the same restatement
    def run(autocast):
        osd = {k: v.to(DEV).clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss, logits, amap, conv = SO.loss_fn(osd, image, length, text_input, text_gt)
        loss.float().backward()
        return loss.detach().float(), {k: v.grad.float() for k, v in osd.items() if v.grad is not None}
    ref_loss, ref_g = run(False)
    amp_loss, amp_g = run(True)
    assert abs(float(ref_loss) - float(g["loss"])) < 1e-3         # the GPU fp32 oracle reproduces the reference's value
    # engine: fused loss + backward through the reference-loop API
    loss = model.loss(image, length, text_input, text_gt)
    loss.backward()
    eng_g = {k: p.grad.float() for k, p in model.named_parameters() if p.grad is not None}
    assert set(eng_g) == set(ref_g)
    report = {"loss": [float(loss), float(ref_loss), float(amp_loss)], "tensors": {}}
    bad = []
    for k, r in ref_g.items():
        if float(r.abs().max()) < 1e-6:      # conv biases ahead of a train-mode BatchNorm: true gradient 0
            continue
        e, s = _rel(eng_g[k], r), _rel(amp_g[k], r)
        report["tensors"][k] = [e, s]
        # every tensor - the decoder's too, whose inputs are the encoder features - inherits the encoder's bf16 sensitivity
        # (stock autocast is 7-30 % off in the decoder and ~90 % off in most encoder tensors at this batch size); measured
        # engine / stock ratio: 0.88 .. 1.19 over the 147 tensors (profiles/r01i_sld_parity.json)
        if not (e < max(1.5 * s, 5e-2)):
            bad.append((k, e, s))
    es = sorted(v[0] for v in report["tensors"].values())
    ss = sorted(v[1] for v in report["tensors"].values())
    report["median"] = [es[len(es) // 2], ss[len(ss) // 2]]
    report["worst"] = [es[-1], ss[-1]]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sld_parity.json", "w") as f:
        json.dump(report, f)
    assert abs(float(loss) - float(ref_loss)) < max(2 * abs(float(amp_loss) - float(ref_loss)), 2e-2 * float(ref_loss))
    assert not bad, bad[:8]
    assert report["median"][0] < 1.15 * report["median"][1] + 1e-2, report["median"]
    # running statistics follow the reference's update (momentum 0.1, unbiased variance)
    for k, v in g["running_after"].items():
        assert _rel(model.state_dict()[k].cpu(), v) < 3e-2, k


def test_sld_fused_trainer_step_and_reference_loop_agree():
    SO, g, sd, model, image, length, text_input, text_gt = _setup()
    from fudanocr_b200.model.transformer import Transformer
    from fudanocr_b200.trainer_sld import SLDTrainer
    model.train()
    model.dropout_p = 0.0
    twin = Transformer("stroke")
    twin.load_state_dict(sd, strict=False)
    twin = twin.to(DEV).train()
    twin.dropout_p = 0.0
    # reference loop, unchanged (train.py:63-77) on the drop-in module
    opt = torch.optim.Adadelta(twin.parameters(), lr=1.0, rho=0.9)
    opt.zero_grad()
    out = twin(image, length, text_input)
    assert out["pred"].shape == (int(length.sum()), 7)
    l_ref = torch.nn.CrossEntropyLoss()(out["pred"], text_gt)
    l_ref.backward()
    opt.step()
    # fused trainer
    tr = SLDTrainer(model)
    l_fused = tr.step(image, length, text_input, text_gt)
    torch.cuda.synchronize()
    assert abs(float(l_fused) - float(l_ref)) < 1e-5 * abs(float(l_ref)) + 1e-6
    worst = 0.0
    for (k, a), (_, b) in zip(model.named_parameters(), twin.named_parameters()):
        worst = max(worst, _rel(a.detach(), b.detach()))
    # same kernels feed both; they differ only where torch's CE / bf16 cast of its fp32 logits gradient enters
    assert worst < 2e-2, worst
    # a second step runs (gradient buffer re-zeroed, state carried) and lowers the loss on the same batch
    l2 = tr.step(image, length, text_input, text_gt)
    torch.cuda.synchronize()
    assert torch.isfinite(l2) and float(l2) < float(l_fused)
    # dropout on: finite, reproducible masks are covered by the op tests; here only that the path runs
    model.dropout_p = 0.1
    l3 = tr.step(image, length, text_input, text_gt)
    assert torch.isfinite(l3)
    # no CPU fallback
    from fudanocr_b200 import _lib as L
    with pytest.raises(L.FocrError):
        model(image.cpu(), length, text_input)


def test_sld_graph_replay_follows_the_eager_trajectory_and_dropout_epoch_advances():
    """SLDTrainer replays the step as one CUDA graph per input shape (first step eager, second captured): with dropout off the
    graph trajectory must equal the eager one (text length 5 is padded to the bucket of 8 inside the graph: same loss, same
    gradients); the dropout epoch the graph advances changes the masks of a fixed (seed, stream) and 0 restores them"""
    SO, g, sd, model, image, length, text_input, text_gt = _setup()
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.model import recog_ops as ops
    from fudanocr_b200.model.transformer import Transformer
    from fudanocr_b200.trainer_sld import SLDTrainer
    twin = Transformer("stroke")
    twin.load_state_dict(sd, strict=False)
    twin = twin.to(DEV)
    for m in (model, twin):
        m.train()
        m.dropout_p = 0.0
    assert text_input.shape[1] % 8 != 0                      # the padding path is exercised
    eager, graph = SLDTrainer(twin, use_graph=False), SLDTrainer(model, use_graph=True)
    for it in range(4):
        le = eager.step(image, length, text_input, text_gt)
        lg = graph.step(image, length, text_input, text_gt)
        torch.cuda.synchronize()
        # (row order of the decoder's token matrix changes with the padding: reductions differ in the last bits, and Adadelta's
        # first steps are sign-like, so the two runs agree to rounding noise, not bit for bit)
        assert abs(float(le) - float(lg)) < 2e-3 * abs(float(le)) + 1e-6, (it, float(le), float(lg))
    assert len(graph._graphs) == 1 and graph.kernel_launches > 4 * 300
    worst = max(_rel(a.detach(), b.detach()) for (_, a), (_, b) in zip(model.named_parameters(), twin.named_parameters()))
    assert worst < 5e-2, worst
    # a changed dropout rate is a new graph key, not a stale replay
    model.dropout_p = 0.1
    l5 = graph.step(image, length, text_input, text_gt)
    l6 = graph.step(image, length, text_input, text_gt)
    l7 = graph.step(image, length, text_input, text_gt)
    torch.cuda.synchronize()
    assert len(graph._graphs) == 2 and all(torch.isfinite(x) for x in (l5, l6, l7))
    # the epoch word: same (seed, stream) -> different mask after an advance, the original mask again at epoch 0
    x = torch.ones(1 << 16, dtype=torch.bfloat16, device=DEV)
    L.check(L.lib.focr_recog_epoch_set(0, L.cur_stream()))
    m0 = ops.dropout(x, 0.1, 1234, 3)
    L.check(L.lib.focr_recog_epoch_advance(L.cur_stream()))
    m1 = ops.dropout(x, 0.1, 1234, 3)
    L.check(L.lib.focr_recog_epoch_set(0, L.cur_stream()))
    m2 = ops.dropout(x, 0.1, 1234, 3)
    torch.cuda.synchronize()
    assert torch.equal(m0, m2) and not torch.equal(m0, m1)
    assert abs(float((m1 == 0).float().mean()) - 0.1) < 0.01


# ---- batch 32: the conditioned regime (BatchNorm over 8192 positions per channel) ----------------------------------------------------
def _setup_b32():
    from oracle import sld_oracle as SO, synth
    from fudanocr_b200.model.transformer import Transformer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.load(synth.GOLDEN_DIR / "sld_b32.pt", weights_only=False)
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    model = Transformer("stroke")
    model.load_state_dict(sd, strict=False)
    model = model.to(DEV)
    image, strings = SO.synth_batch(g["B"])
    assert strings == g["strings"] and abs(float(image.double().sum()) - g["image_checksum"]) < 1e-6
    return SO, g, sd, model, image.to(DEV), g["length"].to(DEV), g["text_input"].to(DEV), g["text_gt"].to(DEV)


def test_sld_encoder_teacher_forced_per_stage_b32():
    """every stage of the 40-conv train-mode-BatchNorm encoder on its own: the engine stage fed the ORACLE's input activation
    (bf16-rounded) against the oracle's output of that stage - absolute tolerance 2e-2 relative L2 (bf16, north_star 1e-2 per
    op, two to three convs + BatchNorms per stage).  No chaos can enter: each stage starts from the exact input."""
    SO, g, sd, model, image, length, text_input, text_gt = _setup_b32()
    worst = _encoder_stage_errors(SO, model, sd, image)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sld_encoder_stages_b32.json", "w") as f:
        json.dump(worst, f)
    bad = {k: v for k, v in worst.items() if not v < 2e-2}
    assert not bad, bad


def _encoder_stage_errors(SO, model, sd, image):
    dsd = {k: v.to(DEV) for k, v in sd.items()}
    from fudanocr_b200.model import recog_ops as ops
    from fudanocr_b200.model.transformer import _Conv, _ConvFirst, _MaxPool
    model.train()
    taps = {}
    with torch.no_grad():
        SO.encoder(dsd, image, train=True, taps=taps)
        e = model.encoder

        def nhwc(x):
            return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        worst = {}
        xin, xout = taps["stem"]
        y = _MaxPool.apply(model._bn(_ConvFirst.apply(xin.float().contiguous(), e.conv1.weight, e.conv1.bias), e.bn1, ops.ACT_RELU))
        worst["stem"] = _rel(y.float().permute(0, 3, 1, 2), xout)
        xin, xout = taps["conv2"]
        y = model._bn(_Conv.apply(nhwc(xin), e.conv2.weight, e.conv2.bias), e.bn2, ops.ACT_RELU)
        worst["conv2"] = _rel(y.float().permute(0, 3, 1, 2), xout)
        for name, tail in (("layer1", "layer1"), ("layer2", "layer2"), ("layer3", "layer3"), ("layer4", "layer4_conv2")):
            for i, blk in enumerate(getattr(e, name)):
                xin, xout = taps[f"{name}.{i}"]
                y = model._block(nhwc(xin), blk)
                worst[f"{name}.{i}"] = _rel(y.float().permute(0, 3, 1, 2), xout)
            tname = name + ("_conv2" if name == "layer4" else "_conv")
            conv = getattr(e, tail + ("_conv" if name != "layer4" else ""))
            bn = getattr(e, tail + "_bn")
            xin, xout = taps[tname]
            y = model._bn(_Conv.apply(nhwc(xin), conv.weight, conv.bias), bn, ops.ACT_RELU)
            worst[tname] = _rel(y.float().permute(0, 3, 1, 2), xout)
    return worst


def _train_oracle(SO_loss, sd, steps, wd=0.0):
    """conditioning: `steps` Adadelta steps of the fp32 ORACLE on the GPU (test infrastructure).  The synthetic random weights
    put every prediction at the uniform distribution (loss = ln 7): the gradient there is a difference of nearly equal terms
    and NO bf16 evaluation follows it (stock autocast is ~85 % off, measured).  After a few dozen steps the network depends on
    its input and stock bf16 is within a few per cent - the regime a training run lives in, and the one absolute tolerances
    mean something in."""
    from oracle import sld_oracle as SO
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False   # same trajectory on every run
    params = {k: v.to(DEV).clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    full = {k: v.to(DEV) for k, v in sd.items()}
    full.update(params)
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}
    for _ in range(steps):
        for v in params.values():
            v.grad = None
        stats = {}
        loss = SO_loss(full, stats)
        loss.backward()
        with torch.no_grad():
            for k, v in params.items():
                if v.grad is None:
                    continue
                p2, sq, acc = SO.adadelta_update(v, v.grad, *state[k], wd=wd)
                v.copy_(p2)
                state[k] = (sq, acc)
            for k, v in stats.items():
                full[k] = v
    return {k: v.detach().clone() for k, v in full.items()}


def _grad_report(ref_g, amp_g, eng_g):
    """per-tensor relative L2 errors of the engine and of stock autocast against fp32, over the tensors that carry gradient
    (a conv bias ahead of a train-mode BatchNorm has true gradient 0: only rounding noise is left to compare)"""
    gmax = max(float(r.norm()) for r in ref_g.values())
    tensors = {}
    for k, r in ref_g.items():
        if float(r.norm()) < 1e-4 * gmax:
            continue
        tensors[k] = [_rel(eng_g[k], r), _rel(amp_g[k], r)]
    es = sorted(v[0] for v in tensors.values())
    ss = sorted(v[1] for v in tensors.values())
    return {"tensors": tensors, "median": [es[len(es) // 2], ss[len(ss) // 2]], "p90": [es[int(0.9 * len(es))], ss[int(0.9 * len(ss))]],
            "worst": [es[-1], ss[-1]], "n": len(es)}


def test_sld_train_step_b32_absolute_tolerances():
    """whole step at batch 32 against the fp32 oracle, ABSOLUTE tolerances, at conditioned weights.
    (1) pin: at the synthetic weights the GPU fp32 oracle reproduces loss and every gradient norm the UNMODIFIED reference module
    gave at this batch (tests/golden/sld_b32.pt) - and there the engine's loss agrees to 2e-3 although no bf16 gradient can;
    (2) 100 Adadelta steps of the oracle condition the network (_train_oracle); at those weights the engine's loss is within 2e-3,
    the median per-tensor gradient error within 6e-2, the 90th percentile within 0.2 - asserted as numbers; the stock-autocast
    figures ride along as a report (gpurun_out/sld_parity_b32.json)."""
    SO, g, sd, model, image, length, text_input, text_gt = _setup_b32()
    model.train()
    model.dropout_p = 0.0

    def run(weights, autocast):
        osd = {k: v.to(DEV).clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in weights.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss, logits, amap, conv = SO.loss_fn(osd, image, length, text_input, text_gt)
        loss.float().backward()
        return loss.detach().float(), {k: v.grad.float() for k, v in osd.items() if v.grad is not None}
    ref_loss, ref_g = run(sd, False)
    assert abs(float(ref_loss) - float(g["loss"])) < 1e-3 * float(g["loss"])       # GPU fp32 oracle == reference (CPU) value
    nmax = max(float(n) for n in g["grad_norms"].values() if n is not None)
    for k, n in g["grad_norms"].items():
        if n is not None and float(n) > 1e-4 * nmax:   # (a conv bias ahead of a train-mode BatchNorm: true gradient 0, noise only)
            assert abs(float(ref_g[k].norm()) - float(n)) < 2e-2 * float(n), k   # ... and its gradients
    loss0 = model.loss(image, length, text_input, text_gt)
    assert abs(float(loss0) - float(ref_loss)) < 2e-3 * float(ref_loss)
    # (2) conditioned weights
    trained = _train_oracle(lambda full, stats: SO.loss_fn(full, image, length, text_input, text_gt, None, stats)[0], sd, 100)
    model.load_state_dict({k: v.cpu() for k, v in trained.items()}, strict=False)
    model.zero_grad(set_to_none=True)
    ref_loss, ref_g = run(trained, False)
    amp_loss, amp_g = run(trained, True)
    loss = model.loss(image, length, text_input, text_gt)
    loss.backward()
    eng_g = {k: p.grad.float() for k, p in model.named_parameters() if p.grad is not None}
    report = _grad_report(ref_g, amp_g, eng_g)
    report["loss"] = [float(loss), float(ref_loss), float(amp_loss)]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sld_parity_b32.json", "w") as f:
        json.dump(report, f)
    assert float(ref_loss) < 1.8, float(ref_loss)                  # the conditioning did train (ln 7 = 1.946 at the start)
    assert abs(float(loss) - float(ref_loss)) < 2e-3 * float(ref_loss), report["loss"]
    assert report["median"][0] < 6e-2, (report["median"], report["p90"], report["worst"])
    assert report["p90"][0] < 0.2, (report["median"], report["p90"], report["worst"])


# ---- 32 x 320 crops (BASELINE configs[3]): 16 x 160 feature maps, 2 560 image tokens ------------------------------------------------
def _setup_wide(B=8):
    from oracle import sld_oracle as SO, synth
    from fudanocr_b200.model.transformer import Transformer
    from fudanocr_b200.util_recog import ALPHABET_STROKE, converter_sld
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    model = Transformer("stroke")
    model.load_state_dict(sd, strict=False)
    model = model.to(DEV)
    parts = [SO.synth_batch(B, seed=1234 + 17 * i) for i in range(10)]
    image = torch.cat([p[0] for p in parts], dim=3).to(DEV)                       # (B, 3, 32, 320)
    strings = parts[0][1]
    length, text_input, text_gt, _ = converter_sld("character", strings, alp2num_character={c: i for i, c in enumerate(ALPHABET_STROKE)},
                                                   device=DEV)
    return SO, sd, model, image, length, text_input, text_gt


def test_sld_32x320_encoder_stages_and_decoder():
    """the fully-convolutional encoder on 32 x 320 crops (stroke-level-decomposition/model/transformer.py:126-164 has no size in
    it): 16 x 160 maps cut into 32 x 4-pixel TMA boxes, every stage teacher-forced against the fp32 oracle (2e-2 as at 32 x 32);
    then the decoder on the engine's own 2 560-token feature map against the oracle decoder on the same features: logits and the
    (B, 4, T, 2560) attention map"""
    SO, sd, model, image, length, text_input, text_gt = _setup_wide()
    assert image.shape[2:] == (32, 320)
    model.train()
    model.dropout_p = 0.0
    worst = _encoder_stage_errors(SO, model, sd, image)
    dsd = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        out = model(image, length, text_input, test=True)
        assert out["conv"].shape == (image.shape[0], 1024, 16, 160) and out["map"].shape[-1] == 2560
        ref_logits, ref_map, _ = SO.forward(dsd, image, text_input, True, conv_feature=out["conv"].float())
    worst["decoder_logits"] = _rel(out["pred"].float(), ref_logits)
    worst["decoder_map_maxabs"] = float((out["map"] - ref_map).abs().max())
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sld_32x320_stages.json", "w") as f:
        json.dump(worst, f)
    bad = {k: v for k, v in worst.items() if not v < 2e-2}
    assert not bad, bad


def test_sld_32x320_train_step_absolute_tolerances():
    """whole step on 32 x 320 crops against the fp32 oracle at conditioned weights (60 oracle Adadelta steps), the same absolute
    tolerances as the 32 x 32 case: loss 2e-3, median per-tensor gradient error 6e-2, 90th percentile 0.2"""
    SO, sd, model, image, length, text_input, text_gt = _setup_wide()
    model.train()
    model.dropout_p = 0.0

    def run(weights, autocast):
        osd = {k: v.to(DEV).clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in weights.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss, logits, amap, conv = SO.loss_fn(osd, image, length, text_input, text_gt)
        loss.float().backward()
        return loss.detach().float(), {k: v.grad.float() for k, v in osd.items() if v.grad is not None}
    trained = _train_oracle(lambda full, stats: SO.loss_fn(full, image, length, text_input, text_gt, None, stats)[0], sd, 60)
    model.load_state_dict({k: v.cpu() for k, v in trained.items()}, strict=False)
    model.zero_grad(set_to_none=True)
    ref_loss, ref_g = run(trained, False)
    amp_loss, amp_g = run(trained, True)
    loss = model.loss(image, length, text_input, text_gt)
    loss.backward()
    eng_g = {k: p.grad.float() for k, p in model.named_parameters() if p.grad is not None}
    report = _grad_report(ref_g, amp_g, eng_g)
    report["loss"] = [float(loss), float(ref_loss), float(amp_loss)]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sld_parity_32x320.json", "w") as f:
        json.dump(report, f)
    assert abs(float(loss) - float(ref_loss)) < 2e-3 * float(ref_loss), report["loss"]
    assert report["median"][0] < 6e-2, (report["median"], report["p90"], report["worst"])
    assert report["p90"][0] < 0.2, (report["median"], report["p90"], report["worst"])
