"""image-ids-CTR recogniser on the engine (fudanocr_b200/model/ids_transformer.py, trainer_sld.py:IDSTrainer) vs
the oracle restatement (pinned to the unmodified reference module and train.py:63-80 by tests/golden/ids_b4.pt) on the GPU.
Same layered criteria as tests/test_gpu_sld.py: kernels in test_gpu_recog_ops.py, assembly in test_ids_assembly.py, here the
eval forward and the whole train step tensor by tensor against what stock autocast(bf16) does to the same restatement."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def _setup():
    from oracle import ids_oracle as IO, synth
    from fudanocr_b200.model.ids_transformer import Transformer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.load(synth.GOLDEN_DIR / "ids_b4.pt", weights_only=False)
    sd = synth.synth_state_dict(synth.load_spec("ids"), 4321)
    model = Transformer()
    model.load_state_dict(sd, strict=False)
    model = model.to(DEV)
    image, labels = IO.synth_batch(g["B"])
    assert labels == g["labels"]
    tf = IO.synth_text_features().to(DEV)
    return IO, g, sd, model, image.to(DEV), g["length"].to(DEV), g["text_input"].to(DEV), g["text_gt"].to(DEV), tf


def test_ids_train_step_vs_oracle_calibrated_against_stock_bf16():
    IO, g, sd, model, image, length, text_input, text_gt, tf = _setup()
    model.train()
    model.dropout_p = 0.0

    def run(autocast):
        osd = {k: v.to(DEV).clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss, rec, dis, *_ = IO.loss_fn(osd, image, length, text_input, text_gt, tf)
        loss.float().backward()
        return [float(loss.detach()), float(rec.detach()), float(dis.detach())], {k: v.grad.float() for k, v in osd.items() if v.grad is not None}
    ref_l, ref_g = run(False)
    amp_l, amp_g = run(True)
    assert abs(ref_l[0] - float(g["loss"])) < 1e-3 * float(g["loss"])      # the GPU fp32 oracle reproduces the reference's value
    loss, rec, dis = model.loss(image, length, text_input, text_gt, tf)
    loss.backward()
    eng_g = {k: p.grad.float() for k, p in model.named_parameters() if p.grad is not None}
    assert set(eng_g) == set(ref_g)                                        # layer4 / compress_attention_linear stay grad-less
    report = {"loss": [[float(loss), float(rec), float(dis)], ref_l, amp_l], "tensors": {}}
    bad = []
    for k, r in ref_g.items():
        if float(r.abs().max()) < 1e-6:
            continue
        e, s = _rel(eng_g[k], r), _rel(amp_g[k], r)
        report["tensors"][k] = [e, s]
        if not (e < max(1.5 * s, 5e-2)):
            bad.append((k, e, s))
    es = sorted(v[0] for v in report["tensors"].values())
    ss = sorted(v[1] for v in report["tensors"].values())
    report["median"], report["worst"] = [es[len(es) // 2], ss[len(ss) // 2]], [es[-1], ss[-1]]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ids_parity.json", "w") as f:
        json.dump(report, f)
    assert abs(float(loss) - ref_l[0]) < max(2 * abs(amp_l[0] - ref_l[0]), 2e-2 * ref_l[0])
    assert abs(float(dis) - ref_l[2]) < max(2 * abs(amp_l[2] - ref_l[2]), 2e-2 * abs(ref_l[2]))
    assert not bad, bad[:8]
    assert report["median"][0] < 1.15 * report["median"][1] + 1e-2, report["median"]


def test_ids_eval_forward_fused_trainer_and_reference_loop():
    IO, g, sd, model, image, length, text_input, text_gt, tf = _setup()
    from fudanocr_b200.model.ids_transformer import Transformer
    from fudanocr_b200.trainer_sld import IDSTrainer
    dsd = {k: v.to(DEV) for k, v in sd.items()}
    model.eval()
    with torch.no_grad():
        o_pred, o_map, o_conv = IO.forward(dsd, image, text_input, train=False)
        ev = model(image, length, text_input, test=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            a_pred, _, a_conv = IO.forward(dsd, image, text_input, train=False)
        e_conv, s_conv, e_pred, s_pred = _rel(ev["conv"], o_conv), _rel(a_conv, o_conv), _rel(ev["pred"], o_pred), _rel(a_pred, o_pred)
        assert e_conv < max(1.5 * s_conv, 2e-2) and e_pred < max(1.5 * s_pred, 2e-2), (e_conv, s_conv, e_pred, s_pred)
        assert ev["pred"].shape == (g["B"], text_input.shape[1], 2048) and ev["conv"].shape == (g["B"], 1024, 2, 16)
    # the unchanged reference loop (train.py:63-90) on the drop-in module vs the fused trainer
    model.train()
    model.dropout_p = 0.0
    twin = Transformer()
    twin.load_state_dict(sd, strict=False)
    twin = twin.to(DEV).train()
    twin.dropout_p = 0.0
    opt = torch.optim.Adadelta(twin.parameters(), lr=1.0, rho=0.9, weight_decay=1e-4)
    opt.zero_grad()
    reg = torch.cat([tf[item].unsqueeze(0) for item in text_gt], dim=0)
    text_pred = twin(image, length, text_input)["pred"]
    text_pred = text_pred / text_pred.norm(dim=1, keepdim=True)
    final_res = text_pred @ tf.t()
    l_ref = torch.nn.CrossEntropyLoss()(final_res, text_gt) + 0.001 * (-torch.nn.MSELoss()(text_pred, reg))
    l_ref.backward()
    opt.step()
    tr = IDSTrainer(model, tf)
    l_fused = tr.step(image, length, text_input, text_gt)
    torch.cuda.synchronize()
    # the reference loop normalises / multiplies in fp32 torch ops, the fused path in bf16 kernels: 1e-3 on the loss value
    assert abs(float(l_fused) - float(l_ref.detach())) < 2e-3 * abs(float(l_ref.detach()))
    # both took the same step: Adadelta's first update is ~3e-3 * sign(g) per weight, so a wiring difference would show as O(0.1)
    worst = 0.0
    for (k, a), (_, b) in zip(model.named_parameters(), twin.named_parameters()):
        worst = max(worst, _rel(a.detach(), b.detach()))
    assert worst < 2e-2, worst
    # (no monotonicity claim: at lr 1 that first step overshoots on a repeated 4-crop batch - measured 8.29 -> 22.1 here, while
    # the batch-64 run of scripts/bench_cfg4.py --model ids goes 8.3 -> 6.65 in nine steps)
    l2 = tr.step(image, length, text_input, text_gt, lr=0.5)
    torch.cuda.synchronize()
    assert torch.isfinite(l2)
    # a per-step learning rate is part of the graph key: first step at 0.25 eager, second captured, third replayed
    for _ in range(3):
        l3 = tr.step(image, length, text_input, text_gt, lr=0.25)
    torch.cuda.synchronize()
    assert torch.isfinite(l3) and torch.isfinite(tr.loss_rec) and any(k[-1] == 0.25 for k in tr._graphs)
    model.dropout_p = 0.1
    assert torch.isfinite(tr.step(image, length, text_input, text_gt))
    names = set(tr.names)
    assert not any("layer4" in k or "compress_attention_linear" in k for k in names) and "encoder.layer3_conv.weight" in names


def test_ids_train_step_b32_absolute_tolerances():
    """whole step at batch 32 against the fp32 oracle, ABSOLUTE tolerances (see tests/test_gpu_sld.py for the reasoning): pinned to
    the unmodified reference module + train.py:71-80 at this batch by tests/golden/ids_b32.pt at the synthetic weights; then, at
    weights conditioned by 150 oracle Adadelta steps (deterministic), loss terms within 2e-3, median per-tensor gradient error
    within 0.12, 90th percentile within 0.35 - looser than the stroke-level model's 6e-2 / 0.2 because 150 steps move this
    4303-way similarity loss only from 8.25 to 6.19 and stock autocast is itself 0.107 / 0.33 off fp32 there (measured: engine
    0.105 / 0.316); stock-autocast figures recorded as a report only."""
    from oracle import ids_oracle as IO, synth
    from fudanocr_b200.model.ids_transformer import Transformer
    from test_gpu_sld import _grad_report, _train_oracle
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.load(synth.GOLDEN_DIR / "ids_b32.pt", weights_only=False)
    sd = synth.synth_state_dict(synth.load_spec("ids"), 4321)
    image, labels = IO.synth_batch(g["B"])
    assert labels == g["labels"]
    image, tf = image.to(DEV), IO.synth_text_features().to(DEV)
    length, text_input, text_gt = g["length"].to(DEV), g["text_input"].to(DEV), g["text_gt"].to(DEV)

    def run(weights, autocast):
        osd = {k: v.to(DEV).clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in weights.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss, rec, dis, *_ = IO.loss_fn(osd, image, length, text_input, text_gt, tf)
        loss.float().backward()
        return [float(loss.detach()), float(rec.detach()), float(dis.detach())], {k: v.grad.float() for k, v in osd.items() if v.grad is not None}
    ref_l, ref_g = run(sd, False)
    assert abs(ref_l[0] - float(g["loss"])) < 1e-3 * float(g["loss"])
    nmax = max(float(n) for n in g["grad_norms"].values() if n is not None)
    for k, n in g["grad_norms"].items():
        if n is not None and float(n) > 1e-4 * nmax:   # (a conv bias ahead of a train-mode BatchNorm: true gradient 0, noise only)
            assert abs(float(ref_g[k].norm()) - float(n)) < 2e-2 * float(n), k
    trained = _train_oracle(lambda full, stats: IO.loss_fn(full, image, length, text_input, text_gt, tf, stats)[0], sd, 150, wd=1e-4)
    model = Transformer()
    model.load_state_dict({k: v.cpu() for k, v in trained.items()}, strict=False)
    model = model.to(DEV).train()
    model.dropout_p = 0.0
    ref_l, ref_g = run(trained, False)
    amp_l, amp_g = run(trained, True)
    loss, rec, dis = model.loss(image, length, text_input, text_gt, tf)
    loss.backward()
    eng_g = {k: p.grad.float() for k, p in model.named_parameters() if p.grad is not None}
    report = _grad_report(ref_g, amp_g, eng_g)
    report["loss"] = [[float(loss), float(rec), float(dis)], ref_l, amp_l]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ids_parity_b32.json", "w") as f:
        json.dump(report, f)
    assert abs(float(loss) - ref_l[0]) < 2e-3 * abs(ref_l[0]) and abs(float(rec) - ref_l[1]) < 2e-3 * abs(ref_l[1]), report["loss"]
    assert report["median"][0] < 0.12, (report["median"], report["p90"], report["worst"])
    assert report["p90"][0] < 0.35, (report["median"], report["p90"], report["worst"])
