"""TSRN (vertical + horizontal BiGRU residual blocks) on the focr engine vs the oracle / reference golden.  GPU only.
Tolerances as in test_gpu_tbsrn.py: eval SR relative L2 <= 1e-2; train mode at least as close as stock autocast(bf16)."""
import ctypes as C
import json
import os
import statistics

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
REPORT = {}


def _dump():
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(REPORT, open("gpurun_out/tsrn_parity.json", "w"), indent=1)


@pytest.fixture(scope="module")
def env():
    from oracle import synth, tbsrn_oracle as O, tsrn_oracle as TS
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.model.tsrn import TSRN
    sd = synth.synth_state_dict(synth.load_spec("tsrn"), 2468, O.tps_buffers())
    golden = torch.load(synth.GOLDEN_DIR / "tsrn_b4.pt", weights_only=False)
    return dict(synth=synth, O=O, TS=TS, L=L, TSRN=TSRN, sd=sd, golden=golden)


def _model(env, stn, train=True):
    m = env["TSRN"](STN=stn).to(DEV)
    m.load_state_dict({k: v for k, v in env["sd"].items() if stn or not (k.startswith("stn_head") or k.startswith("tps"))})
    m.train(train)
    return m


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _ws(env, m, B, name, shape):
    L = env["L"]
    off, n, eb = C.c_longlong(), C.c_longlong(), C.c_int()
    L.check(L.lib.focr_tsrn_ws_tensor(B, m.srb_nums, name.encode(), C.byref(off), C.byref(n), C.byref(eb)))
    ws = m._ws[B]
    raw = ws[off.value: off.value + n.value * eb.value]
    return raw.view(torch.bfloat16).view(*shape).permute(0, 3, 1, 2).float()


def test_eval_forward_vs_golden(env):
    m = _model(env, True, train=False)
    lr, _ = env["synth"].synth_images(4)
    with torch.no_grad():
        sr = m(lr.to(DEV))
    rel = _rel(sr.cpu(), env["golden"]["eval_sr"])
    REPORT["eval_sr_rel_l2"] = rel
    _dump()
    assert rel < 1e-2, rel


@pytest.mark.parametrize("stn,B", [(False, 4), (True, 64)])
def test_train_forward_backward_vs_oracle(env, stn, B):
    TS, synth = env["TS"], env["synth"]
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    m = _model(env, stn)
    sr = m(lr)
    loss = F.mse_loss(sr, hr)
    (loss * 100).backward()
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = {k: v.to(DEV) for k, v in env["sd"].items()}
    taps = {}
    _, info = TS.train_step(sd, lr, hr, {}, stn=stn, taps=taps)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, cal = TS.train_step(sd, lr, hr, {}, stn=stn)
    rep = {"inter": {}}
    for i in range(5):
        for f in ("r0", "o1", "out"):
            rep["inter"][f"srb{i}.{f}"] = _rel(_ws(env, m, B, f"srb{i}.{f}", (B, 16, 64, 64)), taps[f"block{i + 2}.{f}"])
    rep["sr_rel_l2"] = _rel(sr.detach(), info["sr"])
    rep["torch_bf16_sr_rel_l2"] = _rel(cal["sr"].float(), info["sr"])
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    rep["grad_missing"] = sorted(set(info["grads"]) - set(grads))
    gmax = max(g.norm().item() for g in info["grads"].values())
    rel = {k: _rel(grads[k], g) for k, g in info["grads"].items() if k in grads and g.norm().item() >= 1e-5 * gmax}
    calrel = {k: _rel(cal["grads"][k], info["grads"][k]) for k in rel}
    zero = {k: grads[k].norm().item() / gmax for k, g in info["grads"].items() if k in grads and g.norm().item() < 1e-5 * gmax}
    rep["grad_rel_l2"], rep["torch_bf16_grad_rel_l2"], rep["zero_grad_abs"] = rel, calrel, zero
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    rep["grad_norm"] = [gn, info["grad_norm"].item()]
    rep["loss"] = [loss.item(), info["mse"].item()]
    REPORT[f"train_stn{int(stn)}"] = rep
    _dump()
    assert not rep["grad_missing"]
    assert rep["sr_rel_l2"] < 2.5e-2 and rep["sr_rel_l2"] <= 1.1 * max(rep["torch_bf16_sr_rel_l2"], 5e-3), rep["sr_rel_l2"]
    assert abs(loss.item() - info["mse"].item()) < 5e-3 * info["mse"].item()
    trunk = {k: v for k, v in rel.items() if not k.startswith("stn_head.")}
    # per tensor: within 10 %, or (cancellation-dominated sums such as the single PReLU slope, where stock
    # autocast(bf16) is > 100 % off) at most half the stock-bf16 error
    for k, v in trunk.items():
        assert v < 0.1 or v < 0.5 * calrel[k], (k, v, calrel[k])
    assert statistics.median(trunk.values()) <= 1.5 * max(statistics.median(calrel[k] for k in trunk), 5e-3)
    assert max(zero.values(), default=0.0) < 1e-2
    assert abs(gn - info["grad_norm"].item()) < 0.03 * info["grad_norm"].item()


@pytest.mark.parametrize("stn", [False, True])
def test_train_forward_backward_conditioned_absolute(env, stn):
    """north_star's 1e-2 bf16 contract asserted absolutely in train mode, at weights conditioned by 50 deterministic Adam steps of
    the fp32 oracle (see tests/test_gpu_tbsrn.py::test_train_forward_backward_conditioned_absolute for the reasoning)"""
    TS, synth = env["TS"], env["synth"]
    B = 32
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    det, bench_ = torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    try:
        sd = {k: v.to(DEV) for k, v in env["sd"].items()}
        state = {}
        for _ in range(50):
            sd, _info = TS.train_step(sd, lr, hr, state, stn=True)
        _, info = TS.train_step(sd, lr, hr, {}, stn=stn)
    finally:
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = det, bench_
    m = env["TSRN"](STN=stn).to(DEV)
    m.load_state_dict({k: v for k, v in sd.items() if stn or not (k.startswith("stn_head") or k.startswith("tps"))})
    m.train()
    sr = m(lr)
    loss = F.mse_loss(sr, hr)
    (loss * 100).backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    gmax = max(g.norm().item() for g in info["grads"].values())
    rel = {k: _rel(grads[k], g) for k, g in info["grads"].items()
           if k in grads and g.norm().item() >= 1e-4 * gmax and not k.startswith("stn_head.")}
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    rep = {"sr_rel_l2": _rel(sr.detach(), info["sr"]), "loss": [loss.item(), info["mse"].item()],
           "grad_rel_median": statistics.median(rel.values()), "grad_rel_max": max(rel.values()),
           "grad_rel_worst": max(rel, key=rel.get), "grad_norm": [gn, info["grad_norm"].item()]}
    REPORT[f"train_conditioned_stn{int(stn)}"] = rep
    _dump()
    assert rep["sr_rel_l2"] < 1e-2, rep
    assert abs(loss.item() - info["mse"].item()) < 5e-3 * info["mse"].item(), rep
    assert rep["grad_rel_median"] < 8e-2, rep      # measured 4.5e-2 / 6.5e-2 (the BiGRU recurrences; TBSRN: 3.0e-2)
    assert abs(gn - info["grad_norm"].item()) < 0.03 * info["grad_norm"].item(), rep


def test_trainer_steps_and_full_size(env):
    """fused step on TSRN: loss falls, bit-reproducible, B = 256 (BASELINE configs[2] uses TSRN at batch 256)"""
    from fudanocr_b200.trainer import TBSRNTrainer
    torch.manual_seed(0)
    B = 256
    lr = torch.rand(B, 3, 16, 64, device=DEV)
    hr = F.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False)
    runs = []
    for rep in range(2):
        m = _model(env, True)
        tr = TBSRNTrainer(m, lr=1e-3)
        runs.append([tr.step(lr, hr).item() for _ in range(5)])
    REPORT["b256_losses"] = runs
    _dump()
    assert runs[0] == runs[1] and runs[0][-1] < runs[0][0]
