"""Whole-network parity of the focr TBSRN engine (through the C-ABI / the drop-in nn.Module) against the
oracle restatement and the golden vectors recorded from the real reference modules.  GPU only.

Tolerances.  north_star: SR within 1e-2 relative (bf16 compute) of the fp32 reference.
  * eval mode (the deployed path): relative L2 of SR <= 1e-2, asserted against the golden vectors.
  * train mode (batch-statistics BatchNorm re-normalising bf16 rounding noise 11 times, synthetic gain-1 weights):
    the same restatement run by stock PyTorch under autocast(bf16) lands 1.9e-2 from its own fp32 result; the
    engine must be at least as close as that (x1.1) and within 2.5e-2.  Gradients: per-tensor relative L2 <= 0.1 and
    <= stock-bf16 level, total gradient norm within 3 %.
  * the STN prologue (BatchNorm over B*H*W <= 2B samples in its last layers, BatchNorm1d over B) is ill-conditioned at
    tiny batch (stock autocast(bf16) is 100 % off at B = 4), so its parity case runs at B = 64."""
import ctypes as C
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
REPORT = {}


def _dump():
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tbsrn_parity.json", "w") as f:
        json.dump(REPORT, f, indent=1)


@pytest.fixture(scope="module")
def env():
    from oracle import synth, tbsrn_oracle as O
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.model.tbsrn import TBSRN
    sd = synth.synth_state_dict(synth.load_spec("tbsrn"), 1234, O.tps_buffers())
    golden = torch.load(synth.GOLDEN_DIR / "tbsrn_b4.pt", weights_only=False)
    return dict(synth=synth, O=O, L=L, TBSRN=TBSRN, sd=sd, golden=golden)


def _model(env, p_drop=0.0, train=True, stn=True):
    m = env["TBSRN"](STN=stn).to(DEV)
    m.load_state_dict({k: v for k, v in env["sd"].items() if stn or not (k.startswith("stn_head") or k.startswith("tps"))})
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = p_drop
    m.train(train)
    return m


def _ws_tensor(env, model, B, name):
    L = env["L"]
    off, n, eb = C.c_longlong(), C.c_longlong(), C.c_int()
    L.check(L.lib.focr_tbsrn_ws_tensor(B, model.srb_nums, name.encode(), C.byref(off), C.byref(n), C.byref(eb)))
    ws = model._ws[B]
    base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
    raw = ws[base + off.value: base + off.value + n.value * eb.value]
    return raw.view(torch.bfloat16 if eb.value == 2 else torch.float32)


def _nchw(t, B, H, W, Cc):
    return t.view(B, H, W, Cc).permute(0, 3, 1, 2).float()


def _rel_l2(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _oracle_bf16(env, sd, lr, hr, stn):
    """calibration: the same restatement under torch autocast(bf16) — how far a stock bf16 PyTorch run of the
    reference algorithm lands from its own fp32 result"""
    O = env["O"]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, info = O.train_step(sd, lr, hr, {}, masks=None, stn=stn)
    return info


def _grad_report(grads, ref, cal=None):
    """per-tensor relative L2 vs the fp32 oracle; tensors whose true gradient is (numerically) zero — conv biases
    in front of a train-mode BatchNorm, the key bias of softmax attention — are judged on absolute size"""
    gmax = max(g.norm().item() for g in ref.values())
    rel, zero = {}, {}
    for k, g in ref.items():
        if g.norm().item() < 1e-5 * gmax:
            zero[k] = grads[k].norm().item() / gmax
        else:
            rel[k] = _rel_l2(grads[k], g)
    out = {"rel_l2": rel, "zero_grad_abs_over_gmax": zero}
    if cal is not None:
        out["torch_bf16_rel_l2"] = {k: _rel_l2(cal[k], ref[k]) for k in rel}
    return out


def test_eval_forward_vs_golden(env):
    m = _model(env, train=False)
    lr, hr = env["synth"].synth_images(4)
    with torch.no_grad():
        sr = m(lr.to(DEV))
    ref = env["golden"]["eval_sr"]
    rel = _rel_l2(sr.cpu(), ref)
    REPORT["eval"] = {"sr_rel_l2": rel, "sr_maxabs": (sr.cpu() - ref).abs().max().item()}
    _dump()
    assert sr.shape == (4, 3, 32, 128) and rel < 1e-2, REPORT["eval"]


@pytest.mark.parametrize("stn", [False, True])
def test_train_forward_backward_vs_oracle(env, stn):
    O, synth = env["O"], env["synth"]
    B = 64 if stn else 4
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    m = env["TBSRN"](STN=stn).to(DEV)
    m.load_state_dict({k: v for k, v in env["sd"].items() if stn or not (k.startswith("stn_head") or k.startswith("tps"))})
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train()
    sr = m(lr)
    loss = F.mse_loss(sr, hr)
    (loss * 100).backward()
    torch.cuda.synchronize()
    _strict_fp32()
    sd = {k: v.to(DEV) for k, v in env["sd"].items()}
    taps = {}
    _, info = O.train_step(sd, lr, hr, {}, masks=None, stn=stn, taps=taps)
    cal = _oracle_bf16(env, sd, lr, hr, stn)
    rep = {}
    inter = {}
    if stn:
        inter["ctrl_maxabs"] = (_ws_tensor(env, m, B, "ctrl").view(-1, 64)[:B, :40]
                                - taps["ctrl"].reshape(B, 40)).abs().max().item()
        for i, c in enumerate([(16, 64, 32, 64), (8, 32, 64, 64), (4, 16, 128, 128), (2, 8, 256, 256), (1, 4, 256, 256),
                               (1, 2, 256, 256)]):
            h, w_, co, npad = c
            ya = _ws_tensor(env, m, B, f"stn.yact{i}").view(B, h, w_, co).permute(0, 3, 1, 2).float()
            yp = _ws_tensor(env, m, B, f"stn.ypre{i}").view(B, h, w_, npad)[..., :co].permute(0, 3, 1, 2).float()
            inter[f"stn.ypre{i}"] = _rel_l2(yp, taps[f"stn.ypre{i}"])
            inter[f"stn.yact{i}"] = _rel_l2(ya, taps[f"stn.yact{i}"])
        inter["stn.f1pre"] = _rel_l2(_ws_tensor(env, m, B, "stn.f1pre").view(B, 512).float(), taps["stn.f1pre"])
        inter["stn.f1"] = _rel_l2(_ws_tensor(env, m, B, "stn.f1").view(B, 512).float(), taps["stn.f1"])
        inter["x_tps_maxabs"] = (_ws_tensor(env, m, B, "x_tps").view(B, 3, 16, 64) - taps["x_tps"]).abs().max().item()
    inter["b1"] = _rel_l2(_nchw(_ws_tensor(env, m, B, "b1"), B, 16, 64, 64), taps["b1"])
    for i in range(5):
        for f in ("c1", "a1", "c2", "out"):
            inter[f"srb{i}.{f}"] = _rel_l2(_nchw(_ws_tensor(env, m, B, f"srb{i}.{f}"), B, 16, 64, 64),
                                           taps[f"block{i + 2}.{f}"])
    inter["s7"] = _rel_l2(_nchw(_ws_tensor(env, m, B, "s7"), B, 16, 64, 64), taps["s7"])
    inter["u"] = _rel_l2(_nchw(_ws_tensor(env, m, B, "u"), B, 32, 128, 64), taps["u"])
    inter["opre"] = _rel_l2(_ws_tensor(env, m, B, "opre").view(B, 3, 32, 128), taps["opre"])
    rep["intermediates_rel_l2"] = inter
    rep["sr_rel_l2"] = _rel_l2(sr.detach(), info["sr"])
    rep["sr_maxabs"] = (sr.detach() - info["sr"]).abs().max().item()
    rep["torch_bf16_sr_rel_l2"] = _rel_l2(cal["sr"].float(), info["sr"])
    rep["loss"] = [loss.item(), info["mse"].item()]
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    rep["grad_missing"] = sorted(set(info["grads"]) - set(grads))
    rep["grad_extra"] = sorted(set(grads) - set(info["grads"]))
    if not rep["grad_missing"]:
        rep["grads"] = _grad_report(grads, info["grads"], cal["grads"])
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    rep["grad_norm"] = [gn, info["grad_norm"].item(), cal["grad_norm"].item()]
    REPORT["train_stn" if stn else "train_nostn"] = rep
    _dump()
    assert rep["sr_rel_l2"] < 2.5e-2 and rep["sr_rel_l2"] <= 1.1 * rep["torch_bf16_sr_rel_l2"], \
        (rep["sr_rel_l2"], rep["torch_bf16_sr_rel_l2"])
    assert abs(loss.item() - info["mse"].item()) < 5e-3 * info["mse"].item()
    assert not rep["grad_missing"] and not rep["grad_extra"]
    if stn:
        assert sorted(set(k for k, _ in m.named_parameters()) - set(grads)) == env["golden"]["no_grad_params"]
    rel, cal_rel = rep["grads"]["rel_l2"], rep["grads"]["torch_bf16_rel_l2"]
    trunk = {k: v for k, v in rel.items() if not k.startswith("stn_head.")}
    for k, v in trunk.items():  # within 10 %, or (cancellation-dominated scalars) half the stock-bf16 error
        assert v < 0.1 or v < 0.5 * cal_rel[k], (k, v, cal_rel[k])
    # STN-head gradients pass through 7 batch-statistics BatchNorms over as few as 2B samples and a bilinear
    # resampler: in bf16 they are noise-limited (stock autocast(bf16) is > 100 % off); require half that error
    for k, v in rel.items():
        if k.startswith("stn_head."):
            assert v < 0.5 and v < 0.5 * cal_rel[k], (k, v, cal_rel[k])
    import statistics
    assert statistics.median(rel.values()) <= 1.1 * statistics.median(cal_rel.values())
    assert max(rep["grads"]["zero_grad_abs_over_gmax"].values(), default=0.0) < 1e-2
    assert abs(gn - info["grad_norm"].item()) < 0.03 * info["grad_norm"].item()
    # BatchNorm running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased var)
    msd = m.state_dict()
    new_sd, _ = O.train_step(sd, lr, hr, {}, masks=None, stn=stn)
    for k in msd:
        if "running_" in k and (stn or not k.startswith("stn_head")) and not k.startswith("bn."):
            assert torch.allclose(msd[k], new_sd[k], atol=3e-3, rtol=2e-2), k
    assert int(msd["block2.bn1.num_batches_tracked"]) == 1


@pytest.mark.parametrize("stn", [False, True])
def test_train_forward_backward_conditioned_absolute(env, stn):
    """north_star's 1e-2 contract, asserted ABSOLUTELY in train mode.  The synthetic gain-1 weights of the case above are
    the worst conditioning the network ever sees (MSE 0.34, every BatchNorm input white noise); 50 deterministic oracle
    Adam steps (interfaces/super_resolution.py:69-84, fp32, cudnn deterministic) bring the weights to where training
    actually operates (MSE 0.008) and the engine's train-mode SR is then 7e-3 from the fp32 oracle (stock autocast(bf16):
    8.4e-3 without the STN, 5e-1 with it).  Gradients at the same weights: per-tensor relative L2, absolute bounds."""
    import statistics
    O, synth = env["O"], env["synth"]
    B = 32
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    _strict_fp32()
    det, bench_ = torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    try:
        sd = {k: v.to(DEV) for k, v in env["sd"].items()}
        state = {}
        for _ in range(50):
            sd, _info = O.train_step(sd, lr, hr, state, masks=None, stn=True)
        _, info = O.train_step(sd, lr, hr, {}, masks=None, stn=stn)
    finally:
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = det, bench_
    m = env["TBSRN"](STN=stn).to(DEV)
    m.load_state_dict({k: v for k, v in sd.items() if stn or not (k.startswith("stn_head") or k.startswith("tps"))})
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train()
    sr = m(lr)
    loss = F.mse_loss(sr, hr)
    (loss * 100).backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    rep = {"sr_rel_l2": _rel_l2(sr.detach(), info["sr"]), "sr_maxabs": (sr.detach() - info["sr"]).abs().max().item(),
           "loss": [loss.item(), info["mse"].item()], "grads": _grad_report(grads, info["grads"])}
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    rep["grad_norm"] = [gn, info["grad_norm"].item()]
    rel = rep["grads"]["rel_l2"]
    rep["grad_rel_median"], rep["grad_rel_max"] = statistics.median(rel.values()), max(rel.values())
    REPORT["train_conditioned_stn" if stn else "train_conditioned_nostn"] = rep
    _dump()
    assert rep["sr_rel_l2"] < 1e-2, rep["sr_rel_l2"]                                # the contract
    assert abs(loss.item() - info["mse"].item()) < 5e-3 * info["mse"].item()
    trunk = [v for k, v in rel.items() if not k.startswith("stn_head.")]
    assert statistics.median(trunk) < 5e-2 and max(trunk) < 0.12, (statistics.median(trunk), max(trunk))
    assert abs(gn - info["grad_norm"].item()) < 0.03 * info["grad_norm"].item()


def test_reference_loop_and_fused_trainer_agree(env):
    """the unchanged reference step (torch MSELoss + clip_grad_norm_ + torch Adam on the drop-in module) and the
    fused TBSRNTrainer must walk the same trajectory; both must track the oracle."""
    from fudanocr_b200.trainer import TBSRNTrainer
    O, synth = env["O"], env["synth"]
    B = 4
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    _strict_fp32()
    # STN off: at B = 4 the BatchNorm1d/2-sample-BN prologue amplifies the 1-ulp difference between torch's and the
    # fused MSE gradient into percent-level gradient changes, which says nothing about the two front-ends
    m1, m2 = _model(env, stn=False), _model(env, stn=False)
    opt = torch.optim.Adam(m1.parameters(), lr=1e-4, betas=(0.5, 0.999))
    tr = TBSRNTrainer(m2)
    sd = {k: v.to(DEV) for k, v in env["sd"].items()}
    ost = {}
    p0 = {k: v.detach().clone() for k, v in m1.named_parameters()}
    for it in range(2):
        sr = m1(lr)
        loss = F.mse_loss(sr, hr)
        opt.zero_grad()
        (loss * 100).backward()
        gn1 = torch.nn.utils.clip_grad_norm_(m1.parameters(), 0.25)
        opt.step()
        l2 = tr.step(lr, hr)
        sd, info = O.train_step(sd, lr, hr, ost, masks=None, stn=False)
        torch.cuda.synchronize()
        REPORT[f"step{it}"] = dict(loss_ref_loop=loss.item(), loss_trainer=l2.item(), loss_oracle=info["mse"].item(),
                                   gn_ref_loop=gn1.item(), gn_trainer=tr.grad_norm.item(),
                                   gn_oracle=info["grad_norm"].item())
        if it == 0:
            gmax = max(g.norm().item() for g in info["grads"].values())
            live = {k for k, g in info["grads"].items() if g.norm().item() >= 1e-5 * gmax}
            live_names = live
        if it == 0:  # the two front-ends drive the same kernels: gradients must agree tensor by tensor
            tensors, _ = m2._slots()
            names = [m2._slot_names[i] for i in m2._grad_slots]
            g1 = dict((k, p.grad) for k, p in m1.named_parameters() if p.grad is not None)
            # autograd path grads were clipped in place by clip_grad_norm_: undo for the comparison
            coef = min(1.0, 0.25 / (gn1.item() + 1e-6))
            diffs = {}
            off = 0
            for i, k in zip(m2._grad_slots, names):
                n = tensors[i].numel()
                g2 = tr.flat_g[off:off + n]
                off += (n + 3) // 4 * 4
                d = (g1[k].reshape(-1) / coef - g2).norm().item() / (g2.norm().item() + 1e-20)
                # the two paths differ by ~1 ulp in d_sr (torch's MSE backward vs the fused one); batch-statistics BN at
                # B = 4 amplifies that for cancellation-dominated sums (bias-type gradients).  The kernels themselves are
                # bit-reproducible and workspace-independent, and moving the targets by ONE ulp moves these gradients by
                # 1e-3 .. 2.5e-3 (scripts/ws_independence.py: bf16 rounding flips cascade), so the front-ends agree when the
                # mismatch stays inside half the tensor's own bf16 error against the fp32 oracle.
                e_or = (info["grads"][k].reshape(-1) - g2).norm().item() / (g2.norm().item() + 1e-20)
                if d > (5e-3 if n == 1 else 1e-3) and d > 0.5 * e_or and k in live_names and not k.startswith("stn_head."):
                    diffs[k] = (d, e_or)  # (the B=4 STN prologue amplifies 1-ulp differences of d_sr even more)
            REPORT["frontends_grad_mismatch"] = dict(sorted(diffs.items(), key=lambda kv: -kv[1][0])[:20])
            _dump()
        # step 0: same weights, so the two losses are the same number; step 1: after one Adam step (lr * sign(g) per element)
        # of each front-end - sign flips of numerically-zero gradient elements move the second loss by a few 1e-4 relative
        assert abs(loss.item() - l2.item()) < 1e-6 + (1e-5 if it == 0 else 5e-4) * abs(loss.item())
        assert abs(gn1.item() - tr.grad_norm.item()) < 1e-3 * gn1.item(), REPORT.get("frontends_grad_mismatch")
        if it == 0:
            assert not REPORT["frontends_grad_mismatch"], REPORT["frontends_grad_mismatch"]
        assert abs(l2.item() - info["mse"].item()) < 5e-3 * info["mse"].item()
        if it == 0:
            gmax = max(g.norm().item() for g in info["grads"].values())
            live = {k for k, g in info["grads"].items() if g.norm().item() >= 1e-5 * gmax}
    # Adam's first steps are lr*sign(g) per element, so weights are compared through the update direction on the
    # tensors that carry a real gradient (numerically-zero gradients turn into +-lr noise in torch as well)
    d2 = {k: (v.detach() - p0[k]) for k, v in m2.named_parameters()}
    dref = {k: (sd[k] - env["sd"][k].to(DEV)) for k in d2}
    cos = {k: F.cosine_similarity(d2[k].flatten(), dref[k].flatten(), dim=0).item() for k in d2
           if k in live and dref[k].norm() > 0}
    REPORT["update_cosine_min"] = min(cos.values())
    REPORT["update_cosine_mean"] = sum(cos.values()) / len(cos)
    _dump()
    assert REPORT["update_cosine_mean"] > 0.8, REPORT["update_cosine_mean"]


def test_dropout_on_matches_oracle_with_same_masks(env):
    """dropout live (p = 0.1): the oracle gets the exact masks the kernels regenerate from (seed, layer)"""
    from oracle import dropout_rng as R
    O, synth = env["O"], env["synth"]
    B, seed, p = 2, 4242, 0.1
    _strict_fp32()
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    from fudanocr_b200.trainer import TBSRNTrainer
    m = env["TBSRN"](STN=False).to(DEV)  # the ill-conditioned small-batch STN prologue is covered separately
    m.load_state_dict({k: v for k, v in env["sd"].items() if not (k.startswith("stn_head") or k.startswith("tps"))})
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = p
    m.train()
    tr = TBSRNTrainer(m, lr=0.0)  # lr 0: inspect gradients without moving the weights
    loss = tr.step(lr, hr, seed=seed)
    torch.cuda.synchronize()
    masks = {}
    for i in range(5):
        masks[f"block{i + 2}.feature_enhancer.attn"] = R.attn_keep_mask(B, seed, i, p).to(DEV)
        masks[f"block{i + 2}.feature_enhancer.attn_scale"] = R.attn_keep_scale(p)
        masks[f"block{i + 2}.feature_enhancer.ffn"] = R.ffn_keep_mask(B, seed, i, p).to(DEV)
    sd = {k: v.to(DEV) for k, v in env["sd"].items()}
    _, info = O.train_step(sd, lr, hr, {}, masks=masks, stn=False)
    _, info_nodrop = O.train_step(sd, lr, hr, {}, masks=None, stn=False)
    err = (tr.sr - info["sr"]).abs().max().item()
    sep = (info_nodrop["sr"] - info["sr"]).abs().max().item()
    REPORT["dropout_sr_maxabs"] = err
    REPORT["dropout_effect_maxabs"] = sep
    REPORT["dropout_gn"] = [tr.grad_norm.item(), info["grad_norm"].item()]
    _dump()
    rel = _rel_l2(tr.sr, info["sr"])
    REPORT["dropout_sr_rel_l2"] = rel
    _dump()
    assert rel < 2.5e-2 and sep > 3 * err, (rel, err, sep)
    assert abs(tr.grad_norm.item() - info["grad_norm"].item()) < 0.03 * info["grad_norm"].item()


def test_full_size_properties(env):
    """BASELINE size (B = 256): size-independent properties — eval batches are independent (a 256-crop batch
    equals its two halves, bit for bit), a training step is bit-reproducible (no atomics), the loss falls."""
    from fudanocr_b200.trainer import TBSRNTrainer
    torch.manual_seed(0)
    B = 256
    lr = torch.rand(B, 3, 16, 64, device=DEV)
    hr = F.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False)
    m = _model(env, train=False)
    with torch.no_grad():
        full = m(lr).clone()
        a = m(lr[:128]).clone()
        b = m(lr[128:]).clone()
    assert torch.equal(full, torch.cat([a, b]))
    losses = []
    for rep in range(2):
        m = _model(env, p_drop=0.1)
        tr = TBSRNTrainer(m, lr=1e-3)
        ls = [tr.step(lr, hr, seed=100 + it).item() for it in range(6)]
        losses.append(ls)
    REPORT["b256_losses"] = losses
    REPORT["b256_mem_gib"] = torch.cuda.max_memory_allocated() / 2 ** 30
    _dump()
    assert losses[0] == losses[1], losses
    assert losses[0][-1] < losses[0][0]


def test_graph_replay_is_bit_identical_to_eager(env):
    """the trainer replays the step as CUDA graphs (device-resident dropout seed, static input buffers): five steps with
    dropout ON and changing batches must reproduce the eager launches bit for bit - losses, weights, launch count"""
    from fudanocr_b200.trainer import TBSRNTrainer
    B = 8

    def run(use_graph):
        torch.manual_seed(1234)
        m = _model(env, stn=True)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.1
        tr = TBSRNTrainer(m, use_graph=use_graph)
        g = torch.Generator(device=DEV).manual_seed(5)
        losses = []
        for i in range(5):
            lr = torch.rand(B, 3, 16, 64, device=DEV, generator=g)
            hr = torch.rand(B, 3, 32, 128, device=DEV, generator=g)
            losses.append(tr.step(lr, hr, seed=1000 + i).clone())
        torch.cuda.synchronize()
        return torch.cat(losses), tr.flat_p.clone(), tr.kernel_launches, tr._graphs is not None

    l0, p0, n0, g0 = run(False)
    l1, p1, n1, g1 = run(True)
    assert g1 and not g0
    assert torch.equal(l0, l1), (l0, l1)
    assert torch.equal(p0, p1)
    assert n0 == n1 and n0 > 0
