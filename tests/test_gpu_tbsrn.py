"""Whole-network parity of the focr TBSRN engine (through the C-ABI / the drop-in nn.Module) against the
oracle restatement and the golden vectors recorded from the real reference modules.  GPU only.

Tolerances (north_star): SR pixels within 1e-2 (bf16 compute) of the fp32 reference — SR lives in [-1,1], so
the bound is absolute; gradients are compared per tensor in relative L2."""
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
    golden = torch.load(synth.GOLDEN_DIR / "tbsrn_b2.pt", weights_only=False)
    return dict(synth=synth, O=O, L=L, TBSRN=TBSRN, sd=sd, golden=golden)


def _model(env, p_drop=0.0, train=True):
    m = env["TBSRN"]().to(DEV)
    m.load_state_dict(env["sd"])
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


def test_eval_forward_vs_golden(env):
    m = _model(env, train=False)
    lr, _ = env["synth"].synth_images(2)
    with torch.no_grad():
        sr = m(lr.to(DEV))
    err = (sr.cpu() - env["golden"]["eval_sr"]).abs().max().item()
    REPORT["eval_sr_maxabs"] = err
    _dump()
    assert sr.shape == (2, 3, 32, 128) and err < 1e-2, err


def test_train_forward_backward_vs_oracle(env):
    O, synth = env["O"], env["synth"]
    B = 2
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    m = _model(env)
    sr = m(lr)
    loss = F.mse_loss(sr, hr)
    (loss * 100).backward()
    torch.cuda.synchronize()
    # oracle on the GPU in strict fp32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = {k: v.to(DEV) for k, v in env["sd"].items()}
    taps = {}
    new_sd, info = O.train_step(sd, lr, hr, {}, masks=None, taps=taps)
    # localisation: intermediates
    inter = {}
    inter["x_tps"] = (_ws_tensor(env, m, B, "x_tps").view(B, 3, 16, 64) - taps["x_tps"]).abs().max().item()
    inter["ctrl"] = (_ws_tensor(env, m, B, "ctrl").view(-1, 64)[:B, :40] - taps["ctrl"].reshape(B, 40)).abs().max().item()
    inter["b1"] = _rel_l2(_nchw(_ws_tensor(env, m, B, "b1"), B, 16, 64, 64), taps["b1"])
    for i in range(5):
        for f in ("c1", "a1", "c2", "out"):
            inter[f"srb{i}.{f}"] = _rel_l2(_nchw(_ws_tensor(env, m, B, f"srb{i}.{f}"), B, 16, 64, 64),
                                           taps[f"block{i + 2}.{f}"])
    inter["s7"] = _rel_l2(_nchw(_ws_tensor(env, m, B, "s7"), B, 16, 64, 64), taps["s7"])
    inter["u"] = _rel_l2(_nchw(_ws_tensor(env, m, B, "u"), B, 32, 128, 64), taps["u"])
    inter["opre"] = _rel_l2(_ws_tensor(env, m, B, "opre").view(B, 3, 32, 128), taps["opre"])
    REPORT["intermediates_rel_l2"] = inter
    sr_err = (sr.detach() - info["sr"]).abs().max().item()
    REPORT["train_sr_maxabs_vs_oracle"] = sr_err
    REPORT["train_sr_maxabs_vs_golden"] = (sr.detach().cpu() - env["golden"]["train_sr"]).abs().max().item()
    REPORT["loss"] = [loss.item(), info["mse"].item()]
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    gerr = {k: _rel_l2(grads[k], g) for k, g in info["grads"].items() if k in grads}
    REPORT["grad_rel_l2"] = gerr
    REPORT["grad_missing"] = sorted(set(info["grads"]) - set(grads))
    REPORT["grad_extra"] = sorted(set(grads) - set(info["grads"]))
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    REPORT["grad_norm"] = [gn, info["grad_norm"].item()]
    _dump()
    assert sr_err < 1e-2, sr_err
    assert abs(loss.item() - info["mse"].item()) < 2e-3 * info["mse"].item()
    assert not REPORT["grad_missing"] and not REPORT["grad_extra"]
    assert sorted(set(k for k, _ in m.named_parameters()) - set(grads)) == env["golden"]["no_grad_params"]
    worst = max(gerr.items(), key=lambda kv: kv[1])
    assert worst[1] < 0.1, worst
    assert abs(gn - info["grad_norm"].item()) < 0.03 * info["grad_norm"].item()
    # BatchNorm running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased var)
    msd = m.state_dict()
    for k, v in env["golden"]["new_running"].items():
        assert torch.allclose(msd[k].cpu(), v, atol=3e-3, rtol=2e-2), k
    assert int(msd["block2.bn1.num_batches_tracked"]) == 1


def test_reference_loop_and_fused_trainer_agree(env):
    """the unchanged reference step (torch MSELoss + clip_grad_norm_ + torch Adam on the drop-in module) and the
    fused TBSRNTrainer must walk the same trajectory; both must track the oracle."""
    from fudanocr_b200.trainer import TBSRNTrainer
    O, synth = env["O"], env["synth"]
    B = 2
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    m1, m2 = _model(env), _model(env)
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
        sd, info = O.train_step(sd, lr, hr, ost, masks=None)
        torch.cuda.synchronize()
        REPORT[f"step{it}"] = dict(loss_ref_loop=loss.item(), loss_trainer=l2.item(), loss_oracle=info["mse"].item(),
                                   gn_ref_loop=gn1.item(), gn_trainer=tr.grad_norm.item(),
                                   gn_oracle=info["grad_norm"].item())
        assert abs(loss.item() - l2.item()) < 1e-6 + 1e-4 * abs(loss.item())
        assert abs(gn1.item() - tr.grad_norm.item()) < 1e-3 * gn1.item()
        assert abs(l2.item() - info["mse"].item()) < 3e-3 * info["mse"].item()
    d1 = {k: (v.detach() - p0[k]) for k, v in m1.named_parameters()}
    d2 = {k: (v.detach() - p0[k]) for k, v in m2.named_parameters()}
    dref = {k: (sd[k] - env["sd"][k].to(DEV)) for k in d1}
    cos = {}
    for k in d1:
        if dref[k].norm() == 0:
            assert d1[k].norm() == 0 and d2[k].norm() == 0, k
            continue
        assert torch.allclose(d1[k], d2[k], atol=2e-6), k
        cos[k] = F.cosine_similarity(d2[k].flatten(), dref[k].flatten(), dim=0).item()
    REPORT["update_cosine_min"] = min(cos.values())
    REPORT["update_cosine_mean"] = sum(cos.values()) / len(cos)
    _dump()
    assert REPORT["update_cosine_mean"] > 0.97, REPORT["update_cosine_mean"]


def test_dropout_on_matches_oracle_with_same_masks(env):
    """dropout live (p = 0.1): the oracle gets the exact masks the kernels regenerate from (seed, layer)"""
    from oracle import dropout_rng as R
    O, synth = env["O"], env["synth"]
    B, seed, p = 2, 4242, 0.1
    lr, hr = synth.synth_images(B)
    lr, hr = lr.to(DEV), hr.to(DEV)
    from fudanocr_b200.trainer import TBSRNTrainer
    m = _model(env, p_drop=p)
    tr = TBSRNTrainer(m, lr=0.0)  # lr 0: inspect gradients without moving the weights
    loss = tr.step(lr, hr, seed=seed)
    torch.cuda.synchronize()
    masks = {}
    for i in range(5):
        masks[f"block{i + 2}.feature_enhancer.attn"] = R.attn_keep_mask(B, seed, i, p).to(DEV)
        masks[f"block{i + 2}.feature_enhancer.ffn"] = R.ffn_keep_mask(B, seed, i, p).to(DEV)
    sd = {k: v.to(DEV) for k, v in env["sd"].items()}
    _, info = O.train_step(sd, lr, hr, {}, masks=masks)
    _, info_nodrop = O.train_step(sd, lr, hr, {}, masks=None)
    err = (tr.sr - info["sr"]).abs().max().item()
    sep = (info_nodrop["sr"] - info["sr"]).abs().max().item()
    REPORT["dropout_sr_maxabs"] = err
    REPORT["dropout_effect_maxabs"] = sep
    REPORT["dropout_gn"] = [tr.grad_norm.item(), info["grad_norm"].item()]
    _dump()
    assert err < 1e-2 and sep > 3 * err, (err, sep)
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
