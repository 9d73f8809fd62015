"""CTC forward-backward kernel (csrc/ctc.cu, drop-in fudanocr_b200.loss.ctc_loss) through the C ABI vs the golden values
recorded from torch.nn.functional.ctc_loss (tests/golden/ctc.npz), the float64 oracle, and torch's own CUDA op on the box.
Floating point: fp32 kernel vs float64 truth - loss 1e-5 relative; gradient 1e-4 of the largest entry (the lattices live in
the log domain at magnitudes ~T log C ~ 100, where one fp32 ulp is ~1e-5: occupancies carry a few 1e-5 relative error, as
torch's own fp32 kernel does; north_star asks 1e-3 relative for fp32)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = {"crnn_b8": (26, 8, 37, 12, 1, True), "full_b4": (26, 4, 37, 10, 2, False), "wide_b3": (40, 3, 97, 19, 3, True),
         "tiny_b5": (3, 5, 5, 2, 4, True)}


def _run(logits, targets, il, tl, reduction, zero_infinity=False):
    from fudanocr_b200.loss.ctc_loss import ctc_loss
    x = torch.from_numpy(logits).to(DEV).requires_grad_(True)
    loss = ctc_loss(x, torch.from_numpy(targets), torch.from_numpy(il), torch.from_numpy(tl), 0, reduction, zero_infinity)
    (loss.sum() if reduction == "none" else loss).backward()
    return loss.detach().cpu().double().numpy(), x.grad.cpu().double().numpy()


def test_ctc_matches_torch_golden_and_oracle():
    from oracle import synth, ctc_oracle as CO
    g = np.load(synth.GOLDEN_DIR / "ctc.npz")
    for name, (T, B, C, S, seed, ragged) in CASES.items():
        logits, targets, il, tl = CO.synth_case(T, B, C, S, seed, ragged)
        chk = float(logits.astype(np.float64).sum()) + float(targets.sum()) + float(il.sum() + tl.sum())
        assert abs(chk - float(g[f"{name}/checksum"])) < 1e-9 * max(abs(chk), 1.0)
        for red in ("mean", "sum", "none"):
            loss, grad = _run(logits, targets, il, tl, red)
            ref_l, ref_g = g[f"{name}/{red}/loss"], g[f"{name}/{red}/grad"]
            assert np.allclose(loss, ref_l, rtol=1e-5, atol=1e-6), (name, red, loss, ref_l)
            assert np.abs(grad - ref_g).max() < 1e-4 * np.abs(ref_g).max(), (name, red, np.abs(grad - ref_g).max())
    # frames past the input length and the padded target tail must not matter
    logits, targets, il, tl = CO.synth_case(26, 8, 37, 12, 1, True)
    l0, g0 = _run(logits, targets, il, tl, "mean")
    logits2, targets2 = logits.copy(), targets.copy()
    for b in range(8):
        logits2[il[b]:, b] = 77.0
        targets2[b, tl[b]:] = 36
    l1, g1 = _run(logits2, targets2, il, tl, "mean")
    assert l0 == l1 and np.array_equal(g0, g1)
    for b in range(8):
        assert np.all(g0[il[b]:, b] == 0)


def test_ctc_edge_cases_infeasible_empty_and_errors():
    from oracle import synth, ctc_oracle as CO
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.loss.ctc_loss import CTCLoss, ctc_loss
    g = np.load(synth.GOLDEN_DIR / "ctc.npz")
    logits, targets, il, tl = CO.synth_case(6, 3, 7, 5, 9, False)
    il[0] = 2
    tl[1] = 0
    assert np.array_equal(np.stack([il, tl]), g["inf_b3/lengths"])
    loss, grad = _run(logits, targets, il, tl, "mean", zero_infinity=True)
    assert np.allclose(loss, g["inf_b3/mean/loss"], rtol=1e-5)
    assert np.abs(grad - g["inf_b3/mean/grad"]).max() < 1e-4 * np.abs(g["inf_b3/mean/grad"]).max() and np.all(grad[:, 0] == 0)
    # without zero_infinity the infeasible sample reports inf, as torch does
    nll, _ = _run(logits, targets, il, tl, "none")
    assert np.isinf(nll[0]) and np.isfinite(nll[1:]).all()
    # 1-D concatenated targets == padded targets; module form == functional form; bf16 logits accepted
    x = torch.from_numpy(logits).to(DEV)
    cat = torch.cat([torch.from_numpy(targets[b, :tl[b]]) for b in range(3)])
    a = ctc_loss(x, cat, il.tolist(), tl.tolist(), zero_infinity=True)
    b_ = CTCLoss(zero_infinity=True)(x, torch.from_numpy(targets), torch.from_numpy(il), torch.from_numpy(tl))
    assert float(a) == float(b_) and abs(float(a) - float(loss)) < 1e-6
    with pytest.raises(L.FocrError):
        ctc_loss(torch.from_numpy(logits), torch.from_numpy(targets), il, tl)
    with pytest.raises(ValueError):
        ctc_loss(x[0], torch.from_numpy(targets), il, tl)
    # C-ABI argument errors come back as codes, not crashes
    assert L.lib.focr_ctc_loss(x.data_ptr(), 6, 3, 7, 0, 5, 0, 0, 0, 1, 0, 1.0, 0, 0, 0, 0, 0, 0) != 0
    assert b"null" in L.lib.focr_last_error()


def test_ctc_full_batch_vs_torch_cuda_and_determinism():
    """CRNN-shaped batch at the eval pipeline's size (B = 1024, T = 26, C = 37): against torch's own CUDA ctc_loss on the
    same device (library call used as a checker only), bit-identical across two runs, sum of softmax-minus-occupancy = 0"""
    from oracle import ctc_oracle as CO
    import torch.nn.functional as F
    logits, targets, il, tl = CO.synth_case(26, 1024, 37, 12, 21, True)
    l0, g0 = _run(logits, targets, il, tl, "mean")
    l1, g1 = _run(logits, targets, il, tl, "mean")
    assert l0 == l1 and np.array_equal(g0, g1)
    x = torch.from_numpy(logits).to(DEV).double().requires_grad_(True)
    ref = F.ctc_loss(F.log_softmax(x, 2), torch.from_numpy(targets).to(DEV), torch.from_numpy(il).to(DEV),
                     torch.from_numpy(tl).to(DEV), blank=0, reduction="mean")
    ref.backward()
    assert abs(float(ref.detach()) - float(l0)) < 1e-5 * float(ref.detach())
    gref = x.grad.cpu().numpy()
    assert np.abs(gref - g0).max() < 1e-4 * np.abs(gref).max()
    # per frame the gradient sums to zero over classes (softmax sums to 1, occupancies sum to 1)
    assert np.abs(g0.sum(2)).max() < 1e-6
