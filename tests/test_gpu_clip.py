"""CCR-CLIP contrastive head (csrc/clip_loss.cu) against the oracle restatement of image-ids-CTR/CCR-CLIP/main.py:98-110 +
model.py:209-222, through the C-ABI and the autograd wrapper.  fp32 kernel vs fp64 oracle: loss 1e-5, gradients 1e-4 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    """relative L2; a reference that is exactly zero (B = 1: the soft-max of one logit has no gradient) is judged absolutely -
    the kernel's fp32 exp / log leave ~1e-7 there"""
    if float(b.double().norm()) == 0.0:
        return float(a.double().abs().max()) * 1e1
    return float((a.double() - b.double()).norm() / b.double().norm())


def _case(B, D, seed, n_chars):
    rs = np.random.RandomState(seed)
    chars = [chr(0x4E00 + i) for i in range(n_chars)]
    labels = [chars[i] for i in rs.randint(0, n_chars, size=B)]          # single-character labels with repeats, as in the font set
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, D, generator=g) * 3.0
    txt = (0.5 * img + torch.randn(B, D, generator=g)) * 0.7               # correlated towers: a non-trivial soft-max
    return labels, img, txt


@pytest.mark.parametrize("B,D,n_chars", [(128, 2048, 90), (37, 100, 1000), (256, 512, 40), (1, 64, 3)])
def test_clip_contrastive_loss_and_gradients(B, D, n_chars):
    from fudanocr_b200.loss.clip_contrastive import clip_contrastive_loss, ground_truth_from_labels
    from oracle import clip_oracle as CO
    labels, img, txt = _case(B, D, B + D, n_chars)
    gt = ground_truth_from_labels(labels)
    assert torch.equal(gt, CO.ground_truth(labels))
    ls = torch.tensor(np.log(1 / 0.07), dtype=torch.float32)
    ri, rt, rl = (t.double().clone().requires_grad_(True) for t in (img, txt, ls))
    ref, _ = CO.contrastive_loss(ri, rt, rl, gt)
    ref.backward()
    ei, et, el = (t.to(DEV).clone().requires_grad_(True) for t in (img, txt, ls))
    loss = clip_contrastive_loss(ei, et, el, gt.to(DEV))
    (loss * 3.0).backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref)) + 1e-6, (float(loss), float(ref))
    assert _rel(ei.grad.cpu() / 3.0, ri.grad) < 1e-4
    assert _rel(et.grad.cpu() / 3.0, rt.grad) < 1e-4
    assert abs(float(el.grad) / 3.0 - float(rl.grad)) < 1e-4 * abs(float(rl.grad)) + 2e-5   # (fp32 exp / log noise x exp(logit_scale) = 14)
    # bf16 tower outputs are accepted (cast to fp32 inside), gradients come back in the input dtype
    bi, bt = (t.to(DEV).to(torch.bfloat16).requires_grad_(True) for t in (img, txt))
    lb = clip_contrastive_loss(bi, bt, ls.to(DEV), gt.to(DEV))
    lb.backward()
    assert bi.grad.dtype == torch.bfloat16 and torch.isfinite(lb)
    # value-only call, twice: bit-identical
    with torch.no_grad():
        a = clip_contrastive_loss(ei, et, el, gt.to(DEV))
        b = clip_contrastive_loss(ei, et, el, gt.to(DEV))
    assert torch.equal(a, b) and torch.equal(a, loss.detach())


def test_clip_contrastive_argument_errors():
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.loss.clip_contrastive import clip_contrastive_loss
    x = torch.randn(4, 8, device=DEV)
    ls = torch.zeros((), device=DEV)
    with pytest.raises(ValueError):
        clip_contrastive_loss(x, x[:3], ls, torch.arange(4, device=DEV))
    with pytest.raises(ValueError):
        clip_contrastive_loss(x, x, ls, torch.arange(5, device=DEV))
    with pytest.raises(L.FocrError):
        clip_contrastive_loss(x.cpu(), x.cpu(), ls.cpu(), torch.arange(4))
    L.set_status_checks(True)
    try:
        with pytest.raises(IndexError):
            clip_contrastive_loss(x, x, ls, torch.tensor([0, 1, 2, 9], device=DEV))
    finally:
        L.set_status_checks(False)
