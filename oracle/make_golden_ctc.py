"""Golden CTC values recorded from torch.nn.functional.ctc_loss + autograd (the third-party call behind the north star's
"CTC forward-backward"; the reference has no CTC-loss call site - SURVEY.md §0 D2), and the check that
oracle/ctc_oracle.py reproduces them.  CRNN-shaped case (T=26, C=37: model/crnn/crnn.py:78-80 with CRNN(32,1,37,256),
interfaces/base.py:310) plus ragged / repeated / empty / infeasible targets."""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ctc_oracle as CO, synth  # noqa: E402

CASES = {  # name: (T, B, C, S_max, seed, ragged)
    "crnn_b8": (26, 8, 37, 12, 1, True),
    "full_b4": (26, 4, 37, 10, 2, False),
    "wide_b3": (40, 3, 97, 19, 3, True),
    "tiny_b5": (3, 5, 5, 2, 4, True),
}


def torch_ref(logits, targets, il, tl, reduction, zero_infinity):
    x = torch.from_numpy(logits).double().requires_grad_(True)
    loss = F.ctc_loss(F.log_softmax(x, 2), torch.from_numpy(targets), torch.from_numpy(il), torch.from_numpy(tl), blank=0,
                      reduction=reduction, zero_infinity=zero_infinity)
    (loss.sum() if reduction == "none" else loss).backward()
    return loss.detach().numpy(), x.grad.numpy()


def main():
    out = {}
    for name, (T, B, C, S, seed, ragged) in CASES.items():
        logits, targets, il, tl = CO.synth_case(T, B, C, S, seed, ragged)
        for red in ("mean", "sum", "none"):
            loss, grad = torch_ref(logits, targets, il, tl, red, False)
            o_loss, o_nll, o_grad = CO.ctc_loss(logits, targets, il, tl, 0, red)
            assert np.allclose(o_loss, loss, rtol=1e-10, atol=1e-12), (name, red, o_loss, loss)
            assert np.allclose(o_grad, grad, rtol=1e-8, atol=1e-12), (name, red, np.abs(o_grad - grad).max())
            out[f"{name}/{red}/loss"] = np.asarray(loss, np.float64)
            out[f"{name}/{red}/grad"] = grad.astype(np.float64)
        out[f"{name}/checksum"] = np.asarray(float(logits.astype(np.float64).sum()) + float(targets.sum()) + float(il.sum() + tl.sum()))
    # infeasible sample (target longer than the input) with zero_infinity, and an empty target
    logits, targets, il, tl = CO.synth_case(6, 3, 7, 5, 9, False)
    il[0] = 2
    tl[1] = 0
    loss, grad = torch_ref(logits, targets, il, tl, "mean", True)
    o_loss, o_nll, o_grad = CO.ctc_loss(logits, targets, il, tl, 0, "mean", zero_infinity=True)
    assert np.allclose(o_loss, loss, rtol=1e-10) and np.allclose(o_grad, grad, rtol=1e-8, atol=1e-12)
    assert o_nll[0] == 0.0 and np.all(o_grad[:, 0] == 0)
    out["inf_b3/lengths"] = np.stack([il, tl])
    out["inf_b3/mean/loss"] = np.asarray(loss, np.float64)
    out["inf_b3/mean/grad"] = grad
    gd = synth.GOLDEN_DIR
    np.savez_compressed(gd / "ctc.npz", **out)
    h = hashlib.sha256((gd / "ctc.npz").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "ctc.npz" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  ctc.npz"]) + "\n")
    print("ctc golden:", len(out), "arrays; oracle == torch", torch.__version__)


if __name__ == "__main__":
    main()
