"""Does a briefly TRAINED recogniser make the focus-loss gradient follow fp32 under bf16?  Trains the fp32 oracle recogniser
(text-gestalt/loss/transformer_english_decomposition.py restated in oracle/focus_oracle.py, eval-mode BatchNorm) on a fixed
synthetic batch with Adam(1e-4) for N steps, then prints the relative L2 error of d(loss)/d(sr) and of the attention maps for the
engine and for stock autocast(bf16) against the fp32 oracle.  Usage: python scripts/focus_conditioning.py [steps ...]"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import focus_oracle as FO, sld_oracle as SO, synth

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
marks = [int(a) for a in sys.argv[1:]] or [0, 40, 120]
g = torch.load(synth.GOLDEN_DIR / "focus_b2.pt", weights_only=False)
sd = {k: v.to(DEV) for k, v in FO.synth_recogniser_state_dict(synth.load_spec("focus"), g["bn_stats"]).items()}
dic = FO.synth_decomposition()
chars = sorted(dic)
import numpy as np
rs = np.random.RandomState(5)
Bt = 32
_, hr_t = synth.synth_images(Bt, seed=99)
hr_t = hr_t.to(DEV)
labels_t = ["".join(rs.choice(chars, size=int(rs.randint(1, 7)))) for _ in range(Bt)]
length_t, inp_t, gt_t = [t.to(DEV) for t in FO.label_stroke_encoder(labels_t, dic)]


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def measure(tag, sd):
    from fudanocr_b200.loss.stroke_focus_loss import StrokeFocusLoss
    crit = StrokeFocusLoss(types.SimpleNamespace(text_focus=True, stroke_lambda=50), decomposition=dic,
                           transformer_state_dict={k: v.detach().cpu() for k, v in sd.items()}).to(DEV)
    lr, hr = synth.synth_images(6, seed=23)
    sr = F.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=False).clamp(0, 1).to(DEV)
    hr = hr.to(DEV)
    labels = ["a", "focus", "B200", "stroke9", "xyzzy", "Q"]

    def oracle(autocast):
        s = sr.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss, mse, att, info = FO.stroke_focus_loss(sd, s, hr, labels, dic, 50.0, nm=FO.Numerics(fold=True))
        (loss.float() * 100).backward()
        return att.detach().float(), info["map_sr"].detach().float(), s.grad
    att_o, ms_o, g_o = oracle(False)
    att_b, ms_b, g_b = oracle(True)
    _, text_input, _ = crit.label_stroke_encoder(labels, DEV)
    losses, d_sr, mh, ms = crit._run(sr, hr, text_input, 50.0, 100.0, maps=True)
    torch.cuda.synchronize()
    mse_g = 200.0 * (sr - hr) / sr.numel()
    print(f"{tag}: d_sr engine {rel(d_sr, g_o):.3e} stock {rel(g_b, g_o):.3e} | attention part only: engine {rel(d_sr - mse_g, g_o - mse_g):.3e} "
          f"stock {rel(g_b - mse_g, g_o - mse_g):.3e} | map_sr engine {rel(ms, ms_o):.3e} stock {rel(ms_b, ms_o):.3e} | att {float(losses[2]):.5f} "
          f"fp32 {float(att_o):.5f} stock {float(att_b):.5f}", flush=True)


params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
          if v.is_floating_point() and "running" not in k and not k.startswith("pe.")}
full = dict(sd)
full.update(params)
opt = torch.optim.Adam(list(params.values()), lr=float(os.environ.get("FOCUS_LR", "1e-4")))
step = 0
for m in marks:
    while step < m:
        opt.zero_grad(set_to_none=True)
        logits, _, _ = FO.transformer_forward(full, FO.to_gray_tensor(hr_t), length_t, inp_t)
        loss = F.cross_entropy(logits, gt_t)
        loss.backward()
        opt.step()
        step += 1
        if step % 20 == 0:
            print(f"  step {step}: recogniser CE {float(loss):.4f}", flush=True)
    measure(f"after {step} recogniser steps", {k: v.detach() for k, v in full.items()})
