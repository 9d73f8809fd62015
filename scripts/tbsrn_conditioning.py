"""How far is the engine's train-mode SR from the fp32 oracle as the weights get trained?  Prints relative-L2 of SR (engine and
stock autocast(bf16) of the oracle) at the synthetic weights and after N oracle Adam steps (the step body of
interfaces/super_resolution.py:69-84).  Usage: python scripts/tbsrn_conditioning.py [B] [lr]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fudanocr_b200.model.tbsrn import TBSRN
from oracle import synth, tbsrn_oracle as O

dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lr_adam = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
sd = {k: v.to(dev) for k, v in synth.synth_state_dict(synth.load_spec("tbsrn"), 1234, O.tps_buffers()).items()}
lr, hr = synth.synth_images(B)
lr, hr = lr.to(dev), hr.to(dev)


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def measure(tag, sd):
    for stn in (False, True):
        m = TBSRN(STN=stn).to(dev)
        m.load_state_dict({k: v for k, v in sd.items() if stn or not (k.startswith("stn_head") or k.startswith("tps"))})
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.train()
        with torch.no_grad():
            sr = m(lr)
            ref = O.tbsrn_forward(sd, lr, training=True, stn=stn)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                amp = O.tbsrn_forward(sd, lr, training=True, stn=stn)
        print(f"{tag} stn={stn}: engine {rel(sr, ref):.3e}  stock autocast {rel(amp, ref):.3e}  max|d| {float((sr - ref).abs().max()):.3e} "
              f"mse {float(torch.nn.functional.mse_loss(ref, hr)):.5f}", flush=True)


measure("synthetic weights", sd)
state = {}
step = 0
for target in (50, 200, 600):
    while step < target:
        sd, info = O.train_step(sd, lr, hr, state, masks=None, stn=True) if lr_adam == 1e-4 else O.train_step(sd, lr, hr, state, masks=None, stn=True)
        step += 1
    measure(f"after {step} oracle steps", sd)
