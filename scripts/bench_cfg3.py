#!/usr/bin/env python
"""BASELINE configs[2]: text-gestalt TSRN + stroke-focus loss (lambda 50), 32 crops per GPU (global 256 on 8 GPUs),
one process per GPU under torchrun.  A step = TSRN forward, StrokeFocusLoss value+gradient (frozen recogniser: HR forward,
SR forward, SR input-gradient chain), TSRN backward, gradient all-reduce (N > 1), clip + Adam.  Not the headline bench
(bench.py); prints one JSON line with the per-family device-time breakdown."""
import argparse
import json
import os
import random
import string
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.model.tsrn import TSRN
    from fudanocr_b200.trainer import TBSRNTrainer
    from fudanocr_b200.loss.stroke_focus_loss import StrokeFocusLoss
    from fudanocr_b200.loss.transformer_english_decomposition import Transformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    B, K = args.batch, args.steps
    model = TSRN(STN=True).to(dev)
    model.train()
    # synthetic stand-ins for the git-ignored assets: random-init recogniser, 1-4 strokes per character
    chars = string.digits + string.ascii_lowercase + string.ascii_uppercase
    dic = {c: "".join(str(1 + (i * 3 + 5 * k) % 9) for k in range(1 + (i * 7) % 4)) for i, c in enumerate(chars)}
    crit = StrokeFocusLoss(types.SimpleNamespace(text_focus=True, stroke_lambda=50), decomposition=dic,
                           transformer_state_dict=Transformer("tg").state_dict()).to(dev)
    trainer = TBSRNTrainer(model, criterion=crit)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    lr = torch.rand(B, 3, 16, 64, device=dev, generator=g)
    hr = torch.rand(B, 3, 32, 128, device=dev, generator=g)
    rnd = random.Random(1234 + rank)
    labels = ["".join(rnd.choice(string.digits + string.ascii_lowercase) for _ in range(rnd.randint(1, 12))) for _ in range(B)]
    for it in range(max(args.warmup, 3)):
        trainer.step(lr, hr, seed=it, labels=labels)
    torch.cuda.synchronize()
    L.prof_enable(1, b"")
    trainer.step(lr, hr, seed=99, labels=labels)
    breakdown = L.prof_collect()
    L.prof_enable(0, b"")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.lib.focr_launch_count()
    e0.record()
    for it in range(K):
        trainer.step(lr, hr, seed=100 + it, labels=labels)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        T = max(len("".join(dic[c] for c in s)) + 1 for s in labels)
        flop = B * (3 * 26.7e9 + 5.4e9)   # recogniser 2 fwd + 1 dgrad, TSRN fwd+bwd
        print(json.dumps({
            "metric": "tsrn_strokefocus_train_images_per_sec", "value": world * B * K / (ms * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": K, "ms_per_step": ms / K, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "TSRN (STN) + StrokeFocusLoss(lambda 50) train step, batch %d per GPU "
                                   "(BASELINE configs[2])" % B, "decoder_T": T},
            "algorithmic_tflop_per_step": flop / 1e12, "achieved_tflops": flop / (ms / K * 1e-3) / 1e12,
            "launches_per_step": (L.lib.focr_launch_count() - n0) / K,
            "losses": [float(x) for x in trainer.losses.cpu()],
            "breakdown_ms_per_step": {k: round(v[1], 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1][1])},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
