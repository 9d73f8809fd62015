#!/usr/bin/env python
"""BASELINE configs[3]: stroke-level-decomposition transformer recogniser train step, one process per GPU
under torchrun.  A step = encoder (40 convs, train-mode BN) + decoder forward, packed cross entropy, backward, gradient
all-reduce (N > 1), Adadelta.  The reference trains on 32x32 crops (config.py:14, batch 32); BASELINE's 32x320 / batch 512
variant is not a reference configuration and its 16x160 maps do not tile into the implicit GEMM's 128-pixel
boxes yet, so this script reports 32x32 at 64 crops per GPU by default.  Not the headline bench (bench.py); prints one JSON
line with the per-family device-time breakdown."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_baseline(kind: str, batch: int = 8, steps: int = 2, width: int = 32):
    """the reference's own CPU path for this step, timed on the host cores: the oracle restatement (oracle/sld_oracle.py /
    ids_oracle.py, pinned to the unmodified reference modules) - forward, loss, backward, Adadelta - torch CPU fp32, all threads.
    The one place outside tests/ where oracle/ may run (bench `cpu_baseline` leg); a reported baseline, not a target."""
    import time
    import torch
    from oracle import ids_oracle as IO, sld_oracle as SO, synth
    torch.set_num_threads(os.cpu_count())
    ids = kind == "ids"
    sd = synth.synth_state_dict(synth.load_spec("ids" if ids else "sld"), 1234)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}
    full = dict(sd)
    full.update(params)
    if ids:
        image, labels = IO.synth_batch(batch)
        length, text_input, text_gt = IO.converter(labels)
        feats = IO.synth_text_features()
    else:
        image, strings = SO.synth_batch(batch)
        if width != 32:   # 32 x width crops: width / 32 synthetic squares side by side
            image = torch.cat([image] + [SO.synth_batch(batch, seed=1234 + 17 * i)[0] for i in range(1, width // 32)], dim=3)
        length, text_input, text_gt = SO.converter_stroke(strings)

    def step():
        for v in params.values():
            v.grad = None
        stats = {}
        if ids:
            loss = IO.loss_fn(full, image, length, text_input, text_gt, feats, stats)[0]
        else:
            loss = SO.loss_fn(full, image, length, text_input, text_gt, None, stats)[0]
        loss.backward()
        with torch.no_grad():
            for k, v in params.items():
                if v.grad is None:
                    continue
                p2, sq, acc = SO.adadelta_update(v, v.grad, *state[k], wd=1e-4 if ids else 0.0)
                v.copy_(p2)
                state[k] = (sq, acc)
            for k, v in stats.items():
                full[k] = v
        return float(loss)
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch / dt, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle {kind} train step (torch CPU fp32, {os.cpu_count()} threads), batch {batch}, {steps} timed steps "
                      f"({dt:.2f} s/step)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--width", type=int, default=32)
    ap.add_argument("--cpu-baseline", action="store_true", help="also time the oracle's CPU step on rank 0 (adds ~20-60 s)")
    ap.add_argument("--model", choices=("sld", "ids"), default="sld",
                    help="ids: image-ids-CTR recogniser (SURVEY A22), 32x256 crops, CLIP-feature similarity loss, weight decay 1e-4")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.model.transformer import Transformer
    from fudanocr_b200.trainer_sld import SLDTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    B, K = args.batch, args.steps
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    rs = np.random.RandomState(1234 + rank)
    ids = args.model == "ids"
    if ids:
        from fudanocr_b200.model.ids_transformer import Transformer as IDSTransformer, N_CLASS
        from fudanocr_b200.trainer_sld import IDSTrainer
        args.width = 256 if args.width == 32 else args.width
        model = IDSTransformer().to(dev).train()
        feats = torch.randn(N_CLASS, 2048, device=dev, generator=g) * 0.3
        trainer = IDSTrainer(model, feats)
        lens, hi, last = rs.randint(2, 26, size=B), N_CLASS - 1, N_CLASS - 1   # 1..25 characters + END
    else:
        model = Transformer("stroke").to(dev).train()
        trainer = SLDTrainer(model)
        lens, hi, last = rs.randint(2, 31, size=B), 6, 6                   # stroke strings of 2..30 symbols incl. '$'
    image = torch.rand(B, 3, 32, args.width, device=dev, generator=g) * 2 - 1
    T = int(lens.max())
    text_input = torch.zeros(B, T, dtype=torch.long)
    gt = []
    for b, n in enumerate(lens):
        s = rs.randint(1, hi, size=n)
        s[-1] = last
        text_input[b, 1:n] = torch.from_numpy(s[:n - 1])
        gt.extend(s.tolist())
    length = torch.from_numpy(lens.astype(np.int64)).to(dev)
    text_input, text_gt = text_input.to(dev), torch.tensor(gt, dtype=torch.long, device=dev)
    for _ in range(max(args.warmup, 3)):
        loss = trainer.step(image, length, text_input, text_gt)
    torch.cuda.synchronize()
    L.prof_enable(1, b"")
    trainer.step(image, length, text_input, text_gt)
    breakdown = L.prof_collect()
    L.prof_enable(0, b"")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = trainer.kernel_launches
    e0.record()
    for _ in range(K):
        loss = trainer.step(image, length, text_input, text_gt)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ranks_identical = None
    if world > 1:   # every rank must have taken the identical step: compare a checksum of the flat parameter buffer
        chk = trainer.flat_p.double().abs().sum().reshape(1)
        got = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(got, chk)
        ranks_identical = all(bool(g == got[0]) for g in got)
    if rank == 0:
        # forward 31.6 GFLOP/img (SLD, 32x32), 18.2 GFLOP/img (IDS, 32x256)
        flop = B * 3 * (18.2e9 * args.width / 256.0 if ids else 31.6e9 * args.width / 32.0)
        top = sorted(((k, round(v[1], 3)) for k, v in breakdown.items()), key=lambda kv: -kv[1])[:12]
        print(json.dumps({"metric": f"{args.model}_train_images_per_sec", "value": world * B * K / (ms / 1e3), "unit": "images/s",
                          "n_gpus": world, "steps": K, "ms_per_step": ms / K, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": (f"IDS Transformer train step, 32x{args.width} crops, batch {B} per GPU, similarity CE + 0.001 * "
                                                  "distance term, Adadelta(lr 1, rho 0.9, wd 1e-4), dropout 0.1") if ids else
                                                 (f"SLD Transformer('stroke') train step, 32x{args.width} crops, batch {B} per GPU, "
                                                  "CE + Adadelta(lr 1, rho 0.9), dropout 0.1"), "T": T},
                          "tflops_required": flop / (ms / K / 1e3) / 1e12, "launches_per_step": (trainer.kernel_launches - n0) / K,
                          "final_loss": float(loss), "ranks_identical": ranks_identical, "breakdown_ms_per_step": dict(top),
                          "cpu_baseline": cpu_baseline(args.model) if args.cpu_baseline else None}))
    if world > 1:   # NCCL work captured in the step graph: leave without destroy_process_group (it can hang on the graph's references)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
