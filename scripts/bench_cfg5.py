#!/usr/bin/env python
"""BASELINE configs[4]: the joint TBSRN -> CRNN + CTC evaluation pipeline (scene-text-telescope TextSR.eval), 128 crops per GPU
(global 1024 on 8 GPUs; no collective: ranks own disjoint shards).  A step = SR forward (eval), PSNR + SSIM, bicubic + gray,
CRNN forward, greedy CTC decode, strings on the host.  Prints one JSON line."""
import argparse, json, os, random, string, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.model.tbsrn import TBSRN
    from fudanocr_b200.model.crnn import CRNN
    from fudanocr_b200.interfaces.recognition import evaluate_batch
    torch.manual_seed(1234)
    dev = "cuda"
    B = args.batch
    model = TBSRN().to(dev).eval()
    crnn = CRNN(32, 1, 37, 256).to(dev).eval()
    lr, hr = torch.rand(B, 3, 16, 64, device=dev), torch.rand(B, 3, 32, 128, device=dev)
    rnd = random.Random(0)
    labels = ["".join(rnd.choice(string.ascii_lowercase) for _ in range(rnd.randint(1, 12))) for _ in range(B)]
    for _ in range(3):
        evaluate_batch(model, crnn, lr, hr, labels)
    torch.cuda.synchronize()
    n0 = L.lib.focr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = evaluate_batch(model, crnn, lr, hr, labels)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"metric": "tbsrn_crnn_eval_images_per_sec", "value": B / (ms * 1e-3), "unit": "images/s", "n_gpus": 1,
                      "ms_per_step": ms, "config": {"workload": "TBSRN eval -> PSNR/SSIM -> bicubic+gray -> CRNN -> greedy CTC decode, "
                                                                "batch %d per GPU (BASELINE configs[4]), strings decoded on the host" % B},
                      "algorithmic_gflop_per_image": 6.5, "achieved_tflops": B * 6.5e9 / (ms * 1e-3) / 1e12,
                      "launches_per_step": (L.lib.focr_launch_count() - n0) / args.steps,
                      "psnr": float(out["psnr"]), "ssim": float(out["ssim"])}))


if __name__ == "__main__":
    main()
