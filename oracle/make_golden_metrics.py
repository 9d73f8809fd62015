"""Golden PSNR / SSIM values from the UNMODIFIED reference module scene-text-telescope/utils/ssim_psnr.py (IPython stubbed)."""
import hashlib
import os
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("FOCR_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "scene-text-telescope"))
from oracle import synth, metrics_oracle as MO  # noqa: E402


def main():
    ip = types.ModuleType("IPython")
    ip.embed = lambda *a, **k: None
    sys.modules.setdefault("IPython", ip)
    from utils import ssim_psnr as ref
    out = {}
    for B, seed in ((5, 3), (2, 9)):
        sr4, hr4 = MO.synth_pair(B, seed)
        psnr, ssim_avg = ref.calculate_psnr(sr4, hr4), ref.SSIM()(sr4, hr4)
        ssim_img = ref.SSIM(size_average=False)(sr4, hr4)
        assert torch.allclose(MO.calculate_psnr(sr4, hr4), psnr, rtol=1e-6)
        assert torch.allclose(MO.ssim(sr4, hr4), ssim_avg, rtol=1e-6) and torch.allclose(MO.ssim(sr4, hr4, 11, False), ssim_img, rtol=1e-6)
        out[f"B{B}"] = {"seed": seed, "checksum": float(sr4.double().sum() + hr4.double().sum()), "psnr": psnr, "ssim": ssim_avg, "ssim_per_image": ssim_img}
        print(B, float(psnr), float(ssim_avg))
    assert ref.calculate_psnr(hr4, hr4) == float("inf")
    gd = synth.GOLDEN_DIR
    torch.save(out, gd / "metrics.pt")
    h = hashlib.sha256((gd / "metrics.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "metrics.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  metrics.pt"]) + "\n")


if __name__ == "__main__":
    main()
