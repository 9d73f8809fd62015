"""PSNR / SSIM kernel (csrc/metrics.cu, drop-in fudanocr_b200.utils.ssim_psnr) vs the golden values recorded from the
reference's utils/ssim_psnr.py and vs the oracle restatement on the GPU.  fp32 throughout: 1e-5 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_psnr_ssim_vs_golden_and_oracle():
    from oracle import synth, metrics_oracle as MO
    from fudanocr_b200.utils import ssim_psnr as M
    g = torch.load(synth.GOLDEN_DIR / "metrics.pt", weights_only=False)
    for key, rec in g.items():
        sr, hr = MO.synth_pair(int(key[1:]), rec["seed"])
        sr, hr = sr.to(DEV), hr.to(DEV)
        psnr, ssim_avg, per = M.psnr_ssim(sr, hr, per_image=True)
        assert abs(float(psnr) - float(rec["psnr"])) < 1e-5 * float(rec["psnr"]), (float(psnr), float(rec["psnr"]))
        assert abs(float(ssim_avg) - float(rec["ssim"])) < 1e-5, (float(ssim_avg), float(rec["ssim"]))
        assert torch.allclose(per.cpu(), rec["ssim_per_image"], rtol=1e-5, atol=1e-6)
        assert abs(float(M.calculate_psnr(sr, hr)) - float(rec["psnr"])) < 1e-5 * float(rec["psnr"])
        assert abs(float(M.SSIM()(sr, hr)) - float(rec["ssim"])) < 1e-5
        assert torch.allclose(M.ssim(sr, hr, size_average=False).cpu(), rec["ssim_per_image"], rtol=1e-5, atol=1e-6)
    # identical images: PSNR = inf, SSIM = 1 (edge case of the reference: `if mse == 0: return inf`)
    p, s = M.psnr_ssim(hr, hr)
    assert float(p) == float("inf") and abs(float(s) - 1.0) < 1e-6
    # a larger random batch against the oracle evaluated on the GPU
    torch.backends.cudnn.allow_tf32 = False
    a, b = torch.rand(64, 3, 32, 128, device=DEV), torch.rand(64, 3, 32, 128, device=DEV)
    p, s = M.psnr_ssim(a, b)
    assert abs(float(p) - float(MO.calculate_psnr(a, b))) < 1e-5 * float(p)
    assert abs(float(s) - float(MO.ssim(a, b))) < 1e-5


def test_metrics_reject_cpu_and_bad_shapes():
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.utils import ssim_psnr as M
    with pytest.raises(L.FocrError):
        M.psnr_ssim(torch.rand(2, 3, 32, 128), torch.rand(2, 3, 32, 128))
    with pytest.raises(ValueError):
        M.psnr_ssim(torch.rand(2, 3, 16, 64, device=DEV), torch.rand(2, 3, 16, 64, device=DEV))
