"""Device-side collate (csrc/resize.cu, drop-in fudanocr_b200.dataset) vs Pillow's 8-bit bicubic resampler: golden outputs
recorded from PIL (tests/golden/resize.npz) and the numpy oracle on fresh ragged batches.  Integer path: bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _u8(t):
    """fp32 (B,3,H,W) in [0,1] -> uint8 (B,H,W,3); exact because the kernel emits k/255 for integer k"""
    return (t * 255.0).round().to(torch.uint8).permute(0, 2, 3, 1).cpu().numpy()


def test_resize_matches_pillow_golden_bit_exact():
    from oracle import synth, resize_oracle as R
    from fudanocr_b200 import dataset as D
    g = np.load(synth.GOLDEN_DIR / "resize.npz")
    crops = R.synth_crops(24, seed=7)
    for size in ((128, 32), (64, 16)):
        out = D.resize_normalize_batch(crops, size)
        assert out.shape == (len(crops), 3, size[1], size[0]) and out.dtype == torch.float32 and out.is_cuda
        got = _u8(out)
        for i in range(len(crops)):
            assert np.array_equal(got[i], g[f"{size[0]}x{size[1]}_{i}"]), (i, crops[i].shape, size)
        # ToTensor semantics: value = uint8 / 255 in fp32, exactly
        ref0 = torch.from_numpy(R.resize_normalize(crops[0], size))
        assert torch.equal(out[0].cpu(), ref0)


def test_resize_ragged_batch_vs_oracle_and_edge_cases():
    from oracle import resize_oracle as R
    from fudanocr_b200 import dataset as D
    rs = np.random.RandomState(11)
    shapes = [(1, 1), (1, 300), (300, 1), (2, 2), (16, 64), (32, 128), (15, 63), (17, 65), (33, 129), (64, 256), (5, 500),
              (120, 40), (31, 127)]
    crops = [rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in shapes]
    crops += [np.zeros((20, 90, 3), np.uint8), np.full((20, 90, 3), 255, np.uint8)]   # clip8 at both ends
    for size in ((128, 32), (64, 16), (100, 32)):
        got = _u8(D.resize_normalize_batch(crops, size))
        for i, c in enumerate(crops):
            assert np.array_equal(got[i], R.resize_bicubic_u8(c, size)), (i, c.shape, size)
    assert D.resize_normalize_batch([], (128, 32)).shape == (0, 3, 32, 128)
    with pytest.raises(ValueError):
        D.resize_normalize_batch([np.zeros((4, 4), np.uint8)], (128, 32))
    with pytest.raises(NotImplementedError):
        D.resizeNormalize((128, 32), mask=True)


def test_collate_classes_match_reference_semantics():
    """alignCollate_real / alignCollate_syn (dataset.py:231-270) on a batch of (HR, LR, label) / (image, label) tuples"""
    from oracle import resize_oracle as R
    from fudanocr_b200 import dataset as D
    crops = R.synth_crops(6, seed=5)
    lrs = [R.resize_bicubic_u8(c, (max(c.shape[1] // 2, 1), max(c.shape[0] // 2, 1))) for c in crops]
    labels = [f"w{i}" for i in range(len(crops))]
    hr, lr, lab = D.alignCollate_real(imgH=32, imgW=128, down_sample_scale=2)(list(zip(crops, lrs, labels)))
    assert tuple(lab) == tuple(labels) and hr.shape == (len(crops), 3, 32, 128) and lr.shape == (len(crops), 3, 16, 64)
    for i in range(len(crops)):
        assert np.array_equal(_u8(hr)[i], R.resize_bicubic_u8(crops[i], (128, 32)))
        assert np.array_equal(_u8(lr)[i], R.resize_bicubic_u8(lrs[i], (64, 16)))
    hr2, lr2, _ = D.alignCollate_syn(imgH=32, imgW=128, down_sample_scale=2)(list(zip(crops, labels)))
    assert torch.equal(hr2, hr)
    for i, c in enumerate(crops):
        small = R.resize_bicubic_u8(c, (c.shape[1] // 2, c.shape[0] // 2))
        assert np.array_equal(_u8(lr2)[i], R.resize_bicubic_u8(small, (64, 16)))
    one = D.resizeNormalize((128, 32))(crops[0])
    assert torch.equal(one, hr[0])


def test_resize_full_batch_properties():
    """BASELINE batch size (256 crops): constant crops stay constant, identity at the target size, deterministic"""
    from fudanocr_b200 import dataset as D
    rs = np.random.RandomState(2)
    crops = [np.full((int(rs.randint(8, 80)), int(rs.randint(16, 320)), 3), int(rs.randint(0, 256)), np.uint8) for _ in range(128)]
    ident = [rs.randint(0, 256, size=(32, 128, 3)).astype(np.uint8) for _ in range(128)]
    out = D.resize_normalize_batch(crops + ident, (128, 32))
    got = _u8(out)
    for i, c in enumerate(crops):
        assert (got[i] == c[0, 0, 0]).all(), i
    for i, c in enumerate(ident):
        assert np.array_equal(got[128 + i], c)
    assert torch.equal(out, D.resize_normalize_batch(crops + ident, (128, 32)))
