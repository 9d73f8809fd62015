"""Golden outputs of Pillow's Image.resize(size, Image.BICUBIC) + ToTensor on deterministic crops (the reference's
resizeNormalize, scene-text-telescope/dataset/dataset.py:136-152), and the check that oracle/resize_oracle.py reproduces
Pillow bit for bit.  Needs PIL (build container)."""
import hashlib
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import resize_oracle as R, synth  # noqa: E402


def main():
    crops = R.synth_crops(24, seed=7)
    outs = {}
    for size in ((128, 32), (64, 16)):
        for i, c in enumerate(crops):
            ref = np.asarray(Image.fromarray(c, "RGB").resize(size, Image.BICUBIC))
            mine = R.resize_bicubic_u8(c, size)
            assert np.array_equal(ref, mine), (i, c.shape, size, np.abs(ref.astype(int) - mine.astype(int)).max())
            outs[f"{size[0]}x{size[1]}_{i}"] = ref
    # the synthetic-LR path of alignCollate_syn: first a //4 down-sample at the crop's own size
    c = crops[3]
    small = np.asarray(Image.fromarray(c, "RGB").resize((c.shape[1] // 4, c.shape[0] // 4), Image.BICUBIC))
    assert np.array_equal(small, R.resize_bicubic_u8(c, (c.shape[1] // 4, c.shape[0] // 4)))
    gd = synth.GOLDEN_DIR
    np.savez_compressed(gd / "resize.npz", **outs)
    h = hashlib.sha256((gd / "resize.npz").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "resize.npz" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  resize.npz"]) + "\n")
    print("resize golden:", len(outs), "images; oracle == Pillow", Image.__version__ if hasattr(Image, "__version__") else "")


if __name__ == "__main__":
    main()
