"""Device-side collate for the TextZoom crops — drop-in for ``resizeNormalize`` / ``alignCollate_real`` /
``alignCollate_syn`` of scene-text-telescope (dataset/dataset.py:136-152, :231-270; text-gestalt has the same classes).

The reference resizes every crop on the CPU through Pillow (``img.resize((W, H), Image.BICUBIC)`` + ``ToTensor``), one Python
call per image, 2 x B per batch.  Here the ragged uint8 crops of a batch are packed into ONE pinned host buffer, copied to
the device once, and ``focr_resize_bicubic_normalize`` (csrc/resize.cu) produces the dense fp32 NCHW batch with one CTA per
crop - bit-exact with Pillow's 8-bit resampler (tests/test_gpu_resize.py).  Outputs are CUDA tensors (the reference returns
CPU tensors that the train loop then ``.to(device)``s - ``interfaces/super_resolution.py:57-58`` - which is a no-op here).
No CPU fallback: without a CUDA device the collate raises ``FocrError``.  ``mask=True`` (4th channel) is rejected like the
model's ``mask=True`` (the reference default is False, ``interfaces/base.py:141-142``)."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L

__all__ = ["resizeNormalize", "alignCollate_syn", "alignCollate_real", "resize_normalize_batch", "pack_crops",
           "hostCollate_real", "hostCollate_syn", "finish_collate"]


def _as_u8_hwc(img) -> np.ndarray:
    """PIL image (any mode convertible to RGB) or (H, W, 3) uint8 array -> contiguous (H, W, 3) uint8"""
    if isinstance(img, np.ndarray):
        a = img
    elif isinstance(img, torch.Tensor):
        a = img.cpu().numpy()
    else:  # PIL.Image without importing PIL here
        if getattr(img, "mode", "RGB") != "RGB":
            img = img.convert("RGB")
        a = np.asarray(img)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"crop must be (H, W, 3) uint8, got {a.dtype} {a.shape}")
    if a.shape[0] < 1 or a.shape[1] < 1:
        raise ValueError(f"empty crop {a.shape}")
    return np.ascontiguousarray(a)


def pack_crops(images: Sequence) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
    """ragged crops -> (pinned uint8 buffer, int64 [B][3] = {byte offset, h, w}, max_h, max_w); host-side only"""
    arrs = [_as_u8_hwc(im) for im in images]
    meta = np.zeros((len(arrs), 3), np.int64)
    off = 0
    for i, a in enumerate(arrs):
        meta[i] = (off, a.shape[0], a.shape[1])
        off += a.size
    buf = torch.empty(max(off, 1), dtype=torch.uint8)
    if torch.cuda.is_available() and torch.utils.data.get_worker_info() is None:
        buf = buf.pin_memory()   # (a forked DataLoader worker must not create a CUDA context: finish_collate pins instead)
    flat = buf.numpy()
    for (o, h, w), a in zip(meta, arrs):
        flat[o:o + a.size] = a.reshape(-1)
    max_h = int(meta[:, 1].max()) if len(arrs) else 0
    max_w = int(meta[:, 2].max()) if len(arrs) else 0
    return buf, torch.from_numpy(meta), max_h, max_w


def resize_normalize_batch(images: Sequence, size: Tuple[int, int], device=None, packed=None) -> torch.Tensor:
    """[resizeNormalize(size)(im) for im in images] stacked: (B, 3, size[1], size[0]) fp32 in [0, 1] on `device`.
    `packed` = a (device pixels, device meta, max_h, max_w) tuple re-uses an upload (HR and LR from the same crops)."""
    if torch.utils.data.get_worker_info() is not None:
        raise L.FocrError(
            "focr collate was called inside a DataLoader worker process: CUDA cannot be (re)initialised in a forked worker. "
            "Use num_workers=0 with alignCollate_real / alignCollate_syn, or keep the workers and give the loader "
            "hostCollate_real / hostCollate_syn (pack the crops in the worker) and call finish_collate(batch) in the main process.")
    if not torch.cuda.is_available():
        raise L.FocrError("focr collate runs on a CUDA device only (no CPU fallback)")
    dev = torch.device("cuda" if device is None else device)
    ow, oh = int(size[0]), int(size[1])
    if packed is None:
        if len(images) == 0:
            return torch.empty(0, 3, oh, ow, dtype=torch.float32, device=dev)
        buf, meta, max_h, max_w = pack_crops(images)
        packed = (buf.to(dev, non_blocking=True), meta.to(dev, non_blocking=True), max_h, max_w)
    pix, meta_d, max_h, max_w = packed
    B = meta_d.shape[0]
    out = torch.empty(B, 3, oh, ow, dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib.focr_resize_bicubic_normalize(pix.data_ptr(), meta_d.data_ptr(), B, max_h, max_w, ow, oh, out.data_ptr(),
                                                    status.data_ptr(), L.cur_stream()), "resize_bicubic_normalize")
    if L.status_checks() and int(status.item()):
        raise ValueError("resize_bicubic_normalize: a crop exceeds the kernel's shared-memory budget (its output was left zero)")
    return out


class resizeNormalize(object):
    """dataset.py:136-152 - callable on ONE image; returns a (3, H, W) CUDA tensor"""

    def __init__(self, size, mask=False, interpolation=3):  # 3 == PIL.Image.BICUBIC
        if mask:
            raise NotImplementedError("focr resizeNormalize: mask=True (4-channel input) is not supported")
        if interpolation != 3:
            raise NotImplementedError("focr resizeNormalize: only Image.BICUBIC (the reference's only use)")
        self.size = size
        self.interpolation = interpolation
        self.mask = mask

    def __call__(self, img):
        return resize_normalize_batch([img], self.size)[0]


class alignCollate_syn(object):
    """dataset.py:231-254 - HR = resize(crop); LR = resize(resize(crop, crop.size // scale)): two chained Pillow resamples
    (uint8 in between), both on the device"""

    def __init__(self, imgH=64, imgW=256, down_sample_scale=4, keep_ratio=False, min_ratio=1, mask=False):
        if mask:
            raise NotImplementedError("focr collate: mask=True (4-channel input) is not supported")
        self.imgH = imgH
        self.imgW = imgW
        self.keep_ratio = keep_ratio
        self.min_ratio = min_ratio
        self.down_sample_scale = down_sample_scale
        self.mask = mask

    def __call__(self, batch):
        images, label_strs = zip(*batch)
        s = self.down_sample_scale
        images_hr = resize_normalize_batch(images, (self.imgW, self.imgH))
        # the intermediate //scale image has a per-crop size: one launch per distinct size class would fragment the
        # batch, so the first resample of this (synthetic-data) path goes crop by crop through the same kernel
        smalls = []
        for im in images:
            a = _as_u8_hwc(im)
            t = resize_normalize_batch([a], (a.shape[1] // s, a.shape[0] // s))[0]
            smalls.append((t * 255.0).round().to(torch.uint8).permute(1, 2, 0).contiguous())
        images_lr = resize_normalize_batch([t.cpu().numpy() for t in smalls], (self.imgW // s, self.imgH // s))
        return images_hr, images_lr, label_strs


class alignCollate_real(alignCollate_syn):
    """dataset.py:257-270 - the TextZoom path: separate HR and LR crops, each resized to its target size"""

    def __call__(self, batch):
        images_HR, images_lr, label_strs = zip(*batch)
        s = self.down_sample_scale
        images_HR = resize_normalize_batch(images_HR, (self.imgW, self.imgH))
        images_lr = resize_normalize_batch(images_lr, (self.imgW // s, self.imgH // s))
        return images_HR, images_lr, label_strs


# ---- worker-safe split: the host half runs inside DataLoader workers (no CUDA), the device half in the main process -------------
class hostCollate_real(object):
    """collate_fn for DataLoaders with num_workers > 0 (the reference's loaders use 8, interfaces/base.py:103-108): packs the
    ragged HR / LR crops of a batch into two uint8 buffers + meta tables and touches no CUDA API; `finish_collate` turns
    the result into the (images_HR, images_lr, label_strs) triple of alignCollate_real in the main process"""
    kind = "real"

    def __init__(self, imgH=64, imgW=256, down_sample_scale=4, keep_ratio=False, min_ratio=1, mask=False):
        if mask:
            raise NotImplementedError("focr collate: mask=True (4-channel input) is not supported")
        self.imgH, self.imgW, self.down_sample_scale = imgH, imgW, down_sample_scale

    def __call__(self, batch):
        images_HR, images_lr, label_strs = zip(*batch)
        return {"kind": self.kind, "hr": pack_crops(images_HR), "lr": pack_crops(images_lr), "labels": label_strs,
                "size": (self.imgW, self.imgH), "scale": self.down_sample_scale}


class hostCollate_syn(hostCollate_real):
    """host half of alignCollate_syn: the single crop list travels; both resamples run on the device in finish_collate"""
    kind = "syn"

    def __call__(self, batch):
        images, label_strs = zip(*batch)
        return {"kind": self.kind, "hr": [_as_u8_hwc(im) for im in images], "labels": label_strs,
                "size": (self.imgW, self.imgH), "scale": self.down_sample_scale}


def finish_collate(batch: dict, device=None):
    """device half (main process): -> (images_HR, images_lr, label_strs) exactly as alignCollate_real / alignCollate_syn"""
    W, H = batch["size"]
    s = batch["scale"]
    if batch["kind"] == "syn":
        hr, lr, labels = alignCollate_syn(H, W, s)([(im, lab) for im, lab in zip(batch["hr"], batch["labels"])])
        return hr, lr, labels
    dev = torch.device("cuda" if device is None else device)

    def up(packed):
        buf, meta, mh, mw = packed
        if not buf.is_pinned():
            buf = buf.pin_memory()
        return buf.to(dev, non_blocking=True), meta.to(dev, non_blocking=True), mh, mw
    hr = resize_normalize_batch(None, (W, H), dev, packed=up(batch["hr"]))
    lr = resize_normalize_batch(None, (W // s, H // s), dev, packed=up(batch["lr"]))
    return hr, lr, batch["labels"]
