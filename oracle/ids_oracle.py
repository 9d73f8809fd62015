"""ORACLE (test infrastructure only).  torch-fp32 functional restatement of the image-ids-CTR recogniser and its training
step (paths relative to /root/reference/image-ids-CTR):
  ResNet encoder with four pooling steps, layer4 unused   model/transformer.py:72-152
  decoder / embeddings / LayerNorm / attention            :154-327 (same code as stroke-level-decomposition; restated in
                                                          oracle/sld_oracle.py and reused here)
  Transformer.forward, Generator(1024, 2048), packing      :329-382
  train step: L2-normalise, @ text_features^T, CE + 0.001 * (-MSE), Adadelta(lr 1, rho 0.9, weight_decay 1e-4)   train.py:28,63-90
  label tensors                                           util.py:101-127
Arithmetic lives in torch (env.yaml pins torch==1.10.1; 2.11 here).  Pinned by oracle/make_golden_ids.py against the UNMODIFIED
reference module (tests/golden/ids_b4.pt).  The CCR-CLIP text features are git-ignored assets: a seeded stand-in with the same
construction (zero row for START, ones row for END, train.py:52-62) replaces them."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import sld_oracle as SO

N_CLASS = 4303   # START + the 4301 characters of data/char_document_Chinese.txt + END (util.py:12-22)
LAYERS = (("layer1", 3), ("layer2", 4), ("layer3", 6))


def encoder(sd, image, train=True, stats_out=None):                    # model/transformer.py:126-152
    e = "encoder."
    x = F.relu(SO._bn(sd, e + "bn1", SO._conv(sd, e + "conv1", image), train, stats_out))
    x = F.max_pool2d(x, (2, 2), (2, 2))
    x = F.relu(SO._bn(sd, e + "bn2", SO._conv(sd, e + "conv2", x), train, stats_out))
    for name, n in LAYERS:
        x = F.max_pool2d(x, (2, 2), (2, 2))
        for i in range(n):
            x = SO.basic_block(sd, f"{e}{name}.{i}", x, train, stats_out)
        x = F.relu(SO._bn(sd, f"{e}{name}_bn", SO._conv(sd, f"{e}{name}_conv", x), train, stats_out))
    return x


def forward(sd, image, text_input, train=True, conv_feature=None, stats_out=None):
    """-> (pred (B, T, 2048), attention map (B, 4, T, tokens), conv feature (B, 1024, H/16, W/16))   model/transformer.py:341-382"""
    if conv_feature is None:
        conv_feature = encoder(sd, image, train, stats_out)
    emb = F.embedding(text_input, sd["embedding_word.lut.weight"]) * (512 ** 0.5)
    pe = SO.positional_encoding(text_input.shape[1], 512, emb.device).unsqueeze(0).expand(emb.shape[0], -1, -1)
    x, amap = SO.decoder(sd, torch.cat([emb, pe], 2), conv_feature, None)
    pred = F.linear(x, sd["generator_word.proj.weight"], sd["generator_word.proj.bias"])
    return pred, amap, conv_feature


def loss_fn(sd, image, length, text_input, text_gt, text_features, stats_out=None):   # train.py:63-80
    pred, amap, conv = forward(sd, image, text_input, True, None, stats_out)
    text_pred = SO.pack(pred, length)
    reg = text_features[text_gt]
    text_pred = text_pred / text_pred.norm(dim=1, keepdim=True)
    final_res = text_pred @ text_features.t()
    loss_rec = F.cross_entropy(final_res, text_gt)
    loss_dis = -F.mse_loss(text_pred, reg)
    return loss_rec + 0.001 * loss_dis, loss_rec, loss_dis, pred, amap, conv


def converter(labels):                                                  # util.py:101-127 with index lists for the characters
    """labels: list of lists of class indices (1..4303), each sample's last position standing for 'END' ->
    (length, text_input, text_gt) as util.converter builds them"""
    length = torch.tensor([len(s) for s in labels], dtype=torch.long)
    T = int(length.max())
    text_input = torch.zeros(len(labels), T, dtype=torch.long)
    gt = []
    for i, s in enumerate(labels):
        for j in range(len(s) - 1):
            text_input[i, j + 1] = s[j]
        gt.extend(list(s[:-1]) + [N_CLASS - 1])
    return length, text_input, torch.tensor(gt, dtype=torch.long)


def synth_text_features(seed: int = 77, n_class: int = N_CLASS):
    """stand-in for the CCR-CLIP text features: zero row (START), one row per character, ones row (END)  (train.py:52-62)"""
    rs = np.random.RandomState(seed)
    f = rs.standard_normal((n_class, 2048)).astype(np.float32) * 0.3
    f[0] = 0.0
    f[-1] = 1.0
    return torch.from_numpy(f)


def synth_batch(B: int, seed: int = 4321, width: int = 256):
    rs = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, 32), np.linspace(0, 1, width), indexing="ij")
    img = np.zeros((B, 3, 32, width), np.float64)
    for b in range(B):
        for c in range(3):
            f = np.zeros_like(yy)
            for _ in range(6):
                f += rs.uniform(0.3, 1.0) * np.cos(2 * np.pi * (rs.uniform(1, 12) * xx + rs.uniform(0.5, 3) * yy) + rs.uniform(0, 6.28))
            for _ in range(8):
                x0, wd = rs.uniform(0.02, 0.98), rs.uniform(0.005, 0.02)
                f += rs.uniform(1, 2) * np.exp(-((xx - x0) / wd) ** 2)
            f = (f - f.min()) / (f.max() - f.min() + 1e-9)
            img[b, c] = 2 * f - 1 + 0.02 * rs.standard_normal(f.shape)
    labels = [list(rs.randint(1, N_CLASS - 1, size=int(rs.randint(2, 8)))) for _ in range(B)]
    return torch.from_numpy(np.clip(img, -1, 1).astype(np.float32)), [[int(v) for v in s] for s in labels]
