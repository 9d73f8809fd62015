"""image-ids-CTR recogniser on the focr engine - drop-in for ``model.transformer.Transformer`` of image-ids-CTR
(model/transformer.py:329-382): the ResNet encoder with FOUR pooling steps (:72-152; (B,3,32,256) -> (B,1024,2,16) = 32 image
tokens; ``layer4`` / ``layer4_conv2`` are constructed but never called, :103-107), the same decoder layer as
stroke-level-decomposition, and a 2048-d generator whose outputs are L2-normalised and matched against frozen CCR-CLIP text
features (train.py:63-90):  loss = CE(pred_n @ text_features^T, gt) + 0.001 * (-MSE(pred_n, text_features[gt])).
SURVEY.md §8 row A22.  Same state_dict keys as the reference (incl. the dead layers), same forward contract; the arithmetic
is the kernel-backed autograd nodes of model/transformer.py.  CUDA only; no CPU fallback."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import recog_ops as ops
from .transformer import (_AddRelu, _BasicBlock, _Container, _Conv, _ConvFirst, _Decoder, _Embeddings, _FeatMSE, _Generator,
                          _L2Norm, _Linear, _MaxPool, _PackedCE, _PositionalEncoding, _ResNet, Transformer as _SLDTransformer)

N_CLASS = 4303   # 'START' + the 4301 characters of data/char_document_Chinese.txt + 'END' (util.py:12-22)


class _ResNetIDS(_ResNet):  # image-ids-CTR/model/transformer.py:72-152
    def __init__(self, num_in, layers):
        _Container.__init__(self)
        self.conv1 = nn.Conv2d(num_in, 64, 3, 1, 1)
        self.bn1 = nn.BatchNorm2d(64)
        self.conv2 = nn.Conv2d(64, 128, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(128)
        self.layer1 = self._make_layer(128, 256, layers[0])
        self.layer1_conv = nn.Conv2d(256, 256, 3, 1, 1)
        self.layer1_bn = nn.BatchNorm2d(256)
        self.layer2 = self._make_layer(256, 512, layers[1])
        self.layer2_conv = nn.Conv2d(512, 512, 3, 1, 1)
        self.layer2_bn = nn.BatchNorm2d(512)
        self.layer3 = self._make_layer(512, 1024, layers[2])
        self.layer3_conv = nn.Conv2d(1024, 1024, 3, 1, 1)
        self.layer3_bn = nn.BatchNorm2d(1024)
        self.layer4 = self._make_layer(512, 512, layers[3])          # never called by the reference forward (:126-152)
        self.layer4_conv2 = nn.Conv2d(512, 1024, 3, 1, 1)
        self.layer4_conv2_bn = nn.BatchNorm2d(1024)


class Transformer(_SLDTransformer):
    """drop-in for image-ids-CTR/model/transformer.py:329-382 (constructor takes no arguments there)"""

    def __init__(self, n_class: int = N_CLASS):
        nn.Module.__init__(self)
        self.word_n_class = n_class
        self.embedding_word = _Embeddings(512, n_class)
        self.pe = _PositionalEncoding(512)
        self.encoder = _ResNetIDS(3, [3, 4, 6, 3])
        self.decoder = _Decoder(ln_names=("a_2", "b_2"))
        self.generator_word = _Generator(1024, 2048)
        self.dropout_p = self.DROPOUT
        self._seed = (0x1D5 ^ int(torch.initial_seed())) & 0x7FFFFFFF

    def encode(self, image: torch.Tensor) -> torch.Tensor:
        """(B, 3, 32, 256) fp32 -> (B, 2, 16, 1024) bf16 NHWC (ResNet.forward, :126-152: conv-bn-relu-pool, conv-bn-relu, then three
        [pool, BasicBlocks, conv-bn-relu] stages)"""
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] % 16 or image.shape[3] % 16:
            raise ValueError(f"focr IDS Transformer: image must be (B,3,H,W) with H and W multiples of 16, got {tuple(image.shape)}")
        ops.require_cuda(image)
        e = self.encoder
        x = _ConvFirst.apply(image.float().contiguous(), e.conv1.weight, e.conv1.bias)
        x = _MaxPool.apply(self._bn(x, e.bn1, ops.ACT_RELU))
        x = self._bn(_Conv.apply(x, e.conv2.weight, e.conv2.bias), e.bn2, ops.ACT_RELU)
        for name in ("layer1", "layer2", "layer3"):
            x = _MaxPool.apply(x)
            for blk in getattr(e, name):
                x = self._block(x, blk)
            conv, bn = getattr(e, name + "_conv"), getattr(e, name + "_bn")
            x = self._bn(_Conv.apply(x, conv.weight, conv.bias), bn, ops.ACT_RELU)
        return x

    def decode(self, feat: torch.Tensor, text_input: torch.Tensor):
        """-> (pred fp32 (rows_pad, 2048), map): Generator(1024, 2048) (:337,355)"""
        r3, amap = self.decode_hidden(feat, text_input)
        g = self.generator_word.proj
        return _Linear.apply(r3, g.weight, g.bias, False, True), amap

    def forward(self, image, text_length, text_input, conv_feature: Optional[torch.Tensor] = None, test: bool = False, att_map=None):
        if conv_feature is None:
            feat = self.encode(image)
        else:
            feat = conv_feature.permute(0, 2, 3, 1).contiguous()
        conv_out = feat.permute(0, 3, 1, 2)
        if text_length is None:
            return {"conv": conv_out}
        B, T = text_input.shape
        pred, amap = self.decode(feat, text_input)
        full = pred[:B * T].view(B, T, -1)
        if test:
            return {"pred": full, "map": amap, "conv": conv_out}
        keep = torch.arange(T, device=full.device)[None, :] < text_length.to(full.device)[:, None]
        return {"pred": full[keep], "map": amap, "conv": conv_out}

    @staticmethod
    def pad_text_features(text_features: torch.Tensor) -> torch.Tensor:
        """(V, 2048) frozen CLIP features -> rows padded with zeros to a multiple of 64 (one GEMM tile); do this once"""
        V = text_features.shape[0]
        return F.pad(text_features.float(), (0, 0, 0, (-V) % 64)).contiguous()

    def loss(self, image, text_length, text_input, text_gt, text_features, text_features_padded=None):
        """fused criterion of image-ids-CTR/train.py:63-80 -> (loss, loss_rec, loss_dis) with
        loss = loss_rec + 0.001 * loss_dis, loss_rec = CE(pred_n @ text_features^T, gt), loss_dis = -MSE(pred_n, text_features[gt])"""
        feat = self.encode(image)
        B, T = text_input.shape
        pred, _ = self.decode(feat, text_input)
        y = _L2Norm.apply(pred)
        tfp = text_features_padded if text_features_padded is not None else self.pad_text_features(text_features)
        sim = _Linear.apply(y, tfp, None, False, True)
        length, gt = text_length.contiguous(), text_gt.contiguous()
        loss_rec = _PackedCE.apply(sim, B, T, text_features.shape[0], length, gt)
        loss_dis = -_FeatMSE.apply(y, B, T, length, gt, text_features.float().contiguous())
        return loss_rec + 0.001 * loss_dis, loss_rec, loss_dis
