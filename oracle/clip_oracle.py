"""TEST INFRASTRUCTURE - oracle of the CCR-CLIP contrastive head (never imported by the product path).

Restates image-ids-CTR/CCR-CLIP/model.py:209-222 (feature normalisation, logit_scale.exp()) and main.py:98-110 (the two logit
matrices, ground truth, symmetric cross entropy) in torch fp32 / fp64.  Parity pin: `tests/test_oracle_golden_more.py::
test_clip_oracle_matches_the_reference_loop` runs the reference's OWN lines - `CLIP.forward`'s tail and the loss lines of main.py,
executed from /root/reference with stand-in towers - against this restatement when the reference checkout is present."""
from typing import Sequence

import torch
import torch.nn.functional as F


def ground_truth(labels: Sequence[str]) -> torch.Tensor:                      # main.py:101-105
    label_str = "".join(labels)
    gt = torch.arange(len(labels), dtype=torch.long)
    for i in range(len(labels)):
        gt[i] = label_str.index(labels[i])
    return gt


def contrastive_loss(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor, gt: torch.Tensor):
    """un-normalised features -> (loss, logits_per_image)"""
    i_n = image_features / image_features.norm(dim=1, keepdim=True)            # model.py:217-218
    t_n = text_features / text_features.norm(dim=1, keepdim=True)
    scale = logit_scale.exp()                                                 # model.py:219
    logits_per_image = scale * i_n @ t_n.t()                                  # main.py:98
    logits_per_text = logits_per_image.t()                                    # main.py:99
    loss = (F.cross_entropy(logits_per_image, gt) + F.cross_entropy(logits_per_text, gt)) / 2   # main.py:106
    return loss, logits_per_image
