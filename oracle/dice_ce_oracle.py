"""Oracle for row F2 -- restatement of ``monai.losses.DiceCELoss(sigmoid=True)``.  TEST INFRASTRUCTURE ONLY.

The reference builds ``seg_loss = monai.losses.DiceCELoss(sigmoid=True)`` (/root/reference/octsam/models/
training_utils.py:32) and calls ``seg_loss(masks, gt_masks)`` (:62) on ``[B, Nmax, H, W]`` logits / masks.
monai is pinned (1.3.0, /root/reference/environment.yml:224) but NOT installed in the build container and
cannot be fetched: PARITY UNPINNED for this row.  Restated from monai 1.3.0's published source
(``monai/losses/dice.py``: ``DiceLoss.forward`` and ``DiceCELoss.ce`` / ``.forward``), defaults only:

* ``DiceLoss(sigmoid=True, include_background=True, squared_pred=False, jaccard=False, reduction="mean",
  smooth_nr=1e-5, smooth_dr=1e-5, batch=False)``: ``s = sigmoid(x)``; per (b, c), reduced over the spatial axes,
  ``f = 1 - (2 * sum(s * t) + smooth_nr) / (sum(t) + sum(s) + smooth_dr)``; mean over (b, c);
* ``DiceCELoss.ce``: target has as many channels as the input, so it is used as class PROBABILITIES:
  ``torch.nn.CrossEntropyLoss(reduction="mean")(x, t)`` = mean over (b, pixels) of ``-sum_c t_c * log_softmax(x)_c``;
* ``lambda_dice = lambda_ce = 1``: ``loss = dice + ce``.

(monai >= 1.3.1 switches single-channel inputs to BCE-with-logits; that branch does not apply to Nmax > 1 and
is not restated.)  The oracle is what ``tests/test_dice_ce.py`` checks the CUDA kernel against; the PyTorch
calls it is made of are the reference's own arithmetic for everything below monai's glue.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dice_ce(logits: torch.Tensor, target: torch.Tensor, smooth_nr: float = 1e-5, smooth_dr: float = 1e-5) -> torch.Tensor:
    """DiceCELoss(sigmoid=True)(logits, target) for [B, C, *spatial] tensors, in the dtype of ``logits``."""
    target = target.to(logits.dtype)
    s = torch.sigmoid(logits)
    axes = tuple(range(2, logits.dim()))
    inter = (s * target).sum(axes)
    denom = target.sum(axes) + s.sum(axes)
    dice = (1.0 - (2.0 * inter + smooth_nr) / (denom + smooth_dr)).mean()
    ce = F.cross_entropy(logits, target)  # probabilities as targets: mean over batch x pixels
    return dice + ce


def dice_ce_explicit(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """The same value without F.cross_entropy (pins the reading of 'probability targets, mean reduction')."""
    target = target.to(logits.dtype)
    B, C = logits.shape[:2]
    x = logits.reshape(B, C, -1)
    t = target.reshape(B, C, -1)
    lse = torch.logsumexp(x, dim=1, keepdim=True)
    ce = -(t * (x - lse)).sum(1).mean()
    s = torch.sigmoid(x)
    dice = (1.0 - (2.0 * (s * t).sum(2) + 1e-5) / (t.sum(2) + s.sum(2) + 1e-5)).mean()
    return dice + ce
