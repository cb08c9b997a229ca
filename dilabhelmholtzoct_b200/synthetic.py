"""Synthetic OCT-like inputs for tests and bench.py (there is no dataset
offline -- the reference's 552-image OCT set is private, README.md:17).

truth  one-hot {0.,1.} float32 [B, 14, H, W]: a padding band (class 13) on top, 9-13 stacked
       retinal layers with smooth random interfaces, 0-6 elliptical fluid blobs (classes 3/4/7).
pred   per-channel sigmoid (as training_utils.py:64, not a softmax) of a blurred signed one-hot
       plus low-pass and iid noise -> roughly 1-3 k H1 pairs per 256x256 map.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

N_CLASSES = 14


def _gauss1d(sigma: float, device) -> torch.Tensor:
    r = max(1, int(math.ceil(3 * sigma)))
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _blur(x: torch.Tensor, sigma: float) -> torch.Tensor:
    """Separable Gaussian blur of [B, C, H, W] with reflect padding."""
    k = _gauss1d(sigma, x.device)
    r = (k.numel() - 1) // 2
    B, C, H, W = x.shape
    y = x.reshape(B * C, 1, H, W)
    r_h, r_w = min(r, H - 1), min(r, W - 1)
    kh = k[r - r_h: r + r_h + 1] / k[r - r_h: r + r_h + 1].sum()
    kw = k[r - r_w: r + r_w + 1] / k[r - r_w: r + r_w + 1].sum()
    y = F.conv2d(F.pad(y, (r_w, r_w, 0, 0), mode="reflect"), kw.view(1, 1, 1, -1))
    y = F.conv2d(F.pad(y, (0, 0, r_h, r_h), mode="reflect"), kh.view(1, 1, -1, 1))
    return y.reshape(B, C, H, W)


def make_labels(B: int, H: int, W: int, gen: torch.Generator, n_classes: int = N_CLASSES) -> torch.Tensor:
    """int64 label maps [B, H, W] on the generator's device."""
    dev = gen.device
    labels = torch.empty((B, H, W), dtype=torch.int64, device=dev)
    rows = torch.arange(H, device=dev, dtype=torch.float32).view(H, 1)
    cols = torch.arange(W, device=dev, dtype=torch.float32).view(1, W)
    for b in range(B):
        n_layers = int(torch.randint(min(9, n_classes - 1), n_classes, (1,), generator=gen, device=dev))
        # interfaces: cumulative sums of positive low-pass 1-D noise
        raw = torch.rand((n_layers + 1, W + 32), generator=gen, device=dev)
        k = _gauss1d(max(2.0, W / 16), dev)
        thick = F.conv1d(raw.unsqueeze(1), k.view(1, 1, -1), padding=k.numel() // 2).squeeze(1)[:, 16:16 + W]
        thick = 0.3 + thick / thick.mean()
        depth = torch.cumsum(thick, 0)
        depth = depth / depth[-1:].clamp_min(1e-6) * (H * 0.92) + H * 0.04  # [n_layers+1, W] row of each interface
        lab = torch.full((H, W), n_classes - 1, dtype=torch.int64, device=dev)  # class 13 above the retina
        for li in range(n_layers):
            below = rows >= depth[li].view(1, W)
            lab = torch.where(below, torch.full_like(lab, li % (n_classes - 1)), lab)
        lab = torch.where(rows >= depth[n_layers].view(1, W), torch.full_like(lab, 0), lab)
        n_blobs = int(torch.randint(0, 7, (1,), generator=gen, device=dev))
        for _ in range(n_blobs):
            u = torch.rand(5, generator=gen, device=dev)
            cy, cx = H * (0.2 + 0.6 * float(u[0])), W * (0.1 + 0.8 * float(u[1]))
            ry, rx = max(1.5, H * (0.01 + 0.05 * float(u[2]))), max(1.5, W * (0.015 + 0.08 * float(u[3])))
            cls = (3, 4, 7)[int(float(u[4]) * 3) % 3] % max(1, n_classes - 1)
            inside = ((rows - cy) / ry) ** 2 + ((cols - cx) / rx) ** 2 <= 1.0
            lab = torch.where(inside, torch.full_like(lab, cls), lab)
        labels[b] = lab
    return labels


def make_batch(B: int, H: int, W: int, seed: int, device="cpu", n_classes: int = N_CLASSES):
    """(pred, truth) float32 [B, n_classes, H, W]; deterministic in (seed, device type)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    labels = make_labels(B, H, W, gen, n_classes)
    truth = F.one_hot(labels, n_classes).permute(0, 3, 1, 2).to(torch.float32).contiguous()
    signed = _blur(4.0 * (2.0 * truth - 1.0), 2.0)
    low = _blur(torch.randn((B, n_classes, H, W), generator=gen, device=device), 4.0)
    low = low / low.std().clamp_min(1e-6)
    iid = torch.randn((B, n_classes, H, W), generator=gen, device=device)
    pred = torch.sigmoid(signed + 1.0 * low + 0.3 * iid).contiguous()
    return pred, truth
