"""Generates tests/golden/resample_small.npz from the REFERENCE implementation of the step (PyTorch on
CPU: torch.sigmoid + F.interpolate(bilinear, align_corners=True), as at
/root/reference/octsam/models/training_utils.py:64 and topological_loss.py:33-46), including its
autograd gradient.  Run here; the GPU box only reads the .npz."""
import os

import numpy as np
import torch
import torch.nn.functional as F

rng = np.random.default_rng(2024)
out = {}
for k, (H, W, S) in enumerate([(31, 37, 10), (64, 64, 50), (124, 128, 50), (50, 50, 50), (20, 24, 33)]):
    x = (3.0 * rng.standard_normal((1, 2, H, W))).astype(np.float32)
    g = rng.standard_normal((1, 2, S, S)).astype(np.float32)
    for sig in (0, 1):
        t = torch.from_numpy(x).clone().requires_grad_(True)
        y = F.interpolate(torch.sigmoid(t) if sig else t, size=(S, S), mode="bilinear", align_corners=True)
        y.backward(torch.from_numpy(g))
        if H * W <= 64 * 64:
            out[f"x{k}"] = x
            out[f"g{k}"] = g
            out[f"y{k}_{sig}"] = y.detach().numpy()
            out[f"gx{k}_{sig}"] = t.grad.numpy()
        else:  # large case: keep a strided sample of the input gradient only
            out[f"x{k}"] = x[:1, :1]
            out[f"g{k}"] = g[:1, :1]
            t = torch.from_numpy(x[:1, :1]).clone().requires_grad_(True)
            y = F.interpolate(torch.sigmoid(t) if sig else t, size=(S, S), mode="bilinear", align_corners=True)
            y.backward(torch.from_numpy(g[:1, :1]))
            out[f"y{k}_{sig}"] = y.detach().numpy()
            out[f"gx{k}_{sig}"] = t.grad.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resample_small.npz"), **out)
print({k: v.shape for k, v in out.items()})
