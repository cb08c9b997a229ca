"""Data-parallel wrapper around the reference's training step for the topological loss.

The reference trains in one process on one device (``/root/reference/octsam/models/training_utils.py:27-80``,
``device = "cuda" if ... else "cpu"`` at :33).  Here the batch axis is sharded over one process per
GPU (``torch.distributed``, NCCL over NVLink): every (image, class) map is independent through
persistence and matching, and the only coupling is ``mean_b`` over images
(``/root/reference/octsam/models/topological_loss.py:85``).  So the data path needs no collective; the
only exchanges are

* an all-reduce of the scalar loss (``topo_loss_sharded``), and
* the all-reduce of the mask-decoder gradients (``DistributedDataParallel`` on
  ``model.mask_decoder`` -- the encoders are frozen, training_utils.py:277-279), or
  ``allreduce_gradients`` when DDP is not used.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import topological_loss as _tl


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  d y / d x_local = 1, so backward passes the gradient through."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def shard_batch(n_images: int, rank: int, world: int) -> slice:
    """Contiguous slice of the batch axis owned by ``rank`` (whole images: all classes of an image
    stay on one GPU because W_b couples the classes of one image)."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def topo_loss_sharded(pred_local, true_local, lamda, interp=0, feat_d=2, loss_q=2, loss_r=False, *,
                      global_batch: Optional[int] = None, group=None,
                      loss_fn: Optional[Callable] = None):
    """``topo_loss`` with the batch axis sharded over the ranks of ``group``.

    Returns the GLOBAL loss ``lamda * mean_b W_b`` (identical on every rank).  Its backward gives
    each rank ``d loss_global / d pred_local`` -- the mean's ``1 / B_global`` is applied inside the
    kernel, no gradient is communicated.  ``global_batch`` defaults to ``B_local * world_size``.

    ``loss_fn(pred, true, lamda, feat_d, loss_q, loss_r, global_batch) -> 0-d tensor`` replaces the
    CUDA op in CPU unit tests of this wrapper (gloo); the product path never passes it.
    """
    if lamda == 0.0:
        return 0.0
    world = _world(group)
    B_local = pred_local.shape[0]
    if global_batch is None:
        global_batch = B_local * world
    if loss_fn is None:
        _tl._check_inputs(pred_local, true_local, feat_d)
    if interp != 0:
        size = (interp,) * 2
        pred_local = F.interpolate(pred_local, size=size, mode="bilinear", align_corners=True)
        true_local = F.interpolate(true_local, size=size, mode="bilinear", align_corners=True)
    B, C, H, W = pred_local.shape
    pred, truth = pred_local.contiguous(), true_local.detach().contiguous()
    if global_batch == 1:  # the .squeeze() quirk depends on the GLOBAL batch, not on the shard
        if C == 1:
            raise ValueError("B == C == 1: the reference crashes here")
        pred, truth = pred.reshape(C, 1, H, W), truth.reshape(C, 1, H, W)
        global_batch = C
    if loss_fn is None:
        part = _tl._TopoLossFn.apply(pred, truth, lamda, feat_d, loss_q, loss_r, int(global_batch))
    else:
        part = loss_fn(pred, truth, lamda, feat_d, loss_q, loss_r, int(global_batch))
    if world == 1:
        return part
    return _AllReduceSum.apply(part, group)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = False) -> None:
    """Bucketed all-reduce of parameter gradients (what DDP does for ``model.mask_decoder``);
    ``average=False`` suits ``topo_loss_sharded`` whose local gradients already carry 1/B_global."""
    world = _world(group)
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def training_step(model, inputs: dict, gt_masks: torch.Tensor, optimizer, seg_loss: Callable, *,
                  topological: bool = True, lamda: float = 0.1, feat_d: int = 1, interp: int = 50,
                  global_batch: Optional[int] = None, group=None, decoder_params=None,
                  loss_fn: Optional[Callable] = None) -> torch.Tensor:
    """One data-parallel training step: the body of the reference's loop, training_utils.py:55-68,
    on this rank's shard of the batch.

    ``model(**inputs, multimask_output=False).pred_masks`` is ``[B, Nmax, 1, 256, 256]``; masks are
    resampled to 1024^2, cropped to ``reshaped_input_sizes`` and resampled to ``original_sizes``
    (:57-59), then ``seg_loss`` (+ the sharded topological loss, :63-64), backward, gradient
    all-reduce over the mask-decoder parameters, ``optimizer.step()``.  Returns the global loss.
    """
    world = _world(group)
    optimizer.zero_grad()
    outputs = model(**inputs, multimask_output=False)
    masks = F.interpolate(outputs.pred_masks.squeeze(2), (1024, 1024), mode="bilinear", align_corners=False)
    rs, osz = inputs["reshaped_input_sizes"], inputs["original_sizes"]
    masks = masks[..., : int(rs[0, 0]), : int(rs[0, 1])]
    masks = F.interpolate(masks, (int(osz[0, 0]), int(osz[0, 1])), mode="bilinear", align_corners=False)
    B_local = masks.shape[0]
    gb = global_batch if global_batch is not None else B_local * world
    # seg_loss is a mean over the local shard: weight it so that the sum over ranks is the global mean
    loss = seg_loss(masks, gt_masks) * (B_local / gb)
    if world > 1:
        loss = _AllReduceSum.apply(loss, group)
    if topological:
        loss = loss + topo_loss_sharded(torch.sigmoid(masks.float()), gt_masks.float(), lamda, feat_d=feat_d,
                                        interp=interp, global_batch=gb, group=group, loss_fn=loss_fn)
    loss.backward()
    params = list(decoder_params) if decoder_params is not None else [p for p in model.parameters() if p.requires_grad]
    allreduce_gradients(params, group=group, average=False)
    optimizer.step()
    return loss.detach()
