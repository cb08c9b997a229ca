"""Data-parallel wrapper around the reference's training step for the topological loss.

The reference trains in one process on one device (``/root/reference/octsam/models/training_utils.py:27-80``,
``device = "cuda" if ... else "cpu"`` at :33).  Here the batch axis is sharded over one process per
GPU (``torch.distributed``, NCCL over NVLink): every (image, class) map is independent through
persistence and matching, and the only coupling is ``mean_b`` over images
(``/root/reference/octsam/models/topological_loss.py:85``).  So the data path needs no collective; the
only exchanges are

* an all-reduce of the scalar loss -- needed for REPORTING only: the gradient of the global mean with
  respect to a rank's maps depends on that rank's maps alone, so ``training_step`` starts it on a side
  stream (``reduce_scalar_async``) and lets it overlap the backward pass;
* the all-reduce of the mask-decoder gradients (the encoders are frozen, training_utils.py:277-279):
  ``allreduce_gradients`` (one flat bucket, SUM) or ``DistributedDataParallel(model.mask_decoder)`` (MEAN).

Gradient reduction and the loss scale must agree -- ``topo_loss_sharded(..., grad_reduce=...)``:

* ``"sum"`` (default; ``allreduce_gradients(average=False)``): local gradients carry ``1 / B_global``;
* ``"mean"`` (DDP averages over ranks): local gradients carry ``world / B_global``, so that DDP's average is
  again the gradient of the global mean.  (Under DDP, plain ``topo_loss`` on the local shard is equivalent
  when every rank holds the same number of images.)
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


_SIDE_STREAMS = {}


class AsyncScalar:
    """Sum of a detached scalar over the ranks, started NOW without stalling the caller's stream: on CUDA the
    collective runs behind a side stream (the current stream does not wait for it), on CPU (gloo) it is an
    ``async_op``.  ``result()`` makes the current stream (or the host, for gloo) wait and returns the tensor."""

    def __init__(self, x: torch.Tensor, group=None):
        self.world = _world(group)
        x = x.detach()
        self._work = self._event = None
        if self.world == 1:
            self.y = x
        elif x.is_cuda:
            cur = torch.cuda.current_stream(x.device)
            side = _SIDE_STREAMS.get(x.device.index)
            if side is None:
                side = _SIDE_STREAMS[x.device.index] = torch.cuda.Stream(device=x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.y = x.clone()
                dist.all_reduce(self.y, op=dist.ReduceOp.SUM, group=group)
                self._event = torch.cuda.Event()
                self._event.record(side)
            x.record_stream(side)
        else:
            self.y = x.clone()
            self._work = dist.all_reduce(self.y, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def result(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._event is not None:
            torch.cuda.current_stream(self.y.device).wait_event(self._event)
            self.y.record_stream(torch.cuda.current_stream(self.y.device))
            self._event = None
        return self.y


def reduce_scalar_async(x: torch.Tensor, group=None) -> AsyncScalar:
    return AsyncScalar(x, group)


def shard_batch(n_images: int, rank: int, world: int) -> slice:
    """Contiguous slice of the batch axis owned by ``rank`` (whole images: all classes of an image
    stay on one GPU because W_b couples the classes of one image)."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def topo_loss_sharded(pred_local, true_local, lamda, interp=0, feat_d=2, loss_q=2, loss_r=False, *,
                      global_batch: Optional[int] = None, group=None, grad_reduce: str = "sum",
                      reduce: str = "sync", from_logits: bool = False,
                      loss_fn: Optional[Callable] = None):
    """``topo_loss`` with the batch axis sharded over the ranks of ``group``.

    ``reduce="sync"`` returns the GLOBAL loss ``lamda * mean_b W_b`` (identical on every rank);
    ``reduce="local"`` returns this rank's additive share of it (no collective: sum the shares yourself, e.g.
    with ``reduce_scalar_async`` so that the collective overlaps the backward pass).  Either way the backward
    gives each rank ``d loss_global / d pred_local`` -- the mean's ``1 / B_global`` is applied inside the
    kernel, no gradient is communicated -- scaled for the way parameter gradients are reduced afterwards:
    ``grad_reduce="sum"`` (``allreduce_gradients(average=False)``) or ``"mean"`` (DistributedDataParallel).
    ``global_batch`` defaults to ``B_local * world_size``.  ``from_logits=True`` takes the decoder's logits and
    the raw ground truth and runs the fused sigmoid + resample of ``topo_loss_from_logits`` (CUDA only).

    ``loss_fn(pred, true, lamda, feat_d, loss_q, loss_r, global_batch) -> 0-d tensor`` replaces the
    CUDA op in CPU unit tests of this wrapper (gloo); the product path never passes it.
    """
    if lamda == 0.0:
        return 0.0
    if grad_reduce not in ("sum", "mean") or reduce not in ("sync", "local"):
        raise ValueError("grad_reduce must be 'sum' or 'mean', reduce 'sync' or 'local'")
    world = _world(group)
    B_local = pred_local.shape[0]
    if global_batch is None:
        global_batch = B_local * world
    if from_logits:
        if loss_fn is not None:
            pred_local, true_local = torch.sigmoid(pred_local.float()), true_local.float()
        else:
            H, W = pred_local.shape[-2:]
            if interp == 0 and H != W:  # nothing to resample (non-square maps: see topological_loss._canonical)
                pred_local, true_local = torch.sigmoid(pred_local.float()), true_local.detach().float()
            else:
                S = int(interp) if interp != 0 else H
                pred_local = _tl.resample(pred_local.float(), S, sigmoid=True)
                true_local = _tl.resample(true_local.detach().float(), S, sigmoid=False)
            interp = 0
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
    if H != W:  # the reference reads the same flat buffer as W rows of H pixels (topological_loss._canonical)
        pred, truth = pred.view(*pred.shape[:2], W, H), truth.view(*truth.shape[:2], W, H)
    # DDP divides the summed gradients by world: pre-multiply by world (mean over B_global / world images)
    scale = world if (grad_reduce == "mean" and world > 1) else 1
    if global_batch % scale:
        raise ValueError("grad_reduce='mean' needs a global batch that is a multiple of the world size")
    gb_eff = global_batch // scale
    if loss_fn is None:
        part = _tl._TopoLossFn.apply(pred, truth, lamda, feat_d, loss_q, loss_r, int(gb_eff))
    else:
        part = loss_fn(pred, truth, lamda, feat_d, loss_q, loss_r, int(gb_eff))
    if scale != 1:  # value: this rank's share of the global mean; gradient: unchanged (scaled for the DDP average)
        part = part + (part.detach() / scale - part.detach())
    if world == 1 or reduce == "local":
        return part
    if scale != 1:
        total = part.detach().clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return part + (total - part.detach())
    return _AllReduceSum.apply(part, group)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = False) -> int:
    """Bucketed all-reduce of parameter gradients (what DDP does for ``model.mask_decoder``);
    ``average=False`` suits ``topo_loss_sharded(grad_reduce="sum")`` whose local gradients already carry
    1/B_global.  Returns the number of bytes put through the collective."""
    world = _world(group)
    if world == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()


def training_step(model, inputs: dict, gt_masks: torch.Tensor, optimizer, seg_loss: Optional[Callable] = None, *,
                  topological: bool = True, lamda: float = 0.1, feat_d: int = 1, interp: int = 50,
                  global_batch: Optional[int] = None, group=None, decoder_params=None,
                  loss_fn: Optional[Callable] = None, stats: Optional[dict] = None) -> torch.Tensor:
    """One data-parallel training step: the body of the reference's loop, training_utils.py:55-68,
    on this rank's shard of the batch.

    ``model(**inputs, multimask_output=False).pred_masks`` is ``[B, Nmax, 1, 256, 256]``.  On CUDA the step
    uses this package's kernels for everything around the model: ``postprocess_masks`` (the 1024^2 resample,
    crop and resample to ``original_sizes`` of :57-59 as one gather), ``dice_ce_loss`` (``seg_loss=None``: the
    reference's ``monai.losses.DiceCELoss(sigmoid=True)``, :32 / :62, reading the masks once) and the fused
    sigmoid + down-sample + topological loss of :63-64.  On CPU (unit tests: ``loss_fn`` given) the same lines
    run in PyTorch.  Then backward, SUM all-reduce of the mask-decoder gradients, ``optimizer.step()``.
    The scalar loss is all-reduced on a side stream while the backward runs; returns the global loss.
    """
    world = _world(group)
    optimizer.zero_grad()
    outputs = model(**inputs, multimask_output=False)
    rs, osz = inputs["reshaped_input_sizes"], inputs["original_sizes"]
    low = outputs.pred_masks.squeeze(2)
    on_gpu = low.is_cuda and loss_fn is None
    if on_gpu:
        masks = _tl.postprocess_masks(low.float(), (int(rs[0, 0]), int(rs[0, 1])), (int(osz[0, 0]), int(osz[0, 1])))
    else:
        masks = F.interpolate(low, (1024, 1024), mode="bilinear", align_corners=False)
        masks = masks[..., : int(rs[0, 0]), : int(rs[0, 1])]
        masks = F.interpolate(masks, (int(osz[0, 0]), int(osz[0, 1])), mode="bilinear", align_corners=False)
    B_local = masks.shape[0]
    gb = global_batch if global_batch is not None else B_local * world
    if seg_loss is None:
        seg_loss = _tl.dice_ce_loss
    # seg_loss is a mean over the local shard: weight it so that the sum over ranks is the global mean
    local = seg_loss(masks, gt_masks) * (B_local / gb)
    if topological:
        local = local + topo_loss_sharded(masks, gt_masks, lamda, feat_d=feat_d, interp=interp, global_batch=gb, group=group,
                                          reduce="local", from_logits=True, loss_fn=loss_fn)
    total = reduce_scalar_async(local, group)  # overlaps the backward pass; only the REPORTED loss needs it
    local.backward()
    params = list(decoder_params) if decoder_params is not None else [p for p in model.parameters() if p.requires_grad]
    nbytes = allreduce_gradients(params, group=group, average=False)
    optimizer.step()
    if stats is not None:
        stats["grad_allreduce_bytes"] = nbytes
    return total.result()
