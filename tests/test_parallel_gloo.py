"""Data-parallel wrapper on CPU: world_size-2 gloo processes, with the CPU oracle standing in for
the CUDA op (test infrastructure only), must reproduce the unsharded loss and gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


class _OracleLoss(torch.autograd.Function):
    """Stand-in with tl_forward / tl_backward semantics (incl. B_global) built on the oracle."""

    @staticmethod
    def forward(ctx, pred, truth, lamda, feat_d, q, loss_r, gb):
        B = pred.shape[0]
        # kernel semantics: the maps are the geometry they come in (the wrapper has already exchanged H and W of non-square maps)
        loss, grad, _ = oracle.topo_loss(pred.detach().contiguous().numpy(), truth.contiguous().numpy(), lamda, feat_d=feat_d,
                                         loss_q=q, loss_r=loss_r, reference_shape_order=False)
        ctx.grad = torch.tensor(grad) * (B / gb)
        return torch.tensor(loss * B / gb)

    @staticmethod
    def backward(ctx, g):
        return ctx.grad * g, None, None, None, None, None, None


def _oracle_fn(pred, truth, lamda, feat_d, q, loss_r, gb):
    return _OracleLoss.apply(pred, truth, lamda, feat_d, q, loss_r, gb)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pred, truth, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dilabhelmholtzoct_b200.parallel import allreduce_gradients, shard_batch, topo_loss_sharded
    sl = shard_batch(pred.shape[0], rank, world)
    w = torch.nn.Parameter(torch.ones(1))
    p = pred[sl].clone().requires_grad_(True)
    loss = topo_loss_sharded(p * w, truth[sl], 0.1, feat_d=1, global_batch=pred.shape[0], loss_fn=_oracle_fn)
    loss.backward()
    allreduce_gradients([w])
    out[rank] = (float(loss), p.grad.numpy().copy(), float(w.grad))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_loss_equals_unsharded():
    rng = np.random.default_rng(3)
    pred = torch.tensor(rng.random((4, 3, 16, 16)).astype(np.float32))
    truth = torch.tensor((rng.random((4, 3, 16, 16)) < 0.4).astype(np.float32))
    want, wgrad, _ = oracle.topo_loss(pred.numpy(), truth.numpy(), 0.1, feat_d=1)
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), pred, truth, out), nprocs=world, join=True)
    from dilabhelmholtzoct_b200.parallel import shard_batch
    for r in range(world):
        loss, grad, wg = out[r]
        assert abs(loss - want) <= 1e-6 * abs(want)           # every rank sees the global loss
        assert np.allclose(grad, wgrad[shard_batch(4, r, world)], rtol=1e-6, atol=1e-9)
        assert abs(wg - float((wgrad * pred.numpy()).sum())) <= 1e-5 * abs(wg)  # all-reduced parameter gradient


class _Scale(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor([1.0, 0.5, 2.0]).view(1, 3, 1, 1))

    def forward(self, x):
        return x * self.w


def _ddp_worker(rank, world, port, pred, truth, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch.nn.parallel import DistributedDataParallel as DDP
    from dilabhelmholtzoct_b200.parallel import reduce_scalar_async, shard_batch, topo_loss_sharded
    sl = shard_batch(pred.shape[0], rank, world)
    model = DDP(_Scale())
    # DDP AVERAGES the gradients: the sharded loss must be told (grad_reduce="mean"), otherwise the topological
    # term would be scaled by 1 / world
    loss = topo_loss_sharded(model(pred[sl]), truth[sl], 0.1, feat_d=1, global_batch=pred.shape[0], grad_reduce="mean",
                             loss_fn=_oracle_fn)
    loss.backward()
    g_mean = model.module.w.grad.clone()
    # the share-of-the-global-loss form: no collective inside, the scalar is summed asynchronously
    model.zero_grad()
    part = topo_loss_sharded(model(pred[sl]), truth[sl], 0.1, feat_d=1, global_batch=pred.shape[0], grad_reduce="mean",
                             reduce="local", loss_fn=_oracle_fn)
    handle = reduce_scalar_async(part)
    part.backward()
    out[rank] = (float(loss), g_mean.numpy().copy(), float(handle.result()), model.module.w.grad.numpy().copy())
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_ddp_with_mean_reduction_equals_single_process():
    rng = np.random.default_rng(5)
    pred = torch.tensor(rng.random((4, 3, 14, 14)).astype(np.float32))
    truth = torch.tensor((rng.random((4, 3, 14, 14)) < 0.4).astype(np.float32))
    single = _Scale()
    want = _oracle_fn(single(pred), truth, 0.1, 1, 2, False, 4)
    want.backward()
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_ddp_worker, args=(world, _free_port(), pred, truth, out), nprocs=world, join=True)
    for r in range(world):
        loss, g_mean, loss_async, g_local = out[r]
        assert abs(loss - float(want)) <= 1e-6 * abs(float(want))
        assert abs(loss_async - float(want)) <= 1e-6 * abs(float(want))
        assert np.allclose(g_mean, single.w.grad.numpy(), rtol=1e-5, atol=1e-8)   # DDP's average == single-process gradient
        assert np.allclose(g_local, single.w.grad.numpy(), rtol=1e-5, atol=1e-8)


def test_shard_batch_covers_the_batch():
    from dilabhelmholtzoct_b200.parallel import shard_batch
    for n in (1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            idx = [i for r in range(world) for i in range(n)[shard_batch(n, r, world)]]
            assert idx == list(range(n))


def test_single_process_wrapper_matches_oracle_and_handles_global_batch_one():
    from dilabhelmholtzoct_b200.parallel import topo_loss_sharded
    rng = np.random.default_rng(4)
    pred = torch.tensor(rng.random((1, 3, 12, 12)).astype(np.float32), requires_grad=True)
    truth = torch.tensor((rng.random((1, 3, 12, 12)) < 0.4).astype(np.float32))
    loss = topo_loss_sharded(pred, truth, 0.1, feat_d=1, loss_fn=_oracle_fn)
    # B == 1: .squeeze() turns every channel into its own image
    want, _, _ = oracle.topo_loss(pred.detach().permute(1, 0, 2, 3).contiguous().numpy(),
                                  truth.permute(1, 0, 2, 3).contiguous().numpy(), 0.1, feat_d=1)
    assert abs(float(loss) - want) <= 1e-6 * abs(want)
    assert topo_loss_sharded(pred, truth, 0.0) == 0.0


def test_wrapper_reads_non_square_maps_like_the_reference():
    """H != W: the wrapper hands the op the flat buffers as W rows of H pixels; the gradient comes
    back in the caller's layout"""
    from dilabhelmholtzoct_b200.parallel import topo_loss_sharded
    rng = np.random.default_rng(6)
    pred = torch.tensor(rng.random((2, 2, 8, 14)).astype(np.float32), requires_grad=True)
    truth = torch.tensor((rng.random((2, 2, 8, 14)) < 0.4).astype(np.float32))
    loss = topo_loss_sharded(pred, truth, 0.1, feat_d=1, loss_fn=_oracle_fn)
    loss.backward()
    want, wgrad, _ = oracle.topo_loss(pred.detach().numpy(), truth.numpy(), 0.1, feat_d=1)  # the reference's reading
    assert abs(float(loss) - want) <= 1e-6 * abs(want)
    assert pred.grad.shape == pred.shape and np.allclose(pred.grad.numpy(), wgrad, rtol=1e-6, atol=1e-9)
    plain, _, _ = oracle.topo_loss(pred.detach().numpy(), truth.numpy(), 0.1, feat_d=1, reference_shape_order=False)
    assert abs(plain - want) > 1e-6 * abs(want)  # ... which is not the loss of the maps read as they look
