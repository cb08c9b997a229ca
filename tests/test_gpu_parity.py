"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.
Bit-exact for persistence pairs / critical pixels; loss and gradient within 1e-5 relative."""
import numpy as np
import pytest
import torch

import oracle
from tests.kats import KATS

pytestmark = pytest.mark.gpu

REL = 1e-5  # north_star: loss values and gradients agree within 1e-5 relative in fp32


def _gpu_pairs(maps, dim):
    import dilabhelmholtzoct_b200 as tlb
    out = tlb.persistence_pairs(torch.as_tensor(maps, dtype=torch.float32, device="cuda"), dim)
    return [p.cpu().numpy() for p in out]


def _assert_same_pairs(maps, dim):
    got = _gpu_pairs(maps, dim)
    for k, f in enumerate(maps):
        want = oracle.cubical_pairs(f, dim)
        assert got[k].shape == want.shape, (k, dim, got[k].shape, want.shape)
        assert np.array_equal(got[k], want), (k, dim)


@pytest.mark.parametrize("name", sorted(KATS))
def test_kats(name):
    img, h0, h1, ess = KATS[name]
    f = np.array(img, dtype=np.float32)
    g0 = _gpu_pairs(f[None], 0)[0]
    g1 = _gpu_pairs(f[None], 1)[0]
    assert sorted(map(tuple, g0[:-1].tolist())) == sorted(h0)
    assert tuple(g0[-1].tolist()) == ess
    assert sorted(map(tuple, g1.tolist())) == sorted(h1)


@pytest.mark.parametrize("dim", [0, 1])
@pytest.mark.parametrize("size", [2, 3, 5, 8, 17, 32, 50, 64])
def test_pairs_random(dim, size):
    rng = np.random.default_rng(100 + size)
    maps = rng.random((12, size, size)).astype(np.float32)
    _assert_same_pairs(maps, dim)


@pytest.mark.parametrize("dim", [0, 1])
@pytest.mark.parametrize("levels", [2, 4, 32, 1024])
def test_pairs_ties(dim, levels):
    """Tie-heavy maps: critical-pixel indices depend on gudhi's (value, dim, position) cell order."""
    rng = np.random.default_rng(7 + levels)
    maps = (rng.integers(0, levels, (16, 24, 24)) / levels).astype(np.float32)
    maps = np.concatenate([maps, np.zeros((1, 24, 24), np.float32), np.ones((1, 24, 24), np.float32)])
    _assert_same_pairs(maps, dim)


@pytest.mark.parametrize("dim", [0, 1])
def test_pairs_256_synthetic(dim):
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(1, 256, 256, seed=4321)
    maps = torch.cat([pred[0], truth[0]]).numpy()
    _assert_same_pairs(maps, dim)


@pytest.mark.parametrize("dim", [0, 1])
def test_pairs_signed_and_negative_zero(dim):
    rng = np.random.default_rng(3)
    maps = (rng.integers(-3, 4, (8, 16, 16))).astype(np.float32)
    maps[maps == 0] = np.where(rng.random((maps == 0).sum()) < 0.5, -0.0, 0.0)
    _assert_same_pairs(maps, dim)


@pytest.mark.parametrize("dim", [0, 1])
def test_pairs_banded_maps(dim):
    """> 65535 nodes: the shared-memory kernel works through the map in bands of whole rows."""
    rng = np.random.default_rng(5)
    maps = rng.random((3, 300, 300)).astype(np.float32)
    maps[2] = np.round(maps[2] * 8) / 8
    _assert_same_pairs(maps, dim)


@pytest.mark.parametrize("dim", [0, 1])
def test_pairs_global_memory_kernel(dim):
    """The global-memory kernel (maps wider than the shared-memory kernel's row limit), forced here."""
    from dilabhelmholtzoct_b200 import _lib
    L = _lib.lib()
    L.tl_set_option(_lib.OPT_FORCE_GLOBAL_KERNEL, 1)
    try:
        rng = np.random.default_rng(8)
        maps = rng.random((3, 70, 70)).astype(np.float32)
        maps[1] = np.round(maps[1] * 4) / 4
        maps[2] = (maps[2] > 0.5).astype(np.float32)
        _assert_same_pairs(maps, dim)
        from dilabhelmholtzoct_b200.synthetic import make_batch
        pred, truth = make_batch(2, 40, 40, seed=8, n_classes=3)
        _check_loss(pred, truth, 0.1, dim)
    finally:
        L.tl_set_option(_lib.OPT_FORCE_GLOBAL_KERNEL, 0)


def test_pairs_many_basins_table_spills_to_global():
    """A checkerboard has ~N/2 basins: the triplet table no longer fits shared memory."""
    rng = np.random.default_rng(6)
    yy, xx = np.mgrid[0:256, 0:256]
    board = ((yy + xx) % 2).astype(np.float32)
    maps = np.stack([board + 0.25 * rng.random((256, 256)).astype(np.float32), board]).astype(np.float32)
    _assert_same_pairs(maps, 1)


def test_pairs_repeatable():
    rng = np.random.default_rng(11)
    maps = rng.random((6, 96, 96)).astype(np.float32)
    a = _gpu_pairs(maps, 1)
    for _ in range(3):
        b = _gpu_pairs(maps, 1)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("q", [1.0, 2.0, 3.0])
def test_wasserstein_vs_oracle(q):
    import dilabhelmholtzoct_b200 as tlb
    rng = np.random.default_rng(int(q) * 13)
    D1, D2 = [], []
    for t in range(40):
        n, m = int(rng.integers(0, 60)), int(rng.integers(0, 12))
        if t % 7 == 0:
            n, m = m, n
        b = rng.random(n).astype(np.float32)
        D1.append(np.stack([b, b + rng.random(n).astype(np.float32)], 1).reshape(-1, 2))
        b = rng.random(m).astype(np.float32)
        d2 = np.stack([b, b + rng.random(m).astype(np.float32)], 1).reshape(-1, 2)
        if t % 5 == 0 and m:
            d2[:] = np.array([0.0, 1.0], np.float32)  # binary ground truth: identical points
        D2.append(d2)
    cost, match = tlb.wasserstein_cost([torch.tensor(d, device="cuda") for d in D1],
                                       [torch.tensor(d, device="cuda") for d in D2], q)
    cost = cost.cpu().numpy()
    for k in range(len(D1)):
        want, _ = oracle.wasserstein(D1[k], D2[k], q)
        assert abs(cost[k] - want) <= 1e-9 + 1e-7 * abs(want), (k, cost[k], want)
        mk = match[k].cpu().numpy()
        used = mk[mk >= 0]
        assert len(set(used.tolist())) == len(used) and (used < len(D2[k])).all()


def _loss_and_grad(pred, truth, lamda, **kw):
    import dilabhelmholtzoct_b200 as tlb
    p = pred.clone().cuda().requires_grad_(True)
    loss = tlb.topo_loss(p, truth.cuda(), lamda, **kw)
    loss.backward()
    return float(loss), p.grad.cpu().numpy()


def _check_loss(pred, truth, lamda, feat_d, q=2, loss_r=False):
    loss, grad = _loss_and_grad(pred, truth, lamda, feat_d=feat_d, loss_q=q, loss_r=loss_r)
    want, wgrad, _ = oracle.topo_loss(pred.numpy(), truth.numpy(), lamda, feat_d=feat_d, loss_q=q, loss_r=loss_r)
    assert abs(loss - want) <= REL * abs(want) + 1e-12, (loss, want)
    scale = np.abs(wgrad).max()
    assert np.array_equal(grad != 0, wgrad != 0), "critical pixels differ"
    assert np.abs(grad - wgrad).max() <= REL * scale + 1e-12, (np.abs(grad - wgrad).max(), scale)


@pytest.mark.parametrize("feat_d", [0, 1])
@pytest.mark.parametrize("q", [1, 2])
def test_loss_small(feat_d, q):
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(3, 48, 48, seed=77, n_classes=5)
    _check_loss(pred, truth, 0.1, feat_d, q=q)


def test_loss_regulariser():
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(2, 40, 40, seed=78, n_classes=4)
    _check_loss(pred, truth, 0.25, 1, q=2, loss_r=True)


def test_loss_nonbinary_truth():
    """interp-style ground truth (fractional values -> several truth points per map)."""
    rng = torch.Generator().manual_seed(5)
    pred = torch.rand((2, 3, 50, 50), generator=rng)
    truth = torch.nn.functional.avg_pool2d(torch.rand((2, 3, 100, 100), generator=rng), 2)
    _check_loss(pred, truth, 0.1, 1)
    _check_loss(pred, truth, 0.1, 0)


def test_loss_256x14():
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(2, 256, 256, seed=1234)
    _check_loss(pred, truth, 0.1, 1)


def test_interp_and_upstream_grad():
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(2, 96, 96, seed=9, n_classes=4)
    logits = torch.logit(pred.clamp(1e-4, 1 - 1e-4)).cuda().requires_grad_(True)
    loss = 3.0 * tlb.topo_loss(torch.sigmoid(logits), truth.cuda(), 0.1, feat_d=1, interp=50)
    loss.backward()
    # oracle on the resampled maps, chained through interpolate + sigmoid by autograd
    lg = logits.detach().cpu().requires_grad_(True)
    ps = torch.nn.functional.interpolate(torch.sigmoid(lg), size=(50, 50), mode="bilinear", align_corners=True)
    ts = torch.nn.functional.interpolate(truth, size=(50, 50), mode="bilinear", align_corners=True)
    want, wgrad, _ = oracle.topo_loss(ps.detach().numpy(), ts.numpy(), 0.1, feat_d=1)
    ps.backward(torch.tensor(wgrad) * 3.0)
    assert abs(float(loss) - 3.0 * want) <= REL * abs(3.0 * want)
    g, w = logits.grad.cpu().numpy(), lg.grad.numpy()
    assert np.abs(g - w).max() <= 1e-4 * np.abs(w).max()


def test_squeeze_quirk_batch_of_one():
    """B == 1: .squeeze() makes every channel its own image."""
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(1, 32, 32, seed=3, n_classes=4)
    loss, grad = _loss_and_grad(pred, truth, 0.1, feat_d=1)
    want, wgrad, _ = oracle.topo_loss(pred.permute(1, 0, 2, 3).contiguous().numpy(),
                                      truth.permute(1, 0, 2, 3).contiguous().numpy(), 0.1, feat_d=1)
    assert abs(loss - want) <= REL * abs(want)
    assert np.abs(grad - wgrad.transpose(1, 0, 2, 3)).max() <= REL * np.abs(wgrad).max()


def test_errors_and_early_out():
    import dilabhelmholtzoct_b200 as tlb
    x = torch.rand((2, 2, 8, 8), device="cuda")
    assert tlb.topo_loss(x, x, 0.0) == 0.0
    with pytest.raises(ValueError):
        tlb.topo_loss(x, x, 0.1)  # default feat_d=2 is invalid on 2-D maps
    with pytest.raises(ValueError):
        tlb.topo_loss(x.cpu(), x.cpu(), 0.1, feat_d=1)  # no CPU fallback
    with pytest.raises(ValueError):
        tlb.topo_loss(torch.rand((2, 2, 8, 9), device="cuda"), torch.rand((2, 2, 9, 8), device="cuda"), 0.1, feat_d=1)  # shapes differ
    with pytest.raises(ValueError):
        tlb.topo_loss(torch.rand((2, 2, 8, 1), device="cuda"), torch.rand((2, 2, 8, 1), device="cuda"), 0.1, feat_d=1)  # squeezed away
    with pytest.raises(ValueError):
        tlb.topo_loss(x[:1, :1], x[:1, :1], 0.1, feat_d=1)


def test_zero_cost_gives_nan_grad():
    """S_b == 0 -> pow(1/q) backward is inf * 0 = NaN in the reference's autograd."""
    import dilabhelmholtzoct_b200 as tlb
    rng = np.random.default_rng(0)
    f = torch.tensor(rng.random((2, 2, 12, 12)).astype(np.float32), device="cuda")
    p = f.clone().requires_grad_(True)
    loss = tlb.topo_loss(p, f, 0.1, feat_d=1)
    loss.backward()
    assert float(loss) == 0.0
    g = p.grad
    assert torch.isnan(g).any() and not torch.isnan(g).all()


def test_golden_fixture_on_gpu():
    import json, os
    data = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pairs_small.json")))
    for case in data["pairs"]:
        f = np.array(case["image"], dtype=np.float32)
        assert _gpu_pairs(f[None], 0)[0].tolist() == case["h0"]
        assert _gpu_pairs(f[None], 1)[0].tolist() == case["h1"]
    for case in data["losses"]:
        pred, truth = torch.tensor(case["pred"]), torch.tensor(case["truth"])
        loss, grad = _loss_and_grad(pred, truth, case["lamda"], feat_d=case["feat_d"], loss_q=case["q"])
        assert abs(loss - case["loss"]) <= REL * abs(case["loss"])
        g = np.array(case["grad"], np.float32)
        assert np.abs(grad - g).max() <= REL * np.abs(g).max() + 1e-12


def test_sharded_wrapper_matches_unsharded_on_gpu():
    """Shards run one after the other on one device (B_global passed to the kernel) must add up to
    the unsharded loss and reproduce its gradient."""
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200.parallel import shard_batch, topo_loss_sharded
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(4, 64, 64, seed=21, n_classes=3)
    full_loss, full_grad = _loss_and_grad(pred, truth, 0.1, feat_d=1)
    total, grads = 0.0, []
    for r in range(2):
        sl = shard_batch(4, r, 2)
        p = pred[sl].cuda().requires_grad_(True)
        part = topo_loss_sharded(p, truth[sl].cuda(), 0.1, feat_d=1, global_batch=4)
        part.backward()
        total += float(part)
        grads.append(p.grad.cpu().numpy())
    assert abs(total - full_loss) <= REL * abs(full_loss)
    assert np.abs(np.concatenate(grads) - full_grad).max() <= REL * np.abs(full_grad).max()


def test_forward_is_deterministic():
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(2, 128, 128, seed=5, n_classes=6)
    a = _loss_and_grad(pred, truth, 0.1, feat_d=1)
    for _ in range(3):
        b = _loss_and_grad(pred, truth, 0.1, feat_d=1)
        assert a[0] == b[0] and np.array_equal(a[1] != 0, b[1] != 0)


def test_host_api_matches_autograd_path():
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(6, 64, 64, seed=31, n_classes=3)
    want_loss, want_grad = _loss_and_grad(pred, truth, 0.1, feat_d=1)
    loss, grad = tlb.topo_loss_from_host(pred.pin_memory(), truth.pin_memory(), 0.1, feat_d=1, chunks=4)
    assert abs(float(loss) - want_loss) <= REL * abs(want_loss)
    assert np.abs(grad.cpu().numpy() - want_grad).max() <= REL * np.abs(want_grad).max()
    loss2, none = tlb.topo_loss_from_host(pred.pin_memory(), truth.pin_memory(), 0.1, feat_d=0, chunks=2, want_grad=False)
    want0, _ = _loss_and_grad(pred, truth, 0.1, feat_d=0)
    assert none is None and abs(float(loss2) - want0) <= REL * abs(want0)


def test_full_c2_batch_against_oracle():
    """BASELINE configs[1] at full size: fp32[64,14,256,256], feat_d=1 -- loss, gradient support and
    per-map pair counts against the (multi-threaded) oracle."""
    import os
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(64, 256, 256, seed=1234 + 2000)
    loss, grad = _loss_and_grad(pred, truth, 0.1, feat_d=1)
    want, wgrad, cnt = oracle.topo_loss(pred.numpy(), truth.numpy(), 0.1, feat_d=1, nthreads=os.cpu_count() or 1)
    assert abs(loss - want) <= REL * abs(want), (loss, want)
    assert np.array_equal(grad != 0, wgrad != 0)
    assert np.abs(grad - wgrad).max() <= REL * np.abs(wgrad).max()
    # size-independent property: every map's gradient entries sum to ~0 for diagonal-matched pairs
    # (each contributes (-g/2, +g/2)); matched pairs break this only on maps with truth points
    per_map = grad.reshape(64 * 14, -1).sum(1)
    no_truth = cnt[:, 1] == 0
    assert np.abs(per_map[no_truth]).max() <= 1e-3 * np.abs(grad).max()


@pytest.mark.parametrize("dim", [0, 1])
def test_pairs_1024(dim):
    """BASELINE configs[4] resolution: one smooth-ish and one tie-heavy 1024x1024 map."""
    rng = np.random.default_rng(9)
    a = rng.random((1024, 1024)).astype(np.float32)
    b = (np.round(rng.random((1024, 1024)) * 3) / 3).astype(np.float32)
    _assert_same_pairs(np.stack([a, b]), dim)


def test_loss_1024_small_batch():
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(1, 1024, 1024, seed=55, n_classes=3)
    pred = torch.cat([pred, pred.flip(-1)]); truth = torch.cat([truth, truth.flip(-1)])
    _check_loss(pred, truth, 0.1, 1)


def test_wasserstein_large_truth_diagrams():
    """Both diagrams large (hundreds of points): the assignment is no longer the easy tall-skinny case."""
    import dilabhelmholtzoct_b200 as tlb
    rng = np.random.default_rng(17)
    D1, D2 = [], []
    for n, m in ((300, 280), (150, 400), (257, 1), (1, 257), (0, 50), (50, 0)):
        b = rng.random(n).astype(np.float32)
        D1.append(np.stack([b, b + rng.random(n).astype(np.float32)], 1).reshape(-1, 2))
        b = rng.random(m).astype(np.float32)
        D2.append(np.stack([b, b + rng.random(m).astype(np.float32)], 1).reshape(-1, 2))
    cost, match = tlb.wasserstein_cost([torch.tensor(d, device="cuda") for d in D1],
                                       [torch.tensor(d, device="cuda") for d in D2], 2.0)
    cost = cost.cpu().numpy()
    for k in range(len(D1)):
        want, _ = oracle.wasserstein(D1[k], D2[k], 2.0)
        assert abs(cost[k] - want) <= 1e-9 + 1e-7 * abs(want), (k, cost[k], want)


def _blobs(rng, size, n_blobs, lo=0.0, hi=1.0):
    yy, xx = np.mgrid[0:size, 0:size]
    f = np.full((size, size), lo, np.float32)
    for _ in range(n_blobs):
        cy, cx = rng.integers(0, size, 2)
        ry, rx = rng.integers(1, max(2, size // 6), 2)
        f[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = hi
    for _ in range(n_blobs // 2):  # holes inside blobs, diagonal contacts
        cy, cx = rng.integers(0, size, 2)
        f[max(0, cy - 1):cy + 1, max(0, cx - 1):cx + 1] = lo
    return f


@pytest.mark.parametrize("size", [2, 7, 33, 64, 100, 256])
def test_pairs_two_valued_maps(size):
    """Two-valued maps take the bit-mask / run union-find path (csrc/ph_binary.cuh): random densities,
    blobs with holes, either value at pixel 0, arbitrary (lo, hi) incl. negative values and -0.0."""
    rng = np.random.default_rng(200 + size)
    maps = []
    for p in (0.05, 0.3, 0.5, 0.7, 0.95):
        maps.append((rng.random((size, size)) < p).astype(np.float32))
    maps.append(1.0 - maps[1])
    maps.append(maps[2] * 3.5 - 1.25)
    maps.append(np.where(maps[3] > 0, np.float32(-0.0), np.float32(-2.0)).astype(np.float32))
    maps.append(_blobs(rng, size, 6))
    maps.append(_blobs(rng, size, 12, lo=1.0, hi=0.0) if size > 2 else maps[0])
    maps.append(np.zeros((size, size), np.float32))
    stripes = np.zeros((size, size), np.float32); stripes[:, 1::2] = 1.0; stripes[0] = stripes[-1] = 0.0
    maps.append(stripes)
    maps = np.stack(maps).astype(np.float32)
    _assert_same_pairs(maps, 1)
    _assert_same_pairs(maps, 0)  # H0 has no two-valued short-cut: generic path on the same maps


def test_pairs_two_valued_1024_and_run_overflow():
    """1024x1024 masks (BASELINE configs[4] ground truth) and a 512x512 checkerboard whose 131 072 runs do
    not fit the run table: the kernel must fall back to the generic path and still be exact."""
    rng = np.random.default_rng(12)
    _assert_same_pairs(np.stack([_blobs(rng, 1024, 40), (rng.random((1024, 1024)) < 0.55).astype(np.float32)]), 1)
    yy, xx = np.mgrid[0:512, 0:512]
    _assert_same_pairs(((yy + xx) % 2).astype(np.float32)[None], 1)
    yy, xx = np.mgrid[0:256, 0:256]
    _assert_same_pairs(((yy + xx) % 2).astype(np.float32)[None], 1)  # 32 768 runs: still on the bit-mask path


def test_pairs_repeatable_at_c2_scale():
    """Lock-free merge + reductions at the headline size: 3 runs over 56 noisy 256x256 maps, identical pairs."""
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, _ = make_batch(4, 256, 256, seed=77)
    maps = pred.reshape(-1, 256, 256).numpy()
    a = _gpu_pairs(maps, 1)
    for _ in range(3):
        b = _gpu_pairs(maps, 1)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    want = oracle.cubical_pairs(maps[5], 1)
    assert np.array_equal(a[5], want)


def test_forward_under_inference_mode():
    """validate_model calls the loss under torch.inference_mode() (training_utils.py:356-378): same value, no graph."""
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(2, 64, 64, seed=41, n_classes=3)
    want, _, _ = oracle.topo_loss(pred.numpy(), truth.numpy(), 0.1, feat_d=1)
    with torch.inference_mode():
        loss = tlb.topo_loss(pred.cuda(), truth.cuda(), 0.1, feat_d=1, interp=0)
        fused = tlb.topo_loss_from_logits(torch.logit(pred.clamp(1e-4, 1 - 1e-4)).cuda(), truth.cuda(), 0.1, feat_d=1, interp=32)
    assert not loss.requires_grad and abs(float(loss) - want) <= REL * abs(want)
    assert torch.isfinite(fused) and not fused.requires_grad
    with torch.no_grad():
        assert float(tlb.topo_loss(pred.cuda(), truth.cuda(), 0.1, feat_d=1)) == float(loss)


def test_nan_and_inf_pixels():
    """A NaN pixel (sigmoid of a diverged logit) makes the cell order undefined: the kernels terminate, the loss is
    NaN and the status word says why; +/-inf are ordinary ordered values."""
    import dilabhelmholtzoct_b200 as tlb
    rng = np.random.default_rng(2)
    f = rng.random((2, 2, 32, 32)).astype(np.float32)
    g = f.copy(); g[0, 1, 5, 7] = np.inf; g[1, 0, 3, 3] = -np.inf
    _assert_same_pairs(g.reshape(-1, 32, 32), 1)
    _assert_same_pairs(g.reshape(-1, 32, 32), 0)
    bad = f.copy(); bad[1, 1, 9, 9] = np.nan
    for dim in (0, 1):
        p = torch.tensor(bad, device="cuda", requires_grad=True)
        loss = tlb.topo_loss(p, torch.tensor(f, device="cuda"), 0.1, feat_d=dim)
        assert torch.isnan(loss)
        with pytest.raises(RuntimeError, match="NaN"):
            tlb.check_status(sync=True)
    two = (f > 0.5).astype(np.float32); two[0, 0, 4, 4] = np.nan  # would-be two-valued map with a NaN
    loss = tlb.topo_loss(torch.tensor(two, device="cuda"), torch.tensor(f, device="cuda"), 0.1, feat_d=1)
    assert torch.isnan(loss)
    with pytest.raises(RuntimeError, match="NaN"):
        tlb.check_status(sync=True)
    tlb.check_status(sync=True)  # reported once


def test_arena_overflow_is_loud():
    """A state buffer too small for the pairs of the batch: nothing is written out of bounds, the loss is NaN,
    tl_status / check_status say 'arena', and a larger arena (set_arena_factor) fixes it."""
    import ctypes
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200 import _lib
    from dilabhelmholtzoct_b200.topological_loss import _buffers
    L = _lib.lib()
    rng = np.random.default_rng(4)
    B, C, S = 2, 3, 64
    pred = torch.tensor(rng.random((B, C, S, S)).astype(np.float32), device="cuda")
    truth = (pred > 0.7).float()
    dev = pred.device
    state, scratch = _buffers(B, C, S, S, 1, dev)
    ns, nc = ctypes.c_size_t(0), ctypes.c_size_t(0)
    L.tl_workspace_bytes(B, C, S, S, 1, ctypes.byref(ns), ctypes.byref(nc))
    per_map = min(2 * (S * S // 2 + 2), max(S * S // 5 + 64, 8192))  # arena records per map that tl_workspace_bytes asks for
    tiny = state[: ns.value - (B * C * per_map - 100) * 24]           # room for 100 records only
    guard = torch.full((4096,), 0x5A, dtype=torch.uint8, device=dev)
    buf = torch.cat([tiny, guard])
    loss = torch.zeros((), device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.tl_forward(pred.data_ptr(), truth.data_ptr(), B, C, S, S, 1, 2.0, 0.1, 0, 0, buf.data_ptr(), tiny.numel(),
                      scratch.data_ptr(), scratch.numel(), loss.data_ptr(), st)
    assert rc == 0
    status = ctypes.c_int(0)
    assert L.tl_status(buf.data_ptr(), ctypes.byref(status), st) == 0
    assert status.value & 1 and torch.isnan(loss)
    assert bool((buf[tiny.numel():] == 0x5A).all()), "records written past the arena"
    grad = torch.empty_like(pred)
    assert L.tl_backward(None, buf.data_ptr(), tiny.numel(), B, C, S, S, 1, 2.0, 0.1, 0, 0, grad.data_ptr(), st) == 0
    torch.cuda.synchronize()
    full = tlb.topo_loss(pred, truth, 0.1, feat_d=1)
    assert torch.isfinite(full)
    tlb.check_status(sync=True)


def test_host_api_uint8_truth_and_batch_of_one():
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(5, 64, 64, seed=32, n_classes=3)
    want_loss, want_grad = _loss_and_grad(pred, truth, 0.1, feat_d=1)
    loss, grad = tlb.topo_loss_from_host(pred.pin_memory(), truth.to(torch.uint8).pin_memory(), 0.1, feat_d=1, chunks=3)
    assert abs(float(loss) - want_loss) <= REL * abs(want_loss)
    assert np.abs(grad.cpu().numpy() - want_grad).max() <= REL * np.abs(want_grad).max()
    # ... and as bits (numpy.packbits order), widened by tl_unpack_mask_bits
    bits = tlb.pack_mask_bits(truth)
    assert bits.dtype == torch.uint8 and tuple(bits.shape) == (5, 3, 64, 8) and bits.is_pinned()
    loss_b, grad_b = tlb.topo_loss_from_host(pred.pin_memory(), bits, 0.1, feat_d=1, chunks=2, truth_packed=True)
    assert float(loss_b) == float(loss)
    assert float((grad_b - grad).abs().max()) <= 1e-6 * float(grad.abs().max())  # atomics may add in another order
    with pytest.raises(ValueError, match="truth_packed"):
        tlb.topo_loss_from_host(pred.pin_memory(), truth.to(torch.uint8).pin_memory(), 0.1, feat_d=1, truth_packed=True)
    p1, t1 = pred[:1].contiguous(), truth[:1].contiguous()
    w1, g1 = _loss_and_grad(p1, t1, 0.1, feat_d=1)
    loss1, grad1 = tlb.topo_loss_from_host(p1.pin_memory(), t1.pin_memory(), 0.1, feat_d=1)
    assert tuple(grad1.shape) == tuple(p1.shape) and abs(float(loss1) - w1) <= REL * abs(w1)
    assert np.abs(grad1.cpu().numpy() - g1).max() <= REL * np.abs(g1).max()
    with pytest.raises(ValueError, match="pinned"):
        tlb.topo_loss_from_host(pred, truth, 0.1, feat_d=1)
    with pytest.raises(ValueError, match="HOST"):
        tlb.topo_loss_from_host(pred.cuda(), truth.cuda(), 0.1, feat_d=1)


def test_double_backward_raises():
    import dilabhelmholtzoct_b200 as tlb
    p = torch.rand((2, 2, 16, 16), device="cuda", requires_grad=True)
    t = (torch.rand((2, 2, 16, 16), device="cuda") > 0.5).float()
    loss = tlb.topo_loss(p, t, 0.1, feat_d=1)
    (g,) = torch.autograd.grad(loss, p, create_graph=True)
    with pytest.raises(RuntimeError):
        g.sum().backward()


@pytest.mark.timeout(600)
def test_c5_scale_one_full_image():
    """BASELINE configs[4] at its real resolution: every map of one full 14-class 1024x1024 image (prediction and
    ground truth) bit-exact against the oracle, then loss and gradient of a 2-image batch within 1e-5."""
    import os
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(1, 1024, 1024, seed=1234 + 5000)
    _assert_same_pairs(torch.cat([pred[0], truth[0]]).numpy(), 1)
    p2, t2 = torch.cat([pred, pred.flip(-2)]), torch.cat([truth, truth.flip(-2)])
    loss, grad = _loss_and_grad(p2, t2, 0.1, feat_d=1)
    want, wgrad, _ = oracle.topo_loss(p2.numpy(), t2.numpy(), 0.1, feat_d=1, nthreads=os.cpu_count() or 1)
    assert abs(loss - want) <= REL * abs(want), (loss, want)
    assert np.array_equal(grad != 0, wgrad != 0)
    assert np.abs(grad - wgrad).max() <= REL * np.abs(wgrad).max()


def test_loss_many_maps_with_two_large_diagrams():
    """More maps with two large diagrams (> 8 points each) than the general matching kernel has scratch slots (16):
    every one of them must be matched exactly once (a loop-index bug once skipped / repeated some)."""
    rng = torch.Generator().manual_seed(9)
    pred = torch.rand((8, 6, 40, 40), generator=rng)
    truth = torch.nn.functional.avg_pool2d(torch.rand((8, 6, 80, 80), generator=rng), 2)
    for _ in range(2):
        _check_loss(pred, truth, 0.1, 1)


def test_fused_gradient_equals_separate_launch():
    """tl_forward_backward (gradient written in the tail of the persistence launch) against tl_forward + tl_backward,
    with images that take the small-R matching, images with maps on the heavy list (left to grad_kernel) and more
    images than SMs' worth of jobs; then the switch TL_OPT_NO_FUSED_GRAD, an upstream gradient != 1 and a second
    backward over a retained graph."""
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200 import _lib
    from dilabhelmholtzoct_b200.synthetic import make_batch
    from dilabhelmholtzoct_b200.topological_loss import _buffers
    L = _lib.lib()
    pred, truth = make_batch(24, 96, 96, seed=41, n_classes=7)
    rng = torch.Generator().manual_seed(3)
    soft = torch.nn.functional.avg_pool2d(torch.rand((24, 7, 192, 192), generator=rng), 2)
    truth[5], truth[17, 2] = soft[5], soft[17, 2]  # image 5: every map heavy; image 17: one heavy map
    pred, truth = pred.cuda(), truth.cuda()
    B, C, H, W = pred.shape
    st = torch.cuda.current_stream().cuda_stream

    def run(fused):
        state, scratch = _buffers(B, C, H, W, 1, pred.device)
        loss, grad = torch.zeros((), device="cuda"), torch.full_like(pred, 7.0)
        args = (pred.data_ptr(), truth.data_ptr(), B, C, H, W, 1, 2.0, 0.1, 0, 0, state.data_ptr(), state.numel(),
                scratch.data_ptr(), scratch.numel(), loss.data_ptr())
        if fused:
            assert L.tl_forward_backward(*args, grad.data_ptr(), st) == 0
        else:
            assert L.tl_forward(*args, st) == 0
            assert L.tl_backward(None, state.data_ptr(), state.numel(), B, C, H, W, 1, 2.0, 0.1, 0, 0, grad.data_ptr(), st) == 0
        torch.cuda.synchronize()
        return float(loss), grad.cpu().numpy()

    l0, g0 = run(False)
    for _ in range(3):
        l1, g1 = run(True)
        assert l1 == l0
        assert np.array_equal(g1 != 0, g0 != 0)
        assert np.abs(g1 - g0).max() <= 1e-6 * np.abs(g0).max()  # same terms; atomics may add them in another order
    L.tl_set_option(_lib.OPT_NO_FUSED_GRAD, 1)
    try:
        l2, g2 = run(True)
    finally:
        L.tl_set_option(_lib.OPT_NO_FUSED_GRAD, 0)
    assert l2 == l0 and np.abs(g2 - g0).max() <= 1e-6 * np.abs(g0).max()
    want, wgrad, _ = oracle.topo_loss(pred.cpu().numpy(), truth.cpu().numpy(), 0.1, feat_d=1)
    assert abs(l0 - want) <= REL * abs(want) and np.abs(g0 - wgrad).max() <= REL * np.abs(wgrad).max()
    # autograd: upstream gradient 2.5, then a second backward over the retained graph with upstream 1
    p = pred.clone().requires_grad_(True)
    loss = tlb.topo_loss(p, truth, 0.1, feat_d=1)
    (loss * 2.5).backward(retain_graph=True)
    assert np.abs(p.grad.cpu().numpy() - 2.5 * g0).max() <= 1e-6 * 2.5 * np.abs(g0).max()
    p.grad = None
    loss.backward()
    assert np.abs(p.grad.cpu().numpy() - g0).max() <= 1e-6 * np.abs(g0).max()


# ---- rectangular maps.  The C ABI takes the geometry gudhi sees (H rows of W pixels); for H != W the
#      reference reads the flat buffer as W rows of H pixels (torch_topological passes the shape un-reversed),
#      which the shim and the oracle's topo_loss both reproduce
@pytest.mark.parametrize("dim", [0, 1])
@pytest.mark.parametrize("shape", [(2, 7), (7, 2), (5, 9), (37, 53), (64, 200), (200, 64), (31, 255), (255, 32)])
def test_pairs_rectangular(dim, shape):
    rng = np.random.default_rng(7000 + shape[0] * 1000 + shape[1])
    maps = rng.random((4,) + shape).astype(np.float32)
    maps[1] = np.round(maps[1] * 6) / 6       # tie-heavy
    maps[2] = (maps[2] > 0.45).astype(np.float32)  # two-valued
    _assert_same_pairs(maps, dim)


@pytest.mark.parametrize("dim", [0, 1])
@pytest.mark.parametrize("shape", [(120, 600), (600, 120), (130, 510), (496, 512)])
def test_pairs_rectangular_multi_band(dim, shape):
    """more than 65536 pixels: column bands of a rectangular map (fast front end for W % 4 == 0, generic otherwise)"""
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(1, shape[0], shape[1], seed=4100 + shape[1], n_classes=5)
    rng = np.random.default_rng(shape[0])
    maps = np.stack([pred[0, 1].numpy(), truth[0, 3].numpy(), rng.random(shape).astype(np.float32)])
    _assert_same_pairs(maps, dim)


@pytest.mark.parametrize("shape", [(2, 3, 40, 56), (2, 3, 56, 40), (3, 1, 24, 40)])
@pytest.mark.parametrize("feat_d", [0, 1])
def test_loss_rectangular_follows_the_reference_shape_order(shape, feat_d):
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200.synthetic import make_batch
    B, C, H, W = shape
    pred, truth = make_batch(B, H, W, seed=91 + H, n_classes=C)
    _check_loss(pred, truth, 0.1, feat_d)
    # the same thing said the long way: the loss of the flat buffers read as W rows of H pixels
    want, wgrad, _ = oracle.topo_loss(pred.numpy().reshape(B, C, W, H), truth.numpy().reshape(B, C, W, H), 0.1, feat_d=feat_d,
                                      reference_shape_order=False)
    loss, grad = _loss_and_grad(pred, truth, 0.1, feat_d=feat_d)
    assert grad.shape == (B, C, H, W)
    assert abs(loss - want) <= REL * abs(want) + 1e-12
    assert np.abs(grad.reshape(B, C, W, H) - wgrad).max() <= REL * np.abs(wgrad).max() + 1e-12
    # pairs at the inner boundary: plain reading by default, the reference's reading on request
    f = pred[0, 0].numpy()
    got = tlb.persistence_pairs(pred[0, 0].cuda(), 1, reference_shape_order=True)[0].cpu().numpy()
    assert np.array_equal(got, oracle.cubical_pairs(f.reshape(W, H), 1))


def test_rectangular_host_api_and_logits_call():
    import dilabhelmholtzoct_b200 as tlb
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(3, 48, 64, seed=17, n_classes=4)
    want, wgrad, _ = oracle.topo_loss(pred.numpy(), truth.numpy(), 0.1, feat_d=1)
    loss, grad = tlb.topo_loss_from_host(pred.pin_memory(), tlb.pack_mask_bits(truth), 0.1, feat_d=1, truth_packed=True, chunks=2)
    assert abs(float(loss) - want) <= REL * abs(want)
    assert grad.shape == pred.shape and np.abs(grad.cpu().numpy() - wgrad).max() <= REL * np.abs(wgrad).max()
    logits = torch.logit(pred.clamp(1e-4, 1 - 1e-4)).cuda().requires_grad_(True)
    l2 = tlb.topo_loss_from_logits(logits, truth.cuda(), 0.1, feat_d=1)
    l2.backward()
    w2, _, _ = oracle.topo_loss(torch.sigmoid(logits.detach()).cpu().numpy(), truth.numpy(), 0.1, feat_d=1)
    assert abs(float(l2) - w2) <= REL * abs(w2) and logits.grad.shape == logits.shape


def test_rectangular_batch_of_one():
    """B == 1 (every channel its own image) and H != W together"""
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(1, 24, 40, seed=5, n_classes=4)
    loss, grad = _loss_and_grad(pred, truth, 0.1, feat_d=1)
    want, wgrad, _ = oracle.topo_loss(pred.permute(1, 0, 2, 3).contiguous().numpy(),
                                      truth.permute(1, 0, 2, 3).contiguous().numpy(), 0.1, feat_d=1)
    assert abs(loss - want) <= REL * abs(want)
    assert grad.shape == (1, 4, 24, 40)
    assert np.abs(grad - wgrad.transpose(1, 0, 2, 3)).max() <= REL * np.abs(wgrad).max()
