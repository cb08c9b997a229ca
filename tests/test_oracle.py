"""CPU suite: pins the fast oracle (oracle/topo_oracle.c) against the literal cell-complex
restatement, the hand-derived KATs, scipy's exact assignment and a torch-autograd restatement of
torch_topological's WassersteinDistance.  (The reference itself has no tests -- SURVEY.md 8c.)"""
import json
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st
from scipy.optimize import linear_sum_assignment

import oracle
from oracle.oracle_literal import cubical_pairs_literal
from tests.kats import KATS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pairs_small.json")


@pytest.mark.parametrize("name", sorted(KATS))
def test_kats_literal_and_fast(name):
    img, h0, h1, ess = KATS[name]
    f = np.array(img, dtype=np.float32)
    l0, l1, less = cubical_pairs_literal(f)
    assert sorted(l0) == sorted(h0) and sorted(l1) == sorted(h1) and less == ess
    f0 = [tuple(x) for x in oracle.cubical_pairs(f, 0)]
    f1 = [tuple(x) for x in oracle.cubical_pairs(f, 1)]
    assert f0 == l0 + [less] and f1 == l1


def test_golden_fixture_matches_fast_oracle():
    data = json.load(open(GOLDEN))
    for case in data["pairs"]:
        f = np.array(case["image"], dtype=np.float32)
        assert oracle.cubical_pairs(f, 0).tolist() == case["h0"]
        assert oracle.cubical_pairs(f, 1).tolist() == case["h1"]


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 9), st.integers(1, 9), st.integers(0, 2), st.integers(0, 2 ** 31 - 1))
def test_fast_equals_literal(h, w, mode, seed):
    rng = np.random.default_rng(seed)
    f = (rng.random((h, w)) if mode == 0 else rng.integers(0, [0, 4, 2][mode], (h, w))).astype(np.float32)
    l0, l1, ess = cubical_pairs_literal(f)
    assert [tuple(x) for x in oracle.cubical_pairs(f, 0)] == l0 + [ess]
    assert [tuple(x) for x in oracle.cubical_pairs(f, 1)] == l1


@pytest.mark.parametrize("kind", ["pred", "truth", "ties32", "iid", "rect"])
def test_fast_equals_literal_at_headline_size(kind):
    """Oracle F (the checker of the GPU parity tests at full size) against the literal boundary-matrix reduction on
    256 x 256 maps of BASELINE's workload -- incl. emission order -- and on a rectangular map."""
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, truth = make_batch(1, 256, 256, seed=1234)
    rng = np.random.default_rng(11)
    f = {"pred": lambda: pred[0, 3].numpy(), "truth": lambda: truth[0, 3].numpy(),
         "ties32": lambda: (np.round(pred[0, 5].numpy() * 32) / 32).astype(np.float32),
         "iid": lambda: rng.random((256, 256)).astype(np.float32),
         "rect": lambda: rng.random((96, 200)).astype(np.float32)}[kind]()
    l0, l1, ess = cubical_pairs_literal(f)
    assert [tuple(x) for x in oracle.cubical_pairs(f, 0)] == l0 + [ess]
    assert [tuple(x) for x in oracle.cubical_pairs(f, 1)] == l1
    if kind != "truth":
        assert len(l1) > 1000


def _torch_cost_matrix(D1, D2, q):
    """torch_topological WassersteinDistance._make_distance_matrix, restated."""
    def proj(d):
        x, y = d[:, 0], d[:, 1]
        return 0.5 * torch.stack(((x + y), (x + y)), 1)
    d11 = torch.linalg.vector_norm(D1 - proj(D1), float("inf"), dim=1)
    d22 = torch.linalg.vector_norm(D2 - proj(D2), float("inf"), dim=1)
    dist = torch.cdist(D1, D2, p=float("inf"))
    up = torch.hstack((dist, d11[:, None]))
    lo = torch.cat((d22, torch.tensor(0.0).unsqueeze(0)))
    return torch.vstack((up, lo)).pow(q)


def _diagrams(rng, n, m, binary_truth=False):
    b = rng.random(n).astype(np.float32)
    D1 = np.stack([b, b + rng.random(n).astype(np.float32)], 1).reshape(-1, 2)
    b = rng.random(m).astype(np.float32)
    D2 = np.stack([b, b + rng.random(m).astype(np.float32)], 1).reshape(-1, 2)
    if binary_truth and m:
        D2[:] = np.array([0.0, 1.0], np.float32)
    return D1, D2


@pytest.mark.parametrize("q", [1, 2, 3])
def test_wasserstein_is_the_exact_lp_optimum(q):
    rng = np.random.default_rng(q)
    for t in range(120):
        n, m = int(rng.integers(0, 10)), int(rng.integers(0, 10))
        D1, D2 = _diagrams(rng, n, m, t % 5 == 0)
        cost, match = oracle.wasserstein(D1, D2, q)
        M = _torch_cost_matrix(torch.tensor(D1), torch.tensor(D2), q).double().numpy()
        big = np.full((n + m, n + m), 1e9)
        big[:n, :m] = M[:n, :m]
        for i in range(n):
            big[i, m + i] = M[i, m]
        for j in range(m):
            big[n + j, j] = M[n, j]
        big[n:, m:] = 0.0
        ref = big[linear_sum_assignment(big)].sum() if n + m else 0.0
        assert abs(cost - ref) < 1e-6
        # the returned matching attains the reported cost
        tot = sum(M[i, match[i]] if match[i] >= 0 else M[i, m] for i in range(n))
        tot += sum(M[n, j] for j in range(m) if j not in set(match[match >= 0].tolist()))
        assert abs(tot - cost) < 1e-6


@pytest.mark.parametrize("q", [1, 2])
def test_wasserstein_equals_the_literal_transport_lp(q):
    """The reference's own formulation, literally: ``ot.emd2(a, b, M)`` on the (n+1) x (m+1) matrix with masses
    a = [1]*n + [m], b = [1]*m + [n] (torch_topological WassersteinDistance.forward, called at
    topological_loss.py:78-82) -- solved as a transport LP with scipy's HiGHS instead of POT's network simplex.  The
    oracle (and the kernels) solve the equivalent partial assignment; this pins the equivalence."""
    from scipy.optimize import linprog
    rng = np.random.default_rng(40 + q)
    for t in range(60):
        n, m = int(rng.integers(0, 9)), int(rng.integers(0, 7))
        D1, D2 = _diagrams(rng, n, m, t % 4 == 0)
        cost, _ = oracle.wasserstein(D1, D2, q)
        M = _torch_cost_matrix(torch.tensor(D1), torch.tensor(D2), q).double().numpy()
        a = np.array([1.0] * n + [float(m)])
        b = np.array([1.0] * m + [float(n)])
        R, Cc = n + 1, m + 1
        A_eq = np.zeros((R + Cc, R * Cc))
        for i in range(R):
            A_eq[i, i * Cc:(i + 1) * Cc] = 1.0          # row sums = a
        for j in range(Cc):
            A_eq[R + j, j::Cc] = 1.0                   # column sums = b
        res = linprog(M.ravel(), A_eq=A_eq, b_eq=np.concatenate([a, b]), bounds=(0, None), method="highs")
        assert res.status == 0
        assert abs(res.fun - cost) < 1e-7 + 1e-7 * abs(cost), (n, m, res.fun, cost)


def test_kat_w2_toy_and_gradient_tie():
    cost, match = oracle.wasserstein(np.array([[0.2, 0.8]], np.float32), np.array([[0.0, 1.0]], np.float32), 2)
    assert abs(cost - 0.04) < 1e-7 and match[0] == 0  # W = 0.2, not 0.34 via the diagonal
    D1 = torch.tensor([[0.25, 0.75]], requires_grad=True)
    M = _torch_cost_matrix(D1, torch.tensor([[0.0, 1.0]]), 2)
    M[0, 0].backward()
    assert torch.allclose(D1.grad, torch.tensor([[0.5, -0.5]]))  # both coordinates at an exact L-inf tie


@pytest.mark.parametrize("feat_d", [0, 1])
@pytest.mark.parametrize("q", [1, 2])
def test_loss_and_gradient_match_torch_autograd_restatement(feat_d, q):
    """Full path restated with torch autograd (gather -> cost matrix -> plan . M -> pow(1/q) -> mean),
    using the oracle's pairs and matching as the (non-differentiable) combinatorial part."""
    rng = np.random.default_rng(10 * feat_d + q)
    B, C, n = 2, 3, 10
    pred = rng.random((B, C, n, n)).astype(np.float32)
    truth = (np.round(rng.random((B, C, n, n)) * 3) / 3).astype(np.float32)
    lam = 0.1
    loss, grad, _ = oracle.topo_loss(pred, truth, lam, feat_d=feat_d, loss_q=q)
    x = torch.tensor(pred, requires_grad=True)
    per_image = []
    for b in range(B):
        total = 0.0
        for c in range(C):
            p1 = torch.as_tensor(oracle.cubical_pairs(pred[b, c], feat_d), dtype=torch.long)
            p2 = torch.as_tensor(oracle.cubical_pairs(truth[b, c], feat_d), dtype=torch.long)
            xf, tf = x[b, c].ravel(), torch.tensor(truth[b, c]).ravel()
            D1 = torch.stack((xf[p1[:, 0]], xf[p1[:, 1]]), 1)
            D2 = torch.stack((tf[p2[:, 0]], tf[p2[:, 1]]), 1)
            M = _torch_cost_matrix(D1, D2, q)
            _, match = oracle.wasserstein(D1.detach().numpy(), D2.numpy(), q)
            G = torch.zeros_like(M)
            for i, j in enumerate(match):
                G[i, j if j >= 0 else len(D2)] = 1.0
            used = set(match[match >= 0].tolist())
            for j in range(len(D2)):
                if j not in used:
                    G[len(D1), j] = 1.0
            total = total + (G * M).sum()
        per_image.append(total.pow(1.0 / q))
    ref = lam * torch.stack(per_image).mean()
    ref.backward()
    assert abs(float(ref) - loss) <= 1e-5 * abs(loss)
    assert np.abs(x.grad.numpy() - grad).max() <= 1e-5 * np.abs(grad).max()


def test_golden_losses():
    data = json.load(open(GOLDEN))
    for case in data["losses"]:
        loss, grad, _ = oracle.topo_loss(np.array(case["pred"], np.float32), np.array(case["truth"], np.float32),
                                         case["lamda"], feat_d=case["feat_d"], loss_q=case["q"])
        assert abs(loss - case["loss"]) <= 1e-6 * abs(case["loss"])
        assert np.allclose(grad, np.array(case["grad"], np.float32), rtol=1e-6, atol=1e-9)


def test_threads_do_not_change_the_result():
    rng = np.random.default_rng(2)
    pred = rng.random((2, 4, 24, 24)).astype(np.float32)
    truth = (rng.random((2, 4, 24, 24)) < 0.4).astype(np.float32)
    a = oracle.topo_loss(pred, truth, 0.1, feat_d=1, nthreads=1)
    b = oracle.topo_loss(pred, truth, 0.1, feat_d=1, nthreads=4)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


# ---- independent pin of the filtration semantics: Betti curves from connected-component labelling
# (scipy.ndimage.label, no union-find of ours).  gudhi's T-construction makes sublevel sets unions of CLOSED
# pixels: components are 8-connected, holes are 4-connected components of the complement that do not touch
# the border.  For every threshold t,
#   beta_0(t) = #{H0 pairs: b <= t < d} (+1 for the essential class once the minimum is in)
#   beta_1(t) = #{H1 pairs: b <= t < d}
# must equal the component / hole counts of {f <= t}.
@pytest.mark.parametrize("seed,levels", [(0, 0), (1, 0), (2, 6), (3, 3), (4, 2)])
def test_betti_curves_match_connected_component_counts(seed, levels):
    from scipy import ndimage
    rng = np.random.default_rng(100 + seed)
    H, W = int(rng.integers(6, 20)), int(rng.integers(6, 20))
    f = rng.random((H, W)).astype(np.float32)
    if levels:
        f = (np.floor(f * levels) / levels).astype(np.float32)
    flat = f.ravel()
    p0 = oracle.cubical_pairs(f, 0)
    p1 = oracle.cubical_pairs(f, 1)
    ess = p0[-1]                      # essential class last: (global-min pixel, argmax)
    fin0 = p0[:-1]
    assert flat[ess[0]] == f.min() and ess[1] == int(np.argmax(f))
    eight = np.ones((3, 3), dtype=int)
    four = ndimage.generate_binary_structure(2, 1)
    for t in np.unique(f):
        sub = f <= t
        n_comp = ndimage.label(sub, structure=eight)[1]
        lab, n_bg = ndimage.label(~sub, structure=four)
        border = set(np.unique(np.concatenate([lab[0], lab[-1], lab[:, 0], lab[:, -1]]))) - {0}
        n_holes = n_bg - len(border)
        b0 = int(np.sum((flat[fin0[:, 0]] <= t) & (t < flat[fin0[:, 1]]))) + 1 if len(fin0) else 1
        b1 = int(np.sum((flat[p1[:, 0]] <= t) & (t < flat[p1[:, 1]]))) if len(p1) else 0
        assert b0 == n_comp, (t, b0, n_comp)
        assert b1 == n_holes, (t, b1, n_holes)


def test_h0_tie_rule_candidates_are_told_apart():
    """The canonical pairing (implemented everywhere in this repo) and gudhi's recalled union-find short cut for
    dimension 0 agree on tie-free maps and on the diagram VALUES always; they differ in the creator pixel when two
    merging components have equal minima.  This pins which images would expose the difference."""
    from oracle.oracle_literal import h0_pairs_gudhi_union_find
    from tests.kats import KATS, TIE_KATS
    for name, kat in TIE_KATS.items():
        f = np.array(kat["image"], dtype=np.float32)
        l0, l1, less = cubical_pairs_literal(f)
        assert sorted(l0) == sorted(kat["canonical"]["h0"]) and less == kat["canonical"]["ess"] and l1 == kat["h1"], name
        assert [tuple(x) for x in oracle.cubical_pairs(f, 0)] == l0 + [less]
        g0, gess = h0_pairs_gudhi_union_find(f)
        if kat["gudhi_union_find"] is not None:
            assert sorted(g0) == sorted(kat["gudhi_union_find"]["h0"]) and gess == kat["gudhi_union_find"]["ess"], name
        # same diagram either way
        vals = lambda pairs: sorted((float(f.ravel()[a]), float(f.ravel()[b])) for a, b in pairs)
        assert vals(g0) == vals(l0)
    assert h0_pairs_gudhi_union_find(np.array(TIE_KATS["two_equal_minima"]["image"], np.float32))[0] != \
        cubical_pairs_literal(np.array(TIE_KATS["two_equal_minima"]["image"], np.float32))[0]
    rng = np.random.default_rng(0)
    for _ in range(30):  # tie-free: identical
        f = rng.permutation(49).reshape(7, 7).astype(np.float32)
        l0, _, less = cubical_pairs_literal(f)
        g0, gess = h0_pairs_gudhi_union_find(f)
        assert g0 == l0 and gess == less
    differ = []
    for name, (img, h0, h1, ess) in KATS.items():
        g0, gess = h0_pairs_gudhi_union_find(np.array(img, np.float32))
        if len(set(np.array(img).ravel().tolist())) == np.array(img).size:  # all pixel values distinct: both rules agree
            assert sorted(g0) == sorted(h0) and gess == ess, name
        elif not (sorted(g0) == sorted(h0) and gess == ess):
            differ.append(name)
    # with tied pixel values even zero-persistence merges move the surviving creator under the recalled rule, so
    # the essential class can land on another pixel of the same value (e.g. the all-zero background)
    assert set(differ) <= {"two_diag_holes", "hole_touching_border", "nested"}


def test_two_valued_h1_rule():
    """csrc/ph_binary.cuh: on a map with two values the H1 pairs are the 4-connected hi blobs off the border,
    destroyer = last raster pixel of the blob, creator = the pixel below it, in raster order of the destroyer."""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(1)
    for t in range(400):
        n = int(rng.integers(2, 14))
        f = (rng.random((n, n)) < rng.choice([0.05, 0.2, 0.5, 0.8, 0.95])).astype(np.float32)
        if t % 3 == 1:
            f = f * 3.5 - 1.25
        if t % 3 == 2:
            f = -f
        want = oracle.cubical_pairs(f, 1)
        got = []
        if f.min() != f.max():
            lab, k = ndi.label(f == f.max())
            for i in range(1, k + 1):
                ys, xs = np.nonzero(lab == i)
                if ys.min() == 0 or xs.min() == 0 or ys.max() == n - 1 or xs.max() == n - 1:
                    continue
                p = int((ys * n + xs).max())
                got.append((p + n, p))
        got.sort(key=lambda x: x[1])
        assert [tuple(x) for x in want.tolist()] == got


def test_nonsquare_maps_follow_the_reference_shape_order():
    """torch_topological passes ``dimensions=x.shape`` un-reversed to gudhi, whose first dimension is the fastest
    one: an H x W map (H != W) is read as W rows of H pixels.  The oracle's topo_loss reproduces that; flat indices
    and the gradient layout are unchanged; square maps are untouched."""
    from oracle import oracle_literal as L
    rng = np.random.default_rng(3)
    x = rng.random((3, 5)).astype(np.float32)
    img = L.gudhi_bitmap_as_image(x.ravel(), x.shape)
    assert img.shape == (5, 3) and np.array_equal(img.ravel(), x.ravel())
    sq = rng.random((4, 4)).astype(np.float32)
    assert np.array_equal(L.gudhi_bitmap_as_image(sq.ravel(), sq.shape), sq)
    pred = rng.random((2, 2, 6, 10)).astype(np.float32)
    truth = (rng.random((2, 2, 6, 10)) > 0.5).astype(np.float32)
    l1, g1, _ = oracle.topo_loss(pred, truth, 0.1, feat_d=1)
    l2, g2, _ = oracle.topo_loss(pred.reshape(2, 2, 10, 6), truth.reshape(2, 2, 10, 6), 0.1, feat_d=1, reference_shape_order=False)
    assert l1 == l2 and g1.shape == pred.shape and np.array_equal(g1.ravel(), g2.ravel())
    # ... and it is NOT the loss of the transposed or of the plainly read image in general
    pairs_plain = oracle.cubical_pairs(pred[0, 0], 1)
    pairs_ref = oracle.cubical_pairs(L.gudhi_bitmap_as_image(pred[0, 0].ravel(), pred[0, 0].shape), 1)
    assert pairs_plain.shape != pairs_ref.shape or not np.array_equal(pairs_plain, pairs_ref)


def test_h0_tie_rule_candidates_agree_on_values_and_on_tie_free_maps():
    """The two candidate rules for gudhi's dimension-0 pairing (canonical total order -- implemented everywhere here -- vs
    the union-find short cut recalled from Persistent_cohomology::update_cohomology_groups_edge):
    identical pairs on maps without exactly tied values (fp32 sigmoid outputs), and on tie-heavy maps the same MULTISET of
    (birth, death) values -- so the loss value never depends on the rule, only which of several equal-valued pixels
    receives the gradient does."""
    from oracle.oracle_literal import h0_pairs_gudhi_union_find
    from dilabhelmholtzoct_b200.synthetic import make_batch
    pred, _ = make_batch(1, 64, 64, seed=77, n_classes=4)
    for levels in (0, 1024, 16):
        f = pred[0, 1].numpy()
        if levels:
            f = (np.round(f * levels) / levels).astype(np.float32)
        h0u, essu = h0_pairs_gudhi_union_find(f)
        can = oracle.cubical_pairs(f, 0)
        h0c, essc = [tuple(int(v) for v in x) for x in can[:-1]], tuple(int(v) for v in can[-1])
        fl = f.ravel()
        assert sorted((fl[a], fl[b]) for a, b in h0u) == sorted((fl[a], fl[b]) for a, b in h0c)
        assert fl[essu[0]] == fl[essc[0]] and essu[1] == essc[1]
        if not levels:
            assert sorted(h0u) == sorted(h0c) and essu == essc
