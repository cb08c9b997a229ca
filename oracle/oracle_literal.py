"""Oracle L -- literal cell-level restatement of the reference's cubical persistence.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may (and there only as the checker / the timed CPU arm).

PARITY UNPINNED: the reference (``/root/reference/octsam/models/topological_loss.py:55-63``)
delegates this arithmetic to ``torch_topological.nn.CubicalComplex`` -> ``gudhi.CubicalComplex``
(neither vendored, pinned nor installed).  This file restates the
published algorithm of those libraries:

* gudhi ``Bitmap_cubical_complex_base`` -- T-construction: pixels are the 2-cells of a
  (2H+1)x(2W+1) bitmap, every lower cell gets the min of its cofaces;
* gudhi ``Bitmap_cubical_complex::is_before_in_filtration`` -- total order
  (filtration value, dimension, bitmap position);
* gudhi ``Persistent_cohomology`` -- the persistence pairing of that total order (unique, so a
  Z/2 boundary-matrix reduction gives the same pairs), intervals kept only when
  death - birth > 0 strictly (``min_persistence = 0``);
* gudhi ``get_top_dimensional_coface_of_a_cell`` / ``cofaces_of_persistence_pairs`` -- every
  paired cell is mapped to a pixel by walking to the FIRST coboundary cell of equal value
  (coboundary enumerated last axis first);
* torch_topological ``CubicalComplex._extract_generators_and_diagrams`` -- the essential H0
  class is paired with ``argmax(x)``; diagram values are re-gathered from ``x``.

Worst case O(cells^2), in practice seconds for a 256 x 256 map (tests/test_oracle.py checks the fast oracle against it
at that size too).

KNOWN RISK, H0 ONLY (unverifiable here): the persistence pairing of a TOTAL order is unique, and
for H1 gudhi derives it from that order through its general cohomology reduction -- which is what this file
computes.  For dimension 0, however, gudhi's ``Persistent_cohomology::update_cohomology_groups_edge`` takes a
union-find short cut and decides which of two merging components dies by comparing the FILTRATION VALUES of
their creators only (``filtration(idx_coc_u) < filtration(idx_coc_v)`` kills v's class, anything else --
including a tie -- kills u's, u being the first boundary vertex of the merging edge, i.e. the one with the
smaller bitmap position) [UPSTREAM-RECALL].  When the minima of the two components are exactly equal the
diagram VALUES are the same under both rules, but the CREATOR PIXEL of the pair (hence where the gradient
lands) can differ from the canonical (value, dim, position) pairing implemented here, in the fast oracle and
in the CUDA kernels.  The same short cut also runs for zero-persistence merges, so with tied PIXEL values the
surviving creator -- and with it the pixel of the essential class -- can move to another pixel of the same
value (all-zero backgrounds of binary maps).  ``h0_pairs_gudhi_union_find`` below restates that short cut so that the two candidates
can be told apart the day gudhi is importable (``tests/kats.py::TIE_KATS`` holds an image on which they
differ; ``tests/golden/make_golden_reference.py`` records what the real library does).  H1 -- the dimension
the reference call site uses (feat_d=1, training_utils.py:64) -- is not affected.
"""
from __future__ import annotations

import numpy as np


def gudhi_bitmap_as_image(top_dimensional_cells, dimensions) -> np.ndarray:
    """gudhi ``Bitmap_cubical_complex_base(dimensions, top_dimensional_cells)``: cell k of the flat list sits at
    coordinates ``(k % dimensions[0], k // dimensions[0])`` -- the FIRST dimension is the fastest-varying one
    [UPSTREAM-RECALL].  Returned as the 2-D image whose raster order is the flat order: ``dimensions[1]`` rows of
    ``dimensions[0]`` pixels.  torch_topological's ``CubicalComplex._forward`` calls it with
    ``dimensions=x.shape, top_dimensional_cells=x.flatten()``, i.e. un-reversed: an H x W tensor is read as W rows
    of H pixels (nothing changes for H == W); the indices ``cofaces_of_persistence_pairs`` returns count cells in
    the same flat order, so ``x.ravel()[idx]`` and ``np.unravel_index(idx, x.shape)`` address the tensor's own
    pixels."""
    d0, d1 = int(dimensions[0]), int(dimensions[1])
    return np.asarray(top_dimensional_cells).reshape(d1, d0)


def _cells(f: np.ndarray):
    """Doubled grid: values (min of cofaces), dims, positions.  SURVEY.md appendix C steps 1-2."""
    H, W = f.shape
    GH, GW = 2 * H + 1, 2 * W + 1
    val = np.full((GH, GW), np.inf, dtype=np.float64)
    for r in range(H):
        for c in range(W):
            v = float(f[r, c])
            Y, X = 2 * r + 1, 2 * c + 1
            blk = val[Y - 1:Y + 2, X - 1:X + 2]
            np.minimum(blk, v, out=blk)
    return val, GH, GW


def _top_cell(pos: int, val: np.ndarray, GH: int, GW: int) -> int:
    """gudhi get_top_dimensional_coface_of_a_cell: first coboundary cell with equal value."""
    Y, X = divmod(pos, GW)
    if (Y & 1) and (X & 1):
        return pos
    v = val[Y, X]
    cob = []
    if Y % 2 == 0:  # last axis (rows of the bitmap) first
        if Y - 1 >= 0:
            cob.append(pos - GW)
        if Y + 1 < GH:
            cob.append(pos + GW)
    if X % 2 == 0:
        if X - 1 >= 0:
            cob.append(pos - 1)
        if X + 1 < GW:
            cob.append(pos + 1)
    for q in cob:
        qy, qx = divmod(q, GW)
        if val[qy, qx] == v:
            return _top_cell(q, val, GH, GW)
    raise AssertionError("no coface of equal value")


def _pixel_of(pos: int, GW: int, W: int) -> int:
    Y, X = divmod(pos, GW)
    return ((Y - 1) // 2) * W + (X - 1) // 2


def cubical_pairs_literal(f: np.ndarray):
    """Return ``(h0_regular, h1_regular, h0_essential)`` for an HxW float image.

    Each regular list holds ``(creator_pixel, destroyer_pixel)`` flat C-order indices in
    gudhi's emission order (filtration order of the death cell); ``h0_essential`` is
    ``(creator_pixel, argmax_pixel)`` (torch_topological's fake destroyer).
    """
    f = np.asarray(f)
    assert f.ndim == 2
    H, W = f.shape
    val, GH, GW = _cells(f)
    ncell = GH * GW
    Ys, Xs = np.divmod(np.arange(ncell), GW)
    dims = (Ys % 2) + (Xs % 2)
    vflat = val.ravel()
    order = sorted(range(ncell), key=lambda p: (vflat[p], dims[p], p))
    rank = np.empty(ncell, dtype=np.int64)
    for i, p in enumerate(order):
        rank[p] = i

    def boundary(p):
        Y, X = divmod(p, GW)
        out = []
        if X & 1:
            out += [p - 1, p + 1]
        if Y & 1:
            out += [p - GW, p + GW]
        return out

    # standard column reduction over Z/2 in filtration order; columns as sets of ranks
    low_to_col = {}
    columns = {}
    paired_birth = {}
    for i, p in enumerate(order):
        col = set(int(rank[q]) for q in boundary(p))
        while col:
            lo = max(col)
            j = low_to_col.get(lo)
            if j is None:
                break
            col ^= columns[j]
        if col:
            lo = max(col)
            low_to_col[lo] = i
            columns[i] = col
            paired_birth[lo] = i
    deaths = set(paired_birth.values())
    h0, h1 = [], []
    for lo in sorted(paired_birth, key=lambda b: paired_birth[b]):  # order of death cell
        i = paired_birth[lo]
        bcell, dcell = order[lo], order[i]
        if not (vflat[dcell] > vflat[bcell]):
            continue
        pb = _pixel_of(_top_cell(bcell, val, GH, GW), GW, W)
        pd = _pixel_of(_top_cell(dcell, val, GH, GW), GW, W)
        (h0 if dims[bcell] == 0 else h1).append((pb, pd))
    ess = [order[i] for i in range(ncell) if i not in paired_birth and i not in deaths]
    assert len(ess) == 1 and dims[ess[0]] == 0, "a filled rectangle has exactly one essential class"
    pe = _pixel_of(_top_cell(ess[0], val, GH, GW), GW, W)
    amax = int(np.argmax(f))  # first max in raster order (torch.argmax on CPU)
    return h0, h1, (pe, amax)


def h0_pairs_gudhi_union_find(f: np.ndarray):
    """H0 pairs under gudhi's union-find short cut for edges as recalled in the module docstring
    [UPSTREAM-RECALL]: vertices create classes in filtration order; an edge joining two classes kills the class
    whose creator has the LARGER filtration value, and on a tie the class of the edge's first boundary vertex
    (smaller bitmap position).  Returns ``(h0_regular, h0_essential)`` in the format of
    ``cubical_pairs_literal``.  Differs from the canonical pairing only when two merging components have exactly
    equal minima."""
    f = np.asarray(f)
    H, W = f.shape
    val, GH, GW = _cells(f)
    vflat = val.ravel()
    ncell = GH * GW
    Ys, Xs = np.divmod(np.arange(ncell), GW)
    dims = (Ys % 2) + (Xs % 2)
    order = sorted(range(ncell), key=lambda p: (vflat[p], dims[p], p))
    parent, creator = {}, {}

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    h0 = []
    for p in order:
        if dims[p] == 0:
            parent[p] = p
            creator[p] = p
        elif dims[p] == 1:
            Y, X = divmod(p, GW)
            u, v = (p - 1, p + 1) if X & 1 else (p - GW, p + GW)
            ru, rv = find(u), find(v)
            if ru == rv:
                continue
            cu, cv = creator[ru], creator[rv]
            if vflat[cu] < vflat[cv]:
                dead, live_c = cv, cu
            else:
                dead, live_c = cu, cv
            parent[ru] = rv
            creator[rv] = live_c
            if vflat[p] > vflat[dead]:
                h0.append((_pixel_of(_top_cell(dead, val, GH, GW), GW, W), _pixel_of(_top_cell(p, val, GH, GW), GW, W)))
    root = find(order[0])
    ess = (_pixel_of(_top_cell(creator[root], val, GH, GW), GW, W), int(np.argmax(f)))
    return h0, ess
