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

It is O(cells^2)-ish and meant for maps up to ~32x32.
"""
from __future__ import annotations

import numpy as np


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
