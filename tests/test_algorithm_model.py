"""Executable model of the GPU algorithm (level-0 elder-linked union-find + order-independent
triplet merge tree) in plain Python, checked against the sequential Kruskal scan with the edges fed
in RANDOM order.  This is the CPU evidence for the exactness argument in csrc/ph_kernel.cuh; the
CUDA kernels are checked against the oracle in test_gpu_parity.py."""
import random

import numpy as np
import pytest

INF = 1 << 70


def _mono(f):
    u = np.float32(f + np.float32(0)).view(np.uint32).item()
    return (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)


def _graph(f, dim):
    """Nodes / edges of the H0 (vertex) or H1 (dual, keys negated => ascending) graph, SURVEY 8a-note."""
    H, W = f.shape
    GW = 2 * W + 1
    fm = [[_mono(f[r, c]) for c in range(W)] for r in range(H)]
    vval = lambda i, j: fm[i][0] if j == 0 else fm[i][W - 1] if j == W else min(fm[i][j - 1], fm[i][j])
    hval = lambda i, j: fm[0][j] if i == 0 else fm[H - 1][j] if i == H else min(fm[i - 1][j], fm[i][j])
    edges = []
    if dim == 1:
        OUT = H * W
        nkey = [-((fm[k // W][k % W] << 32) | k) for k in range(H * W)] + [-INF]
        for i in range(H):
            for j in range(W + 1):
                edges.append((-((vval(i, j) << 32) | (2 * j + (2 * i + 1) * GW)), OUT if j == 0 else i * W + j - 1, OUT if j == W else i * W + j))
        for i in range(H + 1):
            for j in range(W):
                edges.append((-((hval(i, j) << 32) | (2 * j + 1 + 2 * i * GW)), OUT if i == 0 else (i - 1) * W + j, OUT if i == H else i * W + j))
        val = lambda k: (-k) >> 32
    else:
        VW = W + 1
        nkey = []
        for i in range(H + 1):
            for j in range(W + 1):
                v = min(fm[r][c] for r in (i - 1, i) for c in (j - 1, j) if 0 <= r < H and 0 <= c < W)
                nkey.append((v << 32) | (2 * j + 2 * i * GW))
        for i in range(H):
            for j in range(W + 1):
                edges.append(((vval(i, j) << 32) | (2 * j + (2 * i + 1) * GW), i * VW + j, (i + 1) * VW + j))
        for i in range(H + 1):
            for j in range(W):
                edges.append(((hval(i, j) << 32) | (2 * j + 1 + 2 * i * GW), i * VW + j, i * VW + j + 1))
        val = lambda k: k >> 32
    return nkey, edges, val


def _kruskal(nkey, edges):
    par = list(range(len(nkey)))

    def find(x):
        while par[x] != x:
            par[x] = par[par[x]]
            x = par[x]
        return x
    out = []
    for k, a, b in sorted(edges):
        ra, rb = find(a), find(b)
        if ra == rb:
            continue
        if nkey[ra] < nkey[rb]:
            ra, rb = rb, ra  # ra younger: dies
        par[ra] = rb
        out.append((ra, k))
    return out


def _model(nkey, edges, val, rng, level0, boruvka=0, dedup=False, bands=0):
    n = len(nkey)
    T = [(INF, x) for x in range(n)]  # (edge at which x dies, elder target)
    if level0:
        best = [None] * n
        for k, a, b in edges:
            for x, y in ((a, b), (b, a)):
                if best[x] is None or k < best[x][0]:
                    best[x] = (k, y)
        par = list(range(n))

        def find(x):
            while par[x] != x:
                par[x] = par[par[x]]
                x = par[x]
            return x
        order = list(range(n))
        rng.shuffle(order)
        for p in order:
            if nkey[p] == -INF:
                continue
            k, y = best[p]
            if val(k) == val(nkey[p]):  # zero-persistence merge along the earliest incident edge
                ra, rb = find(p), find(y)
                if ra != rb:
                    if nkey[ra] < nkey[rb]:
                        par[rb] = ra
                    else:
                        par[ra] = rb
        for p in range(n):
            r = find(p)
            if r != p:
                T[p] = (-INF, r)

    def rep(x, s):
        while T[x][0] <= s:
            x = T[x][1]
        return x
    es = list(edges)
    if boruvka:
        # elder-rule Boruvka contraction rounds on the basin graph (csrc/ph_small.cuh, boruvka_rounds):
        # a basin whose EARLIEST incident edge leads to an elder basin dies at that edge; survivors stay
        # roots; edges are relabelled to live ancestors and self loops dropped; the rest is merged below
        def root0(x):
            while T[x][0] == -INF:
                x = T[x][1]
            return x
        es = [(s, root0(a), root0(b)) for s, a, b in es]
        es = [e for e in es if e[1] != e[2]]
        dead = set()
        for _ in range(boruvka):
            best = {}
            for s, a, b in es:
                for x, y in ((a, b), (b, a)):
                    if x not in best or (s, y) < best[x]:
                        best[x] = (s, y)
            for x, (s, y) in best.items():
                if nkey[y] < nkey[x]:  # y elder: x dies here, final entry
                    T[x] = (s, y)
                    dead.add(x)

            def live(x):
                while x in dead:
                    x = T[x][1]
                return x
            es = [(s, live(a), live(b)) for s, a, b in es]
            es = [e for e in es if e[1] != e[2]]
    def root0_(x):
        while T[x][0] == -INF:
            x = T[x][1]
        return x
    if dedup:
        # csrc/ph_small.cuh, compaction: of several crossing edges that join the SAME two basins only the earliest can be
        # a tree edge; the kernel drops the later ones it can see in registers -- here ALL of them (the strongest form)
        best = {}
        for s, a, b in es:
            pa, pb = root0_(a), root0_(b)
            if pa == pb:
                continue
            k = (min(pa, pb), max(pa, pb))
            if k not in best or s < best[k][0]:
                best[k] = (s, a, b)
        es = list(best.values())
    rng.shuffle(es)  # ANY order
    if bands:
        # multi-band maps: every band's own edges first (band by band, as the kernel merges them on a band-local table
        # whose entries then move to the global table), the edges across band borders last
        band_of = lambda x: min(bands - 1, x * bands // len(nkey))
        es.sort(key=lambda e: band_of(e[1]) if band_of(e[1]) == band_of(e[2]) else bands)
    for s, x, y in es:
        while True:
            x, y = rep(x, s), rep(y, s)
            if x == y:
                break
            if nkey[y] < nkey[x]:
                x, y = y, x
            so, yo = T[y]
            assert so > s
            T[y] = (s, x)  # the CAS
            if so == INF:
                break
            y, s = yo, so  # re-assert the displaced connection
    return [(y, T[y][0]) for y in range(n) if T[y][0] not in (INF, -INF)]


@pytest.mark.parametrize("seed", range(6))
def test_triplet_merge_equals_kruskal_for_any_edge_order(seed):
    rng, nr = random.Random(seed), np.random.default_rng(seed)
    for t in range(40):
        H = W = int(nr.integers(1, 11))
        if t % 5 == 4:
            H, W = int(nr.integers(1, 9)), int(nr.integers(1, 9))
        mode = t % 3
        f = (nr.random((H, W)) if mode == 0 else nr.integers(0, [0, 4, 2][mode], (H, W))).astype(np.float32)
        for dim in (0, 1):
            nkey, edges, val = _graph(f, dim)
            want = sorted((y, s) for y, s in _kruskal(nkey, edges) if val(s) != val(nkey[y]))
            for level0 in (True, False):
                got = sorted((y, s) for y, s in _model(nkey, edges, val, rng, level0) if val(s) != val(nkey[y]))
                assert got == want, (dim, level0, f.tolist())
            for kw in (dict(dedup=True), dict(bands=3), dict(dedup=True, bands=2)):  # duplicate-edge filter; band-by-band order
                got = sorted((y, s) for y, s in _model(nkey, edges, val, rng, True, **kw) if val(s) != val(nkey[y]))
                assert got == want, (dim, kw, f.tolist())
            for rounds in (1, 2, 4, 50):  # contraction rounds first, lock-free merge for the rest
                got = sorted((y, s) for y, s in _model(nkey, edges, val, rng, True, boruvka=rounds) if val(s) != val(nkey[y]))
                assert got == want, (dim, "boruvka", rounds, f.tolist())
