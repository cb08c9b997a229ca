// Cubical persistence of a batch of maps: one CTA per (image, class) map.
//
// Replaces gudhi's sort-all-cells + persistent-cohomology reduction (reached from
// /root/reference/octsam/models/topological_loss.py:62-63) by a lock-free, sort-free scheme:
//
//   level 0  every node unions with the far end of its EARLIEST incident edge whenever that
//            edge has the node's own value (a zero-persistence merge).  Elder-linked lock-free
//            union-find (CAS on the younger root).  This contracts the grid to its basins.
//   level 1  every remaining edge is merged into a triplet merge tree
//            T[y] = (edge at which y's component dies, an elder node it merges into)
//            with the order-independent, CAS-based Merge of Smirnov & Morozov ("Triplet merge
//            trees"): edges may arrive in any order and from any thread; the fixed point is the
//            elder-rule pairing of the sequential Kruskal scan.
//   emit     every level-0 root y with a recorded edge is one persistence pair; zero-persistence
//            pairs are dropped (gudhi min_persistence = 0, strict).
//
// All comparisons use the exact cell order (value, bitmap position), so persistence pairs and
// critical pixels are bit-exact under ties.
//
// Why the lock-free Merge is exact (tests/test_algorithm_model.py is the executable version):
//   * Read every entry T[y] = (s, x) as the FACT "y and x are connected at level s", x elder than y.
//     Facts are monotone in the level: true at s implies true at every later level.
//   * Merge(a, s, b) first walks both sides to their representatives at level s, using only recorded
//     facts with level <= s.  A stale read is harmless: an entry is only ever replaced by a fact with an
//     EARLIER level for the same node, and the replaced fact is re-asserted by the replacing thread, so
//     everything a thread has read stays true.
//   * The single write is a CAS on the younger representative y: (s_old, x_old) -> (s, x) with s < s_old,
//     after which the same thread continues with Merge(x, s_old, x_old).  Old fact + pending edge and new
//     fact + pending re-assertion have the same closure, so the set of derivable connectivity facts is an
//     invariant of every atomic step; at quiescence it equals the closure of all edges.
//   * Every link points to a strictly elder node, so the pointer graph is acyclic and walks terminate;
//     each successful CAS strictly lowers one entry's level, so the whole process terminates.
//   * At quiescence each T[y].s is y's true death level: if y were connected to an elder node earlier
//     than s, that connection would have to be derivable from facts of level < s, but those only join y
//     to nodes hanging BELOW y (younger), because y's own link has level s.
//   * Level-0 contraction is exact because forest paths are key-monotone towards the basin root: when a
//     crossing edge is scanned, both its ends are already joined to their basin's core, and a basin root
//     whose value exceeds the edge's value is already the eldest node of everything joined to it.
#pragma once
#include "tl_common.cuh"

namespace tl {

constexpr int kPhThreads = 1024;

struct PhArgs {
    const float* maps[2];   // set 0 (pred) / set 1 (truth); maps[1] may be null when n_sets == 1
    PairStore ps;           // record arena, per-map offsets / counts / diagonal-cost sums, status word
    int n_sets, n_maps, H, W;
    int cap;                // most pairs one map can have (per-CTA scratch strides)
    uint64_t* T;            // [gridDim.x][t_stride]
    size_t t_stride;
    unsigned int* job_counter;
    uint64_t magic_W, magic_GW, magic_VW;  // FastDiv magics of W, 2W+1, W+1 (a 64-bit division per thread otherwise)
};

template <int DIM>
struct Ph {
    Geo<DIM> g;
    uint64_t* T;

    __device__ __forceinline__ Ph(const float* f, int H, int W, uint64_t* T_) : g(f, H, W), T(T_) {}

    __device__ __forceinline__ int find0(int x) const {
        for (;;) {
            const uint64_t t = ld_cg_u64(T + x);
            if ((uint32_t)(t >> 32) != kCodeL0) return x;
            const int p = (int)(uint32_t)t;
            const uint64_t tp = ld_cg_u64(T + p);
            if ((uint32_t)(tp >> 32) != kCodeL0) return p;
            T[x] = tp;  // path halving: a level-0 entry only ever holds an ancestor
            x = (int)(uint32_t)tp;
        }
    }

    // read-only variant for the flatten pass (a halving store there could regress an entry that
    // another thread has already pointed at the root)
    __device__ __forceinline__ int find0_ro(int x) const {
        for (;;) {
            const uint64_t t = ld_cg_u64(T + x);
            if ((uint32_t)(t >> 32) != kCodeL0) return x;
            x = (int)(uint32_t)t;
        }
    }

    // level-0 union: link the younger root under the elder one
    __device__ void union0(int a, int b) {
        for (;;) {
            int ra = find0(a), rb = find0(b);
            if (ra == rb) return;
            if (g.nkey(rb) < g.nkey(ra)) { int t = ra; ra = rb; rb = t; }
            unsigned long long expect = ((unsigned long long)kCodeRoot << 32) | (uint32_t)rb;
            unsigned long long want = ((unsigned long long)kCodeL0 << 32) | (uint32_t)ra;
            unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(T + rb), expect, want);
            if (old == expect) return;
            a = ra; b = rb;
        }
    }

    // representative of x at level skey: follow entries merged at or before skey
    __device__ __forceinline__ int rep(int x, uint64_t skey, uint64_t& entry) const {
        for (;;) {
            uint64_t t = ld_cg_u64(T + x);
            uint32_t code = (uint32_t)(t >> 32);
            if (code == kCodeRoot) { entry = t; return x; }
            if (code != kCodeL0 && g.ekey(code - 1) > skey) { entry = t; return x; }
            x = (int)(uint32_t)t;
        }
    }

    // triplet-merge-tree Merge(a, edge, b), lock-free
    __device__ void merge(int a, int b, uint32_t pos, uint64_t skey) {
        uint32_t code = pos + 1;
        for (;;) {
            uint64_t ta, tb;
            int x = rep(a, skey, ta), y = rep(b, skey, tb);
            if (x == y) return;
            if (g.nkey(y) < g.nkey(x)) { int t = x; x = y; y = t; tb = ta; }
            // y is the younger representative: it dies at this edge into x
            unsigned long long want = ((unsigned long long)code << 32) | (uint32_t)x;
            unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(T + y),
                                               (unsigned long long)tb, want);
            if (old == tb) {
                uint32_t ocode = (uint32_t)(tb >> 32);
                if (ocode == kCodeRoot) return;
                // y used to die at `ocode` into b': re-assert that connection for x
                a = x; b = (int)(uint32_t)tb; code = ocode; skey = g.ekey(ocode - 1);
            } else {
                a = x; b = y;
            }
        }
    }
};

template <int DIM>
__global__ void __launch_bounds__(kPhThreads) ph_kernel(PhArgs A) {
    __shared__ unsigned int s_job;
    __shared__ int s_count, s_avail;
    __shared__ unsigned long long s_argmax, s_base;
    __shared__ int s_wcnt[kPhThreads / 32];
    __shared__ double s_red[kPhThreads / 32];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int H = A.H, W = A.W, N = H * W;
    uint64_t* T = A.T + (size_t)blockIdx.x * A.t_stride;
    const unsigned int n_jobs = (unsigned)A.n_sets * (unsigned)A.n_maps;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_job = atomicAdd(A.job_counter, 1u); s_count = 0; s_avail = 0; s_argmax = 0ull; }
        __syncthreads();
        const unsigned int job = s_job;
        if (job >= n_jobs) break;
        // all prediction maps first (heavy), ground-truth maps (light) fill the tail of the launch
        const int set = (int)(job / (unsigned)A.n_maps), map = (int)(job % (unsigned)A.n_maps);
        Ph<DIM> ph(A.maps[set] + (size_t)map * N, H, W, T);
        const Geo<DIM>& g = ph.g;
        const int NN = g.NN, GW = g.GW, VW = g.VW;

        // ---- phase 0: every node is its own root
        for (int x = tid; x < NN; x += nt) T[x] = ((uint64_t)kCodeRoot << 32) | (uint32_t)x;
        {   // argmax(f), first in raster order: torch_topological's fake destroyer (H0); NaN check (the order is undefined)
            unsigned long long best = 0ull;
            bool has_nan = false;
            for (int p = tid; p < N; p += nt) {
                const float fv = __ldg(g.f + p);
                has_nan |= fv != fv;
                unsigned long long k = ((unsigned long long)mono32(fv) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)p);
                best = k > best ? k : best;
            }
            if (DIM == 0) atomicMax(&s_argmax, best);
            if (has_nan) s_avail = -1;
        }
        __syncthreads();
        if (s_avail < 0) {  // block-uniform
            if (tid == 0) { int av; atomicOr(A.ps.status, kStNonFinite); ps_reserve(A.ps, set, map, 0, &av); A.ps.dsum[set][map] = 0.0; }
            continue;
        }

        // ---- phase 1: level-0 contraction along each node's earliest incident edge
        const int n_real = DIM == 1 ? N : NN;
        const int n_real_pad = (n_real + 31) & ~31;
        for (int x = tid; x < n_real_pad; x += nt) {  // warp-uniform trip count
            uint64_t best = ~0ull;
            int other = -1;
            if (x >= n_real) {
            } else if (DIM == 1) {
                const int r = x / W, c = x - r * W;
                const float fp = g.px(r, c);
                // top, bottom: h-edges; left, right: v-edges
                {
                    float v = r == 0 ? fp : fminf(fp, g.px(r - 1, c));
                    uint64_t k = g.make_ekey(v, (uint32_t)(2 * c + 1 + (2 * r) * GW));
                    if (k < best) { best = k; other = r == 0 ? g.OUT : x - W; }
                }
                {
                    float v = r == H - 1 ? fp : fminf(fp, g.px(r + 1, c));
                    uint64_t k = g.make_ekey(v, (uint32_t)(2 * c + 1 + (2 * r + 2) * GW));
                    if (k < best) { best = k; other = r == H - 1 ? g.OUT : x + W; }
                }
                {
                    float v = c == 0 ? fp : fminf(fp, g.px(r, c - 1));
                    uint64_t k = g.make_ekey(v, (uint32_t)(2 * c + (2 * r + 1) * GW));
                    if (k < best) { best = k; other = c == 0 ? g.OUT : x - 1; }
                }
                {
                    float v = c == W - 1 ? fp : fminf(fp, g.px(r, c + 1));
                    uint64_t k = g.make_ekey(v, (uint32_t)(2 * c + 2 + (2 * r + 1) * GW));
                    if (k < best) { best = k; other = c == W - 1 ? g.OUT : x + 1; }
                }
                if ((uint32_t)(best >> 32) != (uint32_t)(g.make_ekey(fp, 0u) >> 32)) other = -1;  // strict local max
            } else {
                const int i = x / VW, j = x - i * VW;
                if (i > 0) {
                    uint64_t k = g.make_ekey(g.vedge_val(i - 1, j), (uint32_t)(2 * j + (2 * i - 1) * GW));
                    if (k < best) { best = k; other = x - VW; }
                }
                if (i < H) {
                    uint64_t k = g.make_ekey(g.vedge_val(i, j), (uint32_t)(2 * j + (2 * i + 1) * GW));
                    if (k < best) { best = k; other = x + VW; }
                }
                if (j > 0) {
                    uint64_t k = g.make_ekey(g.hedge_val(i, j - 1), (uint32_t)(2 * j - 1 + (2 * i) * GW));
                    if (k < best) { best = k; other = x - 1; }
                }
                if (j < W) {
                    uint64_t k = g.make_ekey(g.hedge_val(i, j), (uint32_t)(2 * j + 1 + (2 * i) * GW));
                    if (k < best) { best = k; other = x + 1; }
                }
                // a vertex always has an incident edge of its own value (min of the same pixels)
            }
            if (other >= 0) {
                bool linked = false;
                if (DIM == 1) {
                    // far end strictly higher, or equal with a larger raster index, or OUTSIDE: it and all its
                    // ancestors are elder than x, so x (still a root) links under its root without key lookups
                    bool elder_far = other == g.OUT;
                    if (!elder_far) {
                        const float fp = __ldg(g.f + x), fo = __ldg(g.f + other);
                        elder_far = fo > fp || (fo == fp && other > x);
                    }
                    if (elder_far) {
                        const int rb = ph.find0(other);
                        const unsigned long long expect = ((unsigned long long)kCodeRoot << 32) | (uint32_t)x;
                        const unsigned long long want = ((unsigned long long)kCodeL0 << 32) | (uint32_t)rb;
                        linked = atomicCAS(reinterpret_cast<unsigned long long*>(T + x), expect, want) == expect;
                    }
                }
                if (!linked) ph.union0(x, other);
            }
            __syncwarp();  // reconverge (independent thread scheduling lets the lanes drift apart otherwise)
        }
        __syncthreads();
        // flatten level-0 chains by pointer jumping (own entry only; depth halves per round)
        for (;;) {
            int changed = 0;
            for (int x = tid; x < n_real; x += nt) {
                const uint64_t t = ld_cg_u64(T + x);
                if ((uint32_t)(t >> 32) != kCodeL0) continue;
                const uint64_t tp = ld_cg_u64(T + (uint32_t)t);
                if ((uint32_t)(tp >> 32) == kCodeL0) { T[x] = tp; changed = 1; }
            }
            if (!__syncthreads_or(changed)) break;
        }

        // ---- phase 2: all edges into the triplet merge tree
        const int n_vedges = H * (W + 1), n_hedges = (H + 1) * W;
        const int n_edges_pad = (n_vedges + n_hedges + 31) & ~31;
        for (int e = tid; e < n_edges_pad; e += nt) {  // warp-uniform trip count
            if (e >= n_vedges + n_hedges) { __syncwarp(); continue; }
            int a, b;
            uint32_t pos;
            float val;
            if (e < n_vedges) {
                const int i = e / (W + 1), j = e - i * (W + 1);
                pos = (uint32_t)(2 * j + (2 * i + 1) * GW);
                val = g.vedge_val(i, j);
                if (DIM == 1) { a = j == 0 ? g.OUT : i * W + j - 1; b = j == W ? g.OUT : i * W + j; }
                else { a = i * VW + j; b = a + VW; }
            } else {
                const int e2 = e - n_vedges, i = e2 / W, j = e2 - i * W;
                pos = (uint32_t)(2 * j + 1 + (2 * i) * GW);
                val = g.hedge_val(i, j);
                if (DIM == 1) { a = i == 0 ? g.OUT : (i - 1) * W + j; b = i == H ? g.OUT : i * W + j; }
                else { a = i * VW + j; b = a + 1; }
            }
            // quick reject: same level-0 basin
            uint64_t ta = ld_cg_u64(T + a), tb = ld_cg_u64(T + b);
            int la = (uint32_t)(ta >> 32) == kCodeL0 ? (int)(uint32_t)ta : a;
            int lb = (uint32_t)(tb >> 32) == kCodeL0 ? (int)(uint32_t)tb : b;
            if (la != lb) ph.merge(la, lb, pos, g.make_ekey(val, pos));
            __syncwarp();
        }
        __syncthreads();

        // ---- phase 3: emit pairs of positive persistence.  Two passes over the nodes: count, reserve the
        //      map's records in the arena, then write them in node (= raster) order.
        auto emits = [&](int x, PairRec& rec, uint64_t& sk) {
            const uint64_t t = ld_cg_u64(T + x);
            const uint32_t code = (uint32_t)(t >> 32);
            if (code != kCodeL0 && code != kCodeRoot) {
                const uint32_t pos = code - 1;
                const uint64_t ek = g.ekey(pos), nk = g.nkey(x);
                if ((uint32_t)(ek >> 32) != (uint32_t)(nk >> 32)) {
                    if (DIM == 1) { rec.cre = g.edge_top(pos); rec.des = x; sk = ~nk; }  // death cell = square x
                    else { g.vertex_val(x / VW, x % VW, &rec.cre); rec.des = g.edge_top(pos); sk = ek; }  // death cell = edge
                    return true;
                }
            } else if (DIM == 0 && code == kCodeRoot) {  // the essential class, emitted last
                g.vertex_val(x / VW, x % VW, &rec.cre);
                rec.des = (int)(0xFFFFFFFFu - (uint32_t)s_argmax);
                sk = ~0ull;
                return true;
            }
            return false;
        };
        {
            int mine = 0;
            for (int x = tid; x < NN; x += nt) { PairRec r; uint64_t k; mine += emits(x, r, k) ? 1 : 0; }
            mine = __reduce_add_sync(0xFFFFFFFFu, mine);
            if ((tid & 31) == 0 && mine) atomicAdd(&s_count, mine);
        }
        __syncthreads();
        if (tid == 0) { int avail; s_base = ps_reserve(A.ps, set, map, s_count, &avail); s_avail = avail; }
        __syncthreads();
        PairRec* out = A.ps.arena + s_base;
        uint64_t* skeys = A.ps.skeys ? A.ps.skeys + s_base : nullptr;
        const int avail = s_avail;
        double dacc = 0.0;
        int emit_base = 0;
        for (int x0 = 0; x0 < NN; x0 += nt) {
            const int x = x0 + tid;
            bool emit = false;
            PairRec rec;
            uint64_t sk = 0;
            if (x < NN) emit = emits(x, rec, sk);
            // deterministic slots: block-wide scan in node (= raster) order, no atomics
            const unsigned ballot = __ballot_sync(0xFFFFFFFFu, emit);
            if ((tid & 31) == 0) s_wcnt[tid >> 5] = __popc(ballot);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < kPhThreads / 32; ++w) { const int v = s_wcnt[w]; total += v; if (w < (tid >> 5)) before += v; }
            if (emit) {
                const int slot = emit_base + before + __popc(ballot & lanemask_lt());
                if (slot < avail) {
                    rec.b = __ldg(g.f + rec.cre);
                    rec.d = __ldg(g.f + rec.des);
                    rec.tb = rec.td = __int_as_float(0x7FC00000);
                    out[slot] = rec;
                    if (skeys) skeys[slot] = sk;
                    dacc += (double)cost_diag(rec.b, rec.d, A.ps.q);
                }
            }
            emit_base += total;
            __syncthreads();
        }
        dacc = block_sum(dacc, s_red);
        if (tid == 0) A.ps.dsum[set][map] = dacc;
    }
}

}  // namespace tl
