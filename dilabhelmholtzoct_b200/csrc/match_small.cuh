// Diagram matching of the forward path (tl_forward): see below.  Included by ph_small.cuh as well -- the
// persistence kernel runs the matching of finished maps in the tail of its launch, on SMs that have run out
// of persistence work.
#pragma once
#include "tl_common.cuh"

namespace tl {

constexpr int kMatchThreads = 1024;

// ---- forward path of tl_forward: one pass over the maps.
//
// Segmentation ground truth has 0-5 pairs per map, so the assignment is almost always trivial:
//   * min(n, m) == 0   every point goes to the diagonal; the cost is the two sums the persistence kernel
//                      already formed while emitting (PairStore::dsum) -- no record is read at all;
//   * min(n, m) <= 8   shortest augmenting paths with ALL state in shared memory.  The column potentials
//                      are non-zero only for the <= R(R+1) columns that were ever on a search tree, and the
//                      running column minima of a phase are recomputed from the <= R+1 tree rows instead of
//                      stored, so nothing per column is kept: a step is one sweep over the columns plus one
//                      block-wide arg-min.  Same algorithm, same tie-breaks (smallest column, earliest tree
//                      row) as match_kernel and the oracle;
//   * otherwise        the map is appended to a list that match_kernel (global scratch, a few slots)
//                      works through afterwards.
constexpr int kSmallR = 8;
constexpr int kTouchMax = kSmallR * (kSmallR + 1) + 8;

struct MatchFwdArgs {
    PairStore ps;
    int n_maps, loss_r;
    float q;
    double* cost;    // [n_maps]
    double* tpers;   // [n_maps]
    int32_t* heavy;  // [n_maps] maps left to match_kernel
    unsigned int* n_heavy;
    unsigned int* counter;
};

constexpr int kColCache = 12288;  // columns (b, d) kept in dynamic shared memory: 96 KB

// one map; every thread of a kMatchThreads-wide CTA calls this (block-uniform); s_col: kColCache points of scratch
// returns (block-uniform) whether the map went to the heavy list instead of being matched here
__device__ __noinline__ bool match_one_map(const MatchFwdArgs& A, int k, float2* s_col) {
    __shared__ double s_red[kMatchThreads / 32];
    __shared__ double s_bv[kMatchThreads / 32];
    __shared__ int s_bk[kMatchThreads / 32], s_bw[kMatchThreads / 32];
    __shared__ int s_done, s_nU, s_nT;
    __shared__ float s_rb[kSmallR], s_rd[kSmallR];
    __shared__ double s_rdiag[kSmallR], s_u[kSmallR + 1], s_tv[kTouchMax];
    __shared__ int s_Ucol[kSmallR + 2], s_Urow[kSmallR + 2], s_tc[kTouchMax], s_tp[kTouchMax], s_tway[kTouchMax];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const float q = A.q;
    const double kInf = __longlong_as_double(0x7FF0000000000000LL);
    __syncthreads();  // the previous map's readers of the shared state are done
    const int n = A.ps.counts[0][k], m = A.ps.counts[1][k];
    PairRec* rec1 = A.ps.arena + A.ps.offs[0][k];
    const PairRec* rec2 = A.ps.arena + A.ps.offs[1][k];
    double tp = 0.0;
    if (A.loss_r) {  // total persistence of the prediction diagram (topological_loss.py:88-94)
        for (int i = tid; i < n; i += nt) tp += pow(fabs((double)rec1[i].d - (double)rec1[i].b), (double)q);
        tp = block_sum(tp, s_red);
    }
    const bool swp = n < m;
    const int R = swp ? n : m, Cn = swp ? m : n;
    if (R == 0 || R > kSmallR) {
        if (tid == 0) {
            A.tpers[k] = tp;
            if (R == 0) A.cost[k] = A.ps.dsum[0][k] + A.ps.dsum[1][k];
            else A.heavy[atomicAdd(A.n_heavy, 1u)] = k;
        }
        return R != 0;
    }
    const PairRec* rR = swp ? rec1 : rec2;
    const PairRec* rC = swp ? rec2 : rec1;
    const int NC = Cn + R;
    if (tid < R) {
        const float b = rR[tid].b, d = rR[tid].d;
        s_rb[tid] = b; s_rd[tid] = d; s_rdiag[tid] = (double)cost_diag(b, d, q);
    }
    if (tid <= R) s_u[tid] = 0.0;
    if (tid == 0) s_nT = 0;
    // the columns are read once per step of every phase: keep them on chip (a step is then a shared-memory
    // sweep instead of a round trip to L2)
    const bool cached = Cn <= kColCache;
    if (cached) for (int c = tid; c < Cn; c += nt) s_col[c] = make_float2(rC[c].b, rC[c].d);
    __syncthreads();
    for (int r = 1; r <= R; ++r) {
        if (tid == 0) { s_nU = 1; s_Ucol[0] = 0; s_Urow[0] = r; }
        __syncthreads();
        for (;;) {
            const int nU = s_nU, nT = s_nT;
            double best = kInf; int bestk = 0x7FFFFFFF, bestw = 0;
            for (int c = 1 + tid; c <= NC; c += nt) {
                bool used = false;
                for (int t = 1; t < nU; ++t) used |= s_Ucol[t] == c;
                if (used) continue;
                double vc = 0.0;
                for (int t = 0; t < nT; ++t) if (s_tc[t] == c) vc = s_tv[t];
                float cb = 0.f, cd = 0.f; double cdg = 0.0;
                const bool real = c <= Cn;
                if (real) {
                    if (cached) { const float2 p = s_col[c - 1]; cb = p.x; cd = p.y; } else { cb = rC[c - 1].b; cd = rC[c - 1].d; }
                    cdg = (double)cost_diag(cb, cd, q);
                }
                double mv = kInf; int way = 0;
                for (int t = 0; t < nU; ++t) {  // tree rows in the order they joined: the first strict minimum wins
                    const int i = s_Urow[t] - 1;
                    const double cc = real ? (double)cost_pp(s_rb[i], s_rd[i], cb, cd, q) - cdg : s_rdiag[i];
                    const double cur = cc - s_u[i + 1] - vc;
                    if (cur < mv) { mv = cur; way = s_Ucol[t]; }
                }
                if (mv < best) { best = mv; bestk = c; bestw = way; }  // c ascending per thread: first minimum kept
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_down_sync(0xFFFFFFFFu, best, o);
                const int ok = __shfl_down_sync(0xFFFFFFFFu, bestk, o), ow = __shfl_down_sync(0xFFFFFFFFu, bestw, o);
                if (ob < best || (ob == best && ok < bestk)) { best = ob; bestk = ok; bestw = ow; }
            }
            if (lane == 0) { s_bv[warp] = best; s_bk[warp] = bestk; s_bw[warp] = bestw; }
            __syncthreads();
            if (tid == 0) {
                best = s_bv[0]; bestk = s_bk[0]; bestw = s_bw[0];
                for (int w = 1; w < nt / 32; ++w)
                    if (s_bv[w] < best || (s_bv[w] == best && s_bk[w] < bestk)) { best = s_bv[w]; bestk = s_bk[w]; bestw = s_bw[w]; }
                const double delta = best; const int j1 = bestk;
                for (int t = 0; t < nU; ++t) s_u[s_Urow[t]] += delta;
                for (int t = 1; t < nU; ++t) {
                    const int c = s_Ucol[t];
                    for (int x = 0; x < nT; ++x) if (s_tc[x] == c) s_tv[x] -= delta;
                }
                int tj = -1;
                for (int x = 0; x < nT; ++x) if (s_tc[x] == j1) tj = x;
                if (tj < 0) { tj = nT; s_tc[tj] = j1; s_tv[tj] = 0.0; s_tp[tj] = 0; s_nT = nT + 1; }
                s_tway[tj] = bestw;
                if (s_tp[tj] == 0) {  // free column: flip the path back to the phase's row
                    int j = j1;
                    while (j != 0) {
                        int xj = 0; for (int x = 0; x < s_nT; ++x) if (s_tc[x] == j) xj = x;
                        const int jp = s_tway[xj];
                        int pr = r;
                        if (jp != 0) for (int x = 0; x < s_nT; ++x) if (s_tc[x] == jp) pr = s_tp[x];
                        s_tp[xj] = pr;
                        j = jp;
                    }
                    s_done = 1;
                } else {
                    s_Ucol[nU] = j1; s_Urow[nU] = s_tp[tj]; s_nU = nU + 1;
                    s_done = 0;
                }
            }
            __syncthreads();
            if (s_done) break;
        }
    }
    // cost of the optimum: every column goes to the diagonal (dsum) unless a row took it
    if (tid == 0) {
        double total = A.ps.dsum[swp ? 1 : 0][k];
        for (int x = 0; x < s_nT; ++x) {
            const int row = s_tp[x] - 1, c = s_tc[x];
            if (row < 0) continue;
            if (c <= Cn) {
                const float cb = rC[c - 1].b, cd = rC[c - 1].d;
                total += (double)cost_pp(s_rb[row], s_rd[row], cb, cd, q) - (double)cost_diag(cb, cd, q);
                const int i1 = swp ? row : c - 1;  // record of the prediction in this match
                rec1[i1].tb = swp ? cb : s_rb[row]; rec1[i1].td = swp ? cd : s_rd[row];
            } else total += s_rdiag[row];
        }
        A.cost[k] = total;
        A.tpers[k] = tp;
    }
    return false;
}

__global__ void __launch_bounds__(kMatchThreads) match_small_kernel(const __grid_constant__ MatchFwdArgs A) {
    extern __shared__ __align__(8) float2 s_col_dyn[];  // [kColCache] points of the larger diagram
    __shared__ int s_k;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_k = (int)atomicAdd(A.counter, 1u);
        __syncthreads();
        const int k = s_k;
        if (k >= A.n_maps) break;
        match_one_map(A, k, s_col_dyn);
    }
}

}  // namespace tl
