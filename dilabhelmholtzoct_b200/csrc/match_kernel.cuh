// Diagram matching + loss + gradient scatter.
//
// match_kernel   exact optimum of the transport problem torch_topological's WassersteinDistance
//                hands to POT ot.emd2 (reference call /root/reference/octsam/models/
//                topological_loss.py:78-82): an (n+1)x(m+1) cost matrix with a diagonal row/column,
//                masses [1]*n+[m] / [1]*m+[n].  That LP is a partial assignment, solved here by
//                shortest augmenting paths with the SMALLER diagram as rows (truth diagrams are tiny),
//                one CTA per (image, class) map, columns processed in parallel.
// loss_kernel    S_b = sum_c cost, W_b = S_b^(1/q), loss = lamda * mean_b W_b (+ loss_r term)
//                (topological_loss.py:85-96) and the per-image backward coefficient.
// grad_kernel    analytic backward of training_utils.py:66 through pow / emd2 / cdist(p=inf) /
//                vector_norm(inf) / gather: scatter-add into the critical pixels.
#pragma once
#include <math.h>
#include "tl_common.cuh"
#include "grad_kernel.cuh"
#include "match_small.cuh"

namespace tl {


struct Diagrams {          // strided view of (birth, death) rows
    const char* base;      // first row of diagram 0
    int stride;            // bytes between rows
    const int32_t* off;     // row offsets [n_diag+1] (tl_wasserstein), or null
    const uint32_t* start;  // first row of diagram k (pair arena; used with count when off == null)
    const int32_t* count;   // row counts [n_diag]
    __device__ __forceinline__ int rows(int k) const { return off ? off[k + 1] - off[k] : count[k]; }
    __device__ __forceinline__ size_t first_row(int k) const { return off ? (size_t)off[k] : (size_t)start[k]; }
    __device__ __forceinline__ const char* first(int k) const { return base + first_row(k) * stride; }
};
__device__ __forceinline__ float2 row_at(const char* first, int stride, int i) {
    return *reinterpret_cast<const float2*>(first + (size_t)i * stride);
}

struct MatchArgs {
    Diagrams d1, d2;
    int n_diag;
    float q;
    int loss_r;
    double* cost;        // [n_diag]
    double* tpers;       // [n_diag] or null
    int32_t* match1;     // rows of d1 (same offsets as d1): matched row of d2 or -1; may be null
    PairRec* fill1;      // when non-null: the records behind d1; matched truth points go to .tb / .td
    const int32_t* list; // when non-null: the diagrams to process are list[0 .. *n_list)
    const unsigned int* n_list;
    // per-CTA scratch
    double* v; double* minv; double* u;
    int32_t* way; int32_t* pcol; uint8_t* used;
    size_t stride_c, stride_r;
    unsigned int* counter;  // dynamic work counter (zeroed by the caller) or null: static round-robin over CTAs
};

__global__ void __launch_bounds__(kMatchThreads) match_kernel(MatchArgs A) {
    __shared__ double s_red[kMatchThreads / 32];
    __shared__ double s_bv[kMatchThreads / 32];
    __shared__ int s_bk[kMatchThreads / 32];
    __shared__ int s_j0, s_done;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    double* v = A.v + (size_t)blockIdx.x * A.stride_c;
    double* minv = A.minv + (size_t)blockIdx.x * A.stride_c;
    int32_t* way = A.way + (size_t)blockIdx.x * A.stride_c;
    int32_t* pcol = A.pcol + (size_t)blockIdx.x * A.stride_c;
    uint8_t* used = A.used + (size_t)blockIdx.x * A.stride_c;
    double* u = A.u + (size_t)blockIdx.x * A.stride_r;
    const float q = A.q;
    const double kInf = __longlong_as_double(0x7FF0000000000000LL);

    __shared__ int s_k;
    // maps with a non-empty ground-truth diagram cost 10-100x the others: hand the maps out dynamically
    const int n_work = A.list ? (int)*A.n_list : A.n_diag;
    for (int w = blockIdx.x;; w += gridDim.x) {  // w: position in the work list, k: the diagram it names
        if (A.counter) {
            __syncthreads();
            if (tid == 0) s_k = (int)atomicAdd(A.counter, 1u);
            __syncthreads();
            w = s_k;
        }
        if (w >= n_work) break;
        const int k = A.list ? A.list[w] : w;
        const int n = A.d1.rows(k), m = A.d2.rows(k);
        const char* r1 = A.d1.first(k);
        const char* r2 = A.d2.first(k);
        const int st1 = A.d1.stride, st2 = A.d2.stride;
        const bool swp = n < m;
        const int R = swp ? n : m, Cn = swp ? m : n;
        const char* rR = swp ? r1 : r2; const int stR = swp ? st1 : st2;
        const char* rC = swp ? r2 : r1; const int stC = swp ? st2 : st1;
        int32_t* match1 = A.match1 ? A.match1 + A.d1.first_row(k) : nullptr;
        PairRec* recs1 = A.fill1 ? A.fill1 + A.d1.first_row(k) : nullptr;

        double part = 0.0, tp = 0.0;
#pragma unroll 4
        for (int c = tid; c < Cn; c += nt) { float2 p = row_at(rC, stC, c); part += (double)cost_diag(p.x, p.y, q); }
        if (match1) for (int i = tid; i < n; i += nt) match1[i] = -1;
        if (A.loss_r)
            for (int i = tid; i < n; i += nt) { float2 p = row_at(r1, st1, i); tp += pow(fabs((double)p.y - (double)p.x), (double)q); }
        double total = block_sum(part, s_red);
        if (A.loss_r) tp = block_sum(tp, s_red);

        if (R > 0) {
            const int NC = Cn + R;
            for (int c = tid; c <= NC; c += nt) { v[c] = 0.0; pcol[c] = 0; }
            for (int r = tid; r <= R; r += nt) u[r] = 0.0;
            __syncthreads();
            for (int r = 1; r <= R; ++r) {
                for (int c = tid; c <= NC; c += nt) { minv[c] = kInf; used[c] = 0; }
                if (tid == 0) { pcol[0] = r; s_j0 = 0; }
                __syncthreads();
                for (;;) {
                    const int j0 = s_j0, i0 = pcol[j0];
                    const float2 pr = row_at(rR, stR, i0 - 1);
                    const double ui = u[i0];
                    const double crd = (double)cost_diag(pr.x, pr.y, q);
                    __syncthreads();
                    if (tid == 0) used[j0] = 1;
                    __syncthreads();
                    double best = kInf; int bestk = 0x7FFFFFFF;
                    for (int c = 1 + tid; c <= NC; c += nt) {
                        if (used[c]) continue;
                        double cc;
                        if (c <= Cn) {
                            float2 pc = row_at(rC, stC, c - 1);
                            cc = (double)cost_pp(pr.x, pr.y, pc.x, pc.y, q) - (double)cost_diag(pc.x, pc.y, q);
                        } else cc = crd;
                        const double cur = cc - ui - v[c];
                        double mv = minv[c];
                        if (cur < mv) { mv = cur; minv[c] = cur; way[c] = j0; }
                        if (mv < best) { best = mv; bestk = c; }  // c ascending per thread: first minimum kept
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        double ob = __shfl_down_sync(0xFFFFFFFFu, best, o);
                        int ok = __shfl_down_sync(0xFFFFFFFFu, bestk, o);
                        if (ob < best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
                    }
                    if (lane == 0) { s_bv[warp] = best; s_bk[warp] = bestk; }
                    __syncthreads();
                    best = s_bv[0]; bestk = s_bk[0];
                    for (int w = 1; w < nt / 32; ++w)
                        if (s_bv[w] < best || (s_bv[w] == best && s_bk[w] < bestk)) { best = s_bv[w]; bestk = s_bk[w]; }
                    const double delta = best; const int j1 = bestk;
                    for (int c = tid; c <= NC; c += nt) {
                        if (used[c]) { u[pcol[c]] += delta; v[c] -= delta; }
                        else minv[c] -= delta;
                    }
                    __syncthreads();
                    if (tid == 0) { s_j0 = j1; s_done = pcol[j1] == 0; }
                    __syncthreads();
                    if (s_done) break;
                }
                if (tid == 0) {
                    int j0 = s_j0;
                    do { const int j1 = way[j0]; pcol[j0] = pcol[j1]; j0 = j1; } while (j0);
                }
                __syncthreads();
            }
            double part2 = 0.0;
            for (int c = 1 + tid; c <= NC; c += nt) {
                const int r = pcol[c] - 1;
                if (r < 0) continue;
                const float2 pr = row_at(rR, stR, r);
                if (c <= Cn) {
                    const float2 pc = row_at(rC, stC, c - 1);
                    part2 += (double)cost_pp(pr.x, pr.y, pc.x, pc.y, q) - (double)cost_diag(pc.x, pc.y, q);
                    const int i1 = swp ? r : c - 1;  // row of d1 / of d2 in this match
                    if (match1) match1[i1] = swp ? c - 1 : r;
                    if (recs1) { const float2 p2 = swp ? pc : pr; recs1[i1].tb = p2.x; recs1[i1].td = p2.y; }
                } else part2 += (double)cost_diag(pr.x, pr.y, q);
            }
            total += block_sum(part2, s_red);
        }
        __syncthreads();
        if (tid == 0) { A.cost[k] = total; if (A.tpers) A.tpers[k] = A.loss_r ? tp : 0.0; }
        __syncthreads();
    }
}

struct LossArgs {
    const double* cost; const double* tpers;
    int B, C, B_global, loss_r;
    float q, lamda;
    float* loss_out;
    double* coef;  // [B] d loss / d S_b (NaN when S_b == 0, as autograd's 0 * inf)
    const unsigned int* status;  // device status word: any bit set poisons the loss with NaN
};

__global__ void __launch_bounds__(256) loss_kernel(LossArgs A) {
    __shared__ double s_red[8];
    const int tid = threadIdx.x;
    const int Bg = A.B_global > 0 ? A.B_global : A.B;
    double acc = 0.0, reg = 0.0;
    for (int b = tid; b < A.B; b += blockDim.x) {
        float S;  // total_cost += emd2(...) runs in fp32 in the reference
        A.coef[b] = image_coef(A.cost, b, A.C, A.q, A.lamda, Bg, &S);
        const float Wb = A.q == 2.0f ? sqrtf(S) : powf(S, 1.0f / A.q);
        acc += (double)Wb;
        if (A.loss_r) for (int c = 0; c < A.C; ++c) reg += A.tpers[b * A.C + c];
    }
    acc = block_sum(acc, s_red);
    if (A.loss_r) reg = block_sum(reg, s_red);
    if (tid == 0) {
        double loss = acc / Bg;
        if (A.loss_r) loss += reg / ((double)Bg * A.C);
        // an exhausted arena / basin table or a NaN pixel makes the pairing incomplete: fail loudly
        *A.loss_out = (A.status && *A.status) ? __int_as_float(0x7FC00000) : (float)((double)A.lamda * loss);
    }
}

}  // namespace tl
