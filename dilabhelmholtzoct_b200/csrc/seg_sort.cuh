// Batched segmented LSD radix sort: one CTA per (image, class) map sorts that map's persistence
// pairs by the filtration key of their death cell, i.e. into the order gudhi emits them
// (cofaces_of_persistence_pairs, reached from /root/reference/octsam/models/topological_loss.py:62).
// The persistence kernels emit a map's pairs in raster order of the dying basin (deterministic, and
// what the loss path uses as is); this sort only serves tl_persistence_pairs, whose contract is gudhi's
// emission order.
//
// 8-bit digits over 64-bit keys; passes whose digit is constant over the segment are skipped.
// Stability: every warp owns a contiguous chunk of the segment and walks it in order; within a
// 32-key group, __match_any_sync ranks equal digits by lane.
#pragma once
#include "tl_common.cuh"

namespace tl {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;

struct SortArgs {
    PairStore ps;        // records and sort keys of every map (arena + per-map offsets / counts)
    int n_sets, n_maps;
    int cap;             // most pairs one map can have (scratch stride)
    // per-CTA scratch
    uint64_t* key_tmp;   // [grid][cap]
    uint32_t* idx_a;     // [grid][cap]
    uint32_t* idx_b;     // [grid][cap]
    PairRec* rec_tmp;    // [grid][cap]
};

__global__ void __launch_bounds__(kSortThreads) seg_sort_kernel(SortArgs A) {
    __shared__ uint32_t s_hist[kSortWarps][256];
    __shared__ uint32_t s_total[256];
    __shared__ int s_skip;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_jobs = A.n_sets * A.n_maps;
    uint64_t* kt = A.key_tmp + (size_t)blockIdx.x * A.cap;
    uint32_t* ia = A.idx_a + (size_t)blockIdx.x * A.cap;
    uint32_t* ib = A.idx_b + (size_t)blockIdx.x * A.cap;
    PairRec* rt = A.rec_tmp + (size_t)blockIdx.x * A.cap;

    for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const int set = job % A.n_sets, map = job / A.n_sets;
        int n = A.ps.counts[set][map];
        if (n > A.cap) n = A.cap;
        if (n <= 1) continue;
        PairRec* recs = A.ps.arena + A.ps.offs[set][map];
        uint64_t* k0 = A.ps.skeys + A.ps.offs[set][map];
        uint64_t* ksrc = k0; uint64_t* kdst = kt;
        uint32_t* isrc = ia; uint32_t* idst = ib;
        for (int i = tid; i < n; i += kSortThreads) ia[i] = (uint32_t)i;
        // contiguous chunk per warp, multiple of 32
        const int chunk = ((n + kSortWarps - 1) / kSortWarps + 31) & ~31;
        const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
        __syncthreads();

        for (int shift = 0; shift < 64; shift += 8) {
            for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_hist[0][0])[i] = 0u;
            __syncthreads();
            for (int i = beg + lane; i < end; i += 32)
                atomicAdd(&s_hist[warp][(uint32_t)(ksrc[i] >> shift) & 255u], 1u);
            __syncthreads();
            if (tid < 256) {  // column prefix over warps
                uint32_t run = 0;
                for (int w = 0; w < kSortWarps; ++w) { uint32_t h = s_hist[w][tid]; s_hist[w][tid] = run; run += h; }
                s_total[tid] = run;
            }
            if (tid == 0) s_skip = 0;
            __syncthreads();
            if (tid < 256 && s_total[tid] == (uint32_t)n) s_skip = 1;  // constant digit: nothing to do
            __syncthreads();
            if (s_skip) continue;  // uniform across the block
            if (warp == 0) {  // exclusive scan of the 256 digit totals
                uint32_t v[8], sum = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { v[k] = s_total[lane * 8 + k]; sum += v[k]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
                uint32_t run = incl - sum;
#pragma unroll
                for (int k = 0; k < 8; ++k) { s_total[lane * 8 + k] = run; run += v[k]; }
            }
            __syncthreads();
            for (int i0 = beg; i0 < end; i0 += 32) {
                const int i = i0 + lane;
                const bool valid = i < end;
                uint64_t key = 0; uint32_t idx = 0; uint32_t d = 256u + (uint32_t)lane;
                if (valid) { key = ksrc[i]; idx = isrc[i]; d = (uint32_t)(key >> shift) & 255u; }
                const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
                const int rank = __popc(peers & lanemask_lt());
                if (valid) {
                    const uint32_t pos = s_total[d] + s_hist[warp][d] + (uint32_t)rank;
                    kdst[pos] = key; idst[pos] = idx;
                }
                __syncwarp();
                if (valid && rank == 0) s_hist[warp][d] += (uint32_t)__popc(peers);
                __syncwarp();
            }
            __syncthreads();
            { uint64_t* t = ksrc; ksrc = kdst; kdst = t; }
            { uint32_t* t = isrc; isrc = idst; idst = t; }
        }
        // gather the records into sorted order
        for (int i = tid; i < n; i += kSortThreads) rt[i] = recs[isrc[i]];
        __syncthreads();
        for (int i = tid; i < n; i += kSortThreads) recs[i] = rt[i];
        if (ksrc != k0) for (int i = tid; i < n; i += kSortThreads) k0[i] = ksrc[i];
        __syncthreads();
    }
}

}  // namespace tl
