// Backward of the loss: d loss / d pred scattered into the critical pixels (analytic backward of the reference's
// autograd graph, /root/reference/octsam/models/training_utils.py:66 restricted to topo_loss).  The per-map body is
// shared by grad_kernel (tl_backward) and by the tail of the persistence launch (tl_forward_backward, ph_small.cuh).
#pragma once
#include "tl_common.cuh"

namespace tl {

struct GradArgs {
    const PairRec* arena; const uint32_t* offs; const int32_t* counts; const double* coef; const float* grad_loss;
    int M, C, N, B_global, loss_r;
    float q, lamda;
    float* grad_pred;
    const uint32_t* gfused;  // [B] or null: images whose gradient the persistence launch already wrote (tl_forward_backward)
};

// d loss / d S_b of image b (NaN when S_b == 0, as autograd's 0 * inf); S_b is summed in fp32 like the reference's
// `total_cost += emd2(...)`.  One definition for loss_kernel and the gradient fused into the persistence launch.
__device__ __forceinline__ double image_coef(const double* cost, int b, int C, float q, float lamda, int Bg, float* S_out) {
    float S = 0.f;
    for (int c = 0; c < C; ++c) S += (float)cost[b * C + c];
    if (S_out) *S_out = S;
    return S > 0.f ? (double)lamda / Bg * (1.0 / q) * pow((double)S, 1.0 / (double)q - 1.0)
                   : __longlong_as_double(0x7FF8000000000000LL);
}

// d loss / d (birth, death) of one pair, upstream gradient and the image's coefficient included
__device__ __forceinline__ void pair_gradient(const GradArgs& A, const PairRec& r, double coef, double creg, double& gb, double& gd) {
    const double q = (double)A.q;
    if (A.q == 2.0f && !A.loss_r && isnan(r.tb)) {
        // the usual case, with the fp64 work cut to one multiplication: -0.5 * (2 x) * s is exactly -(s x), so this
        // is bit-identical to the general branch below
        const float h = 0.5f * (r.b + r.d);
        const float x = fmaxf(fabsf(r.b - h), fabsf(r.d - h));
        const double t = (double)(r.d > r.b ? x : (r.d < r.b ? -x : 0.f)) * coef;
        gb = -t; gd = t;
        return;
    }
    if (isnan(r.tb)) {  // matched to the diagonal
        const float h = 0.5f * (r.b + r.d);
        const float x = fmaxf(fabsf(r.b - h), fabsf(r.d - h));
        const double gg = x > 0.f ? (q == 2.0 ? 2.0 * (double)x : q * pow((double)x, q - 1.0)) : (q == 1.0 ? 1.0 : 0.0);
        const double s = r.d > r.b ? 1.0 : (r.d < r.b ? -1.0 : 0.0);
        gb = -0.5 * gg * s; gd = 0.5 * gg * s;
    } else {
        const float xb = r.b - r.tb, xd = r.d - r.td, ab = fabsf(xb), ad = fabsf(xd);
        const float dist = fmaxf(ab, ad);
        const double gg = dist > 0.f ? q * pow((double)dist, q - 1.0) : 0.0;
        gb = ab == dist ? gg * (xb > 0.f ? 1.0 : (xb < 0.f ? -1.0 : 0.0)) : 0.0;
        gd = ad == dist ? gg * (xd > 0.f ? 1.0 : (xd < 0.f ? -1.0 : 0.0)) : 0.0;
    }
    gb *= coef; gd *= coef;  // NaN coefficient (S_b == 0) poisons every entry, like autograd
    if (A.loss_r) {
        const double pers = (double)r.d - (double)r.b, ap = fabs(pers);
        const double gr = ap > 0.0 ? q * pow(ap, q - 1.0) * (pers > 0.0 ? 1.0 : -1.0) * creg : 0.0;
        gb -= gr; gd += gr;
    }
}

// One map, one CTA: zero-fill the map's gradient (128-bit stores, the lines stay in L2), then scatter-add
// the pairs into the critical pixels.  Fusing the fill keeps the atomics off cold DRAM lines and saves the
// separate memset pass.  `coef` already carries the upstream gradient.
__device__ __forceinline__ void grad_one_map(const GradArgs& A, int map, double coef, double gl) {
    const int n = A.counts[map];
    const PairRec* recs = A.arena + A.offs[map];
    float* g = A.grad_pred + (size_t)map * A.N;
    if (((reinterpret_cast<uintptr_t>(g) & 15) == 0) && (A.N & 3) == 0) {
        float4* g4 = reinterpret_cast<float4*>(g);
        for (int i = threadIdx.x; i < (A.N >> 2); i += blockDim.x) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int i = threadIdx.x; i < A.N; i += blockDim.x) g[i] = 0.f;
    }
    __syncthreads();  // the fill is visible to the whole CTA before its atomics land on the same lines
    const double creg = (double)A.lamda / ((double)A.B_global * A.C) * gl;
    // 4 records per thread and trip, loaded first: the records come from DRAM (they were streamed past L2), and a
    // thread that waits for them one at a time makes the map's gradient a chain of DRAM latencies
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {
        PairRec r4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < n) r4[u] = recs[i];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i >= n) continue;
            double gb, gd;
            pair_gradient(A, r4[u], coef, creg, gb, gd);
            atomicAdd(g + r4[u].cre, (float)gb);
            atomicAdd(g + r4[u].des, (float)gd);
        }
    }
    __syncthreads();
}

// The same map through a TILE in shared memory (the tail of the persistence launch has the SM's shared memory to
// itself): per tile of `tile_px` pixels, zero the tile, add the pairs whose pixels fall into it (shared-memory
// atomics), stream the tile out with 128-bit stores.  No global atomics and no separate zero fill: 40 -> 27 us per
// 256 x 256 map on one SM.  The records are read once per tile (from L2 after the first), so this pays for maps of
// a few tiles only.  Measured and dropped (scripts/gpu_r2f.sh): parking the second tile's contributions in shared
// memory to read the records once (39 us), plain read-modify-writes behind an ownership vote instead of atomics
// (32 us), the 8 adds of a round as one batch of compare-and-swap attempts (37 us), prefetching the next job's
// records into L2 (34 us).
__device__ __forceinline__ void grad_one_map_tiled(const GradArgs& A, int map, double coef, double gl, float* tile, int tile_px,
                                                   unsigned long long* dbg = nullptr) {  // dbg (thread 0): cycles in [zero, math + adds, stream out, record wait]
    long long tc = dbg ? clock64() : 0;
#define TL_GDBG(slot) do { if (dbg && threadIdx.x == 0) { const long long t1_ = clock64(); dbg[slot] += (unsigned long long)(t1_ - tc); tc = t1_; } } while (0)
    const int n = A.counts[map], N = A.N;
    const PairRec* recs = A.arena + A.offs[map];
    float* g = A.grad_pred + (size_t)map * N;
    const double creg = (double)A.lamda / ((double)A.B_global * A.C) * gl;
    for (int base = 0; base < N; base += tile_px) {
        const int len = min(tile_px, N - base);
        for (int i = threadIdx.x; i < ((len + 3) >> 2); i += blockDim.x) reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        TL_GDBG(0);
        for (int i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {
            PairRec r4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < n) r4[u] = recs[i];
            }
            if (dbg && threadIdx.x == 0) {  // wait for the records, charged to slot 3
                int sink = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) if (i0 + u * (int)blockDim.x < n) sink += r4[u].cre;
                if (sink == 0x7FFFFFFF) dbg[0] += 1;
                TL_GDBG(3);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i >= n) continue;
                const unsigned int pc = (unsigned int)(r4[u].cre - base), pd = (unsigned int)(r4[u].des - base);
                if (pc >= (unsigned int)len && pd >= (unsigned int)len) continue;
                double gb, gd;
                pair_gradient(A, r4[u], coef, creg, gb, gd);
                if (pc < (unsigned int)len) atomicAdd(tile + pc, (float)gb);
                if (pd < (unsigned int)len) atomicAdd(tile + pd, (float)gd);
            }
            TL_GDBG(1);
        }
        __syncthreads();
        TL_GDBG(1);
        if (((reinterpret_cast<uintptr_t>(g + base) & 15) == 0) && (len & 3) == 0) {
            float4* g4 = reinterpret_cast<float4*>(g + base);
            for (int i = threadIdx.x; i < (len >> 2); i += blockDim.x) g4[i] = reinterpret_cast<const float4*>(tile)[i];
        } else {
            for (int i = threadIdx.x; i < len; i += blockDim.x) g[base + i] = tile[i];
        }
        __syncthreads();
        TL_GDBG(2);
    }
#undef TL_GDBG
}

// ---- 1-D bulk asynchronous copies (the TMA engine without a tensor map) and the mbarrier they complete on
__device__ __forceinline__ void mbar_init(uint32_t bar_s, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar_s, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar_s, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar_s), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void bulk_load(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar_s) {  // global -> shared
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_s, uint32_t bytes) {  // shared -> global
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kGradChunk = 1920;                                   // records per staging buffer
constexpr int kGradStageBytes = kGradChunk * (int)sizeof(PairRec);  // 46080: a multiple of 16 and of two records

// grad_one_map_tiled with the records STREAMED through two staging buffers by bulk asynchronous copies (mbarrier
// completion) instead of loaded into registers round by round -- the job no longer waits a DRAM latency per round of
// four records, only for its first chunk: 31 -> 23 us per 256 x 256 map -- and each finished tile written out by bulk
// copies shared -> global.
//   stage: 2 x kGradStageBytes of shared memory, 16-byte aligned;  bars_s: shared address of two mbarriers (count 1),
//   initialised once per kernel;  phase: their parity bits, returned updated (block-uniform; every thread waits).
// A bulk copy wants 16-byte aligned ends and records are 24 bytes: the stream starts one record early when the
// map's first record has an odd index and ends on an even count (the arena keeps one spare record for that).
// Thread 0 must call bulk_wait_all() before it leaves the kernel (its bulk stores are then in memory).
__device__ __forceinline__ unsigned int grad_one_map_tiled_bulk(const GradArgs& A, int map, double coef, double gl, float* tile, int tile_px,
                                                                unsigned char* stage, uint32_t bars_s, unsigned int phase,
                                                                unsigned long long* dbg = nullptr) {  // dbg: cycles in [zero, math + adds, stream out, record wait]
    long long tc = dbg ? clock64() : 0;
#define TL_GDBG(slot) do { if (dbg && threadIdx.x == 0) { const long long t1_ = clock64(); dbg[slot] += (unsigned long long)(t1_ - tc); tc = t1_; } } while (0)
    const int n = A.counts[map], N = A.N, nt = blockDim.x;
    const unsigned int r0 = A.offs[map];
    const int skip = (int)(r0 & 1u), total = n + skip;
    const PairRec* src = A.arena + (r0 - (unsigned int)skip);
    const int nch = (total + kGradChunk - 1) / kGradChunk;
    float* g = A.grad_pred + (size_t)map * N;
    const double creg = (double)A.lamda / ((double)A.B_global * A.C) * gl;
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage), tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    auto issue = [&](int k) {  // thread 0: chunk k of the record stream into buffer k & 1
        const int cnt = min(kGradChunk, total - k * kGradChunk);
        const uint32_t bytes = (uint32_t)((cnt + 1) & ~1) * (uint32_t)sizeof(PairRec);
        const uint32_t bar = bars_s + 8u * (uint32_t)(k & 1);
        fence_async_proxy();  // the buffer's last readers (generic proxy) are behind a barrier; order them before the async write
        mbar_arrive_expect_tx(bar, bytes);
        bulk_load(stage_s + (uint32_t)(k & 1) * (uint32_t)kGradStageBytes, src + (size_t)k * kGradChunk, bytes, bar);
    };
    for (int base = 0; base < N; base += tile_px) {
        const int len = min(tile_px, N - base);
        if (threadIdx.x == 0) { if (nch > 0) issue(0); if (nch > 1) issue(1); }
        for (int i = threadIdx.x; i < ((len + 3) >> 2); i += nt) reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        TL_GDBG(0);
        for (int k = 0; k < nch; ++k) {
            const int b = k & 1;
            while (!mbar_try_wait(bars_s + 8u * (uint32_t)b, (phase >> b) & 1u)) {}
            phase ^= 1u << b;
            TL_GDBG(3);
            const PairRec* buf = reinterpret_cast<const PairRec*>(stage + (size_t)b * kGradStageBytes);
            const int cnt = min(kGradChunk, total - k * kGradChunk);
            for (int j = threadIdx.x; j < cnt; j += nt) {
                if (k == 0 && j < skip) continue;  // the record before the map's first
                const PairRec r = buf[j];
                const unsigned int pc = (unsigned int)(r.cre - base), pd = (unsigned int)(r.des - base);
                if (pc >= (unsigned int)len && pd >= (unsigned int)len) continue;
                double gb, gd;
                pair_gradient(A, r, coef, creg, gb, gd);
                if (pc < (unsigned int)len) atomicAdd(tile + pc, (float)gb);
                if (pd < (unsigned int)len) atomicAdd(tile + pd, (float)gd);
            }
            if (k + 1 == nch) fence_async_proxy();  // last chunk: the tile is complete, the bulk store below reads it
            __syncthreads();  // everybody is done with buffer b
            if (threadIdx.x == 0 && k + 2 < nch) issue(k + 2);
            TL_GDBG(1);
        }
        if (nch == 0) { fence_async_proxy(); __syncthreads(); }
        if (((reinterpret_cast<uintptr_t>(g + base) & 15) == 0) && (len & 3) == 0) {
            if (threadIdx.x == 0) {
                for (int off = 0; off < len * 4; off += 32768) bulk_store(reinterpret_cast<char*>(g + base) + off, tile_s + (uint32_t)off, (uint32_t)min(32768, len * 4 - off));
                bulk_commit();
                bulk_wait_read();  // the tile has been read: it may be zeroed again
            }
        } else {
            for (int i = threadIdx.x; i < len; i += nt) g[base + i] = tile[i];
        }
        __syncthreads();
        TL_GDBG(2);
    }
#undef TL_GDBG
    return phase;
}

// One CTA per map (maps of images the persistence launch has already served are skipped).
__global__ void __launch_bounds__(512) grad_kernel(GradArgs A) {
    const double gl = A.grad_loss ? (double)__ldg(A.grad_loss) : 1.0;
    for (int map = blockIdx.x; map < A.M; map += gridDim.x) {
        if (A.gfused && A.gfused[map / A.C]) continue;  // block-uniform
        grad_one_map(A, map, A.coef[map / A.C] * gl, gl);
    }
}

// grad[i] *= *grad_loss, skipped altogether when the upstream gradient is exactly 1 (the usual `loss.backward()`)
__global__ void __launch_bounds__(256) scale_kernel(const float* __restrict__ grad_loss, float* __restrict__ g, long long n) {
    const float gl = __ldg(grad_loss);
    if (gl == 1.0f) return;
    const long long n4 = ((reinterpret_cast<uintptr_t>(g) & 15) == 0) ? n >> 2 : 0;
    float4* g4 = reinterpret_cast<float4*>(g);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = g4[i];
        v.x *= gl; v.y *= gl; v.z *= gl; v.w *= gl;
        g4[i] = v;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) g[i] *= gl;
}

}  // namespace tl
