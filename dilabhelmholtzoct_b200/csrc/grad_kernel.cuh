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
