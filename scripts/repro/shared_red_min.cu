// Stand-alone probe of the pattern that failed in the two-reduction form of the Boruvka rounds
// (DESIGN.md section 6): native 32-bit shared-memory reductions (red.shared.min.u32) on the two halves of a
// 64-bit entry, in two sweeps separated by membar.cta + bar.sync, then a consistency check:
//   sweep 1   hi[a] = min(hi[a], v)           for both ends of every edge
//   sweep 2   if (hi[a] == v) lo[a] = min(lo[a], id)
//   check     hi set  <=>  lo set
// Prints the number of inconsistent entries over all rounds (expected 0).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o shared_red_min shared_red_min.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_min32(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024, 1) probe(const uint4* edges, int n, int K, int rounds, int fence, unsigned long long* bad) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t T = (uint32_t)__cvta_generic_to_shared(smem);
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint4* my = edges + (size_t)blockIdx.x * n;
    unsigned long long local_bad = 0;
    for (int r = 0; r < rounds; ++r) {
        for (int c = tid; c <= K; c += nt) asm volatile("st.volatile.shared.u64 [%0], %1;" :: "r"(T + c * 8u), "l"(~0ull) : "memory");
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const uint4 e = my[i];  // x: value, y: id, z: a, w: b
            red_min32(T + e.z * 8u + 4u, e.x + r);
            red_min32(T + e.w * 8u + 4u, e.x + r);
        }
        if (fence) __threadfence_block();
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const uint4 e = my[i];
            if (lds32(T + e.z * 8u + 4u) == e.x + r) red_min32(T + e.z * 8u, e.y);
            if (lds32(T + e.w * 8u + 4u) == e.x + r) red_min32(T + e.w * 8u, e.y);
        }
        if (fence) __threadfence_block();
        __syncthreads();
        for (int c = tid; c <= K; c += nt) {
            const uint64_t v = lds64(T + c * 8u);
            const bool hi_set = (uint32_t)(v >> 32) != 0xFFFFFFFFu, lo_set = (uint32_t)v != 0xFFFFFFFFu;
            if (hi_set != lo_set) ++local_bad;
        }
        __syncthreads();
    }
    if (local_bad) atomicAdd(bad, local_bad);
}

int main(int argc, char** argv) {
    const int K = 10000, n = 61000, rounds = argc > 1 ? atoi(argv[1]) : 50, grid = 148;
    std::vector<uint4> h((size_t)grid * n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
    for (size_t i = 0; i < h.size(); ++i) {
        // neighbouring records share an end, like the compaction's list
        const uint32_t a = (uint32_t)((i / 3) % K) + 1, b = (a + 1 + rnd() % 7) % K + 1;
        h[i] = make_uint4(0x40000000u + (rnd() >> 4), (uint32_t)(i % n), a, b == a ? (a % K) + 1 : b);
    }
    uint4* d; unsigned long long* bad;
    cudaMalloc(&d, h.size() * sizeof(uint4)); cudaMalloc(&bad, 8);
    cudaMemcpy(d, h.data(), h.size() * sizeof(uint4), cudaMemcpyHostToDevice);
    const int smem = (K + 1) * 8;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int fence = 0; fence < 2; ++fence) {
        cudaMemset(bad, 0, 8);
        probe<<<grid, 1024, smem>>>(d, n, K, rounds, fence, bad);
        unsigned long long hb = 0;
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
        printf("fence=%d rounds=%d ctas=%d: %s, inconsistent entries: %llu\n", fence, rounds, grid, cudaGetErrorString(e), hb);
    }
    return 0;
}
