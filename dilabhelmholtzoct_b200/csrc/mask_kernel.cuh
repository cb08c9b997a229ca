// Ground-truth masks that crossed PCIe bit-packed (numpy.packbits order: the first pixel is bit 7 of the first
// byte) are widened to the fp32 {0, 1} maps the persistence kernel reads.  The reference builds these masks on
// the CPU as {0.0, 1.0} arrays (/root/reference/octsam/models/training_utils.py:398, :413, :432).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tl {

// one thread per 32 pixels: a 4-byte read, eight 16-byte writes (consecutive threads write consecutive 128-byte lines)
__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint8_t* __restrict__ bits, float* __restrict__ out, long long n_bytes) {
    const long long n_words = n_bytes >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(bits) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (aligned) {
        for (long long w = t0; w < n_words; w += stride) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(bits) + w);  // little endian: byte k = bits [8k, 8k+8)
            float4* o = reinterpret_cast<float4*>(out + w * 32);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t b = (v >> (8 * k)) & 0xFFu;
                __stcs(o + 2 * k, make_float4((b >> 7) & 1u ? 1.f : 0.f, (b >> 6) & 1u ? 1.f : 0.f, (b >> 5) & 1u ? 1.f : 0.f, (b >> 4) & 1u ? 1.f : 0.f));
                __stcs(o + 2 * k + 1, make_float4((b >> 3) & 1u ? 1.f : 0.f, (b >> 2) & 1u ? 1.f : 0.f, (b >> 1) & 1u ? 1.f : 0.f, b & 1u ? 1.f : 0.f));
            }
        }
    }
    for (long long i = (aligned ? n_words * 4 : 0) + t0; i < n_bytes; i += stride) {  // tail bytes (or everything when unaligned)
        const uint32_t b = bits[i];
#pragma unroll
        for (int k = 0; k < 8; ++k) out[i * 8 + k] = (b >> (7 - k)) & 1u ? 1.f : 0.f;
    }
}

}  // namespace tl
