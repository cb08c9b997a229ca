// F1: the step in front of the persistence path at the reference call site
//
//   topo_loss(torch.sigmoid(masks.float()), gt_masks.float(), 0.1, feat_d=1, interp=50)
//       /root/reference/octsam/models/training_utils.py:64
//   -> F.interpolate(x, size=(interp, interp), mode='bilinear', align_corners=True) on both inputs
//       /root/reference/octsam/models/topological_loss.py:33-46
//
// fused into one gather: an S x S output pixel reads its <= 4 source pixels and applies the sigmoid to
// those only (the reference takes the sigmoid of the whole H x W map and then keeps ~4 S^2 of its
// values: at 496x512 -> 50x50 that is 4 % of the map).  Arithmetic follows ATen's
// upsample_bilinear2d (align_corners=True) in fp32:
//   scale = (in - 1) / (out - 1)   (0 when out == 1)
//   src = scale * dst;  i0 = (int)src;  i1 = i0 + (i0 < in - 1);  l1 = src - i0;  l0 = 1 - l1
//   out = l0y * (l0x * v00 + l1x * v01) + l1y * (l0x * v10 + l1x * v11)
// and torch.sigmoid's 1 / (1 + exp(-x)).
//
// Backward: grad_in is dense (autograd hands it to the mask decoder), so it is zero-filled with
// 128-bit stores and the S x S upstream gradients are scattered with atomics (4 per output pixel),
// times sigmoid'(x) = s (1 - s) recomputed from the logit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tl {

struct ResampleArgs {
    const float* in;     // [n_maps][H][W]
    float* out;          // forward: [n_maps][S][S]
    const float* gout;   // backward: [n_maps][S][S]
    float* gin;          // backward: [n_maps][H][W]
    int n_maps, H, W, S;
    int apply_sigmoid;
    float sy, sx;        // (H - 1) / (S - 1), (W - 1) / (S - 1)
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
    const float src = scale * (float)dst;
    i0 = (int)src;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.0f - l1;
}

// one thread per output pixel; consecutive threads -> consecutive output columns
__global__ void __launch_bounds__(256) resample_fwd_kernel(ResampleArgs a) {
    const long long total = (long long)a.n_maps * a.S * a.S;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(o % a.S);
        const long long t = o / a.S;
        const int oy = (int)(t % a.S);
        const long long m = t / a.S;
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        src_index(a.sy, oy, a.H, y0, y1, ly0, ly1);
        src_index(a.sx, ox, a.W, x0, x1, lx0, lx1);
        const float* p = a.in + m * (long long)a.H * a.W;
        float v00 = __ldg(p + (long long)y0 * a.W + x0), v01 = __ldg(p + (long long)y0 * a.W + x1);
        float v10 = __ldg(p + (long long)y1 * a.W + x0), v11 = __ldg(p + (long long)y1 * a.W + x1);
        if (a.apply_sigmoid) { v00 = sigmoid_f(v00); v01 = sigmoid_f(v01); v10 = sigmoid_f(v10); v11 = sigmoid_f(v11); }
        a.out[o] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    }
}

// dense zero-fill of grad_in with 128-bit stores (n4 float4 then the scalar tail)
__global__ void __launch_bounds__(256) zero_fill_kernel(float* p, long long n) {
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) p4[i] = z;
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = 0.f;
}

__global__ void __launch_bounds__(256) resample_bwd_kernel(ResampleArgs a) {
    const long long total = (long long)a.n_maps * a.S * a.S;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const float g = __ldg(a.gout + o);
        if (g == 0.f) continue;  // the topological gradient is sparse (critical pixels only)
        const int ox = (int)(o % a.S);
        const long long t = o / a.S;
        const int oy = (int)(t % a.S);
        const long long m = t / a.S;
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        src_index(a.sy, oy, a.H, y0, y1, ly0, ly1);
        src_index(a.sx, ox, a.W, x0, x1, lx0, lx1);
        const long long base = m * (long long)a.H * a.W;
        const long long i00 = base + (long long)y0 * a.W + x0, i01 = base + (long long)y0 * a.W + x1;
        const long long i10 = base + (long long)y1 * a.W + x0, i11 = base + (long long)y1 * a.W + x1;
        float w00 = ly0 * lx0 * g, w01 = ly0 * lx1 * g, w10 = ly1 * lx0 * g, w11 = ly1 * lx1 * g;
        if (a.apply_sigmoid) {
            const float s00 = sigmoid_f(__ldg(a.in + i00)), s01 = sigmoid_f(__ldg(a.in + i01));
            const float s10 = sigmoid_f(__ldg(a.in + i10)), s11 = sigmoid_f(__ldg(a.in + i11));
            w00 *= s00 * (1.0f - s00); w01 *= s01 * (1.0f - s01); w10 *= s10 * (1.0f - s10); w11 *= s11 * (1.0f - s11);
        }
        atomicAdd(a.gin + i00, w00);
        atomicAdd(a.gin + i01, w01);
        atomicAdd(a.gin + i10, w10);
        atomicAdd(a.gin + i11, w11);
    }
}

}  // namespace tl
