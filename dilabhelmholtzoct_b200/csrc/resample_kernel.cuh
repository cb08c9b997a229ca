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

// ---------------------------------------------------------------------------------------------
// F3: SAM post-processing of the mask decoder output at the reference call site
//
//   masks = F.interpolate(outputs.pred_masks.squeeze(2), (1024, 1024), mode="bilinear", align_corners=False)
//   masks = masks[..., :reshaped_h, :reshaped_w]
//   masks = F.interpolate(masks, (orig_h, orig_w), mode="bilinear", align_corners=False)
//       /root/reference/octsam/models/training_utils.py:57-59
//
// as ONE gather: an output pixel blends 4 pixels of the (never materialised) T x T intermediate, each of
// which blends 4 source pixels -- 16 taps on the small 256 x 256 map instead of writing and re-reading a
// [B, N, 1024, 1024] tensor.  Both stages follow ATen's upsample_bilinear2d (align_corners=False) in
// fp32, and the intermediate values are rounded to fp32 exactly where the two-step form rounds them:
//   scale = in / out;  src = max(0, scale * (dst + 0.5) - 0.5);  i0 = (int)src;  i1 = i0 + (i0 < in - 1)
//   l1 = src - i0;  l0 = 1 - l1
namespace tl {

struct PostArgs {
    const float* in;    // [n_maps][Hs][Ws]
    float* out;         // forward: [n_maps][oh][ow]
    const float* gout;  // backward: [n_maps][oh][ow]
    float* gin;         // backward: [n_maps][Hs][Ws]
    int n_maps, Hs, Ws, T, rh, rw, oh, ow;
    float s1y, s1x;     // Hs / T, Ws / T
    float s2y, s2x;     // rh / oh, rw / ow
};

__device__ __forceinline__ void half_pixel(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.0f - l1;
}

// One thread per output COLUMN of a kFwdCols x kFwdRows tile: the column's taps (second-stage columns X0, X1
// and their first-stage source columns and weights) are computed once and kept in registers while the thread
// walks down the tile's rows; the rows' taps are computed once per tile by kFwdRows threads and broadcast
// from shared memory.  Consecutive threads write consecutive columns (coalesced 128-byte stores).
constexpr int kFwdCols = 128, kFwdRows = 16;

struct RowTaps { int a[2][2]; float la[2][2]; float ly[2]; };  // source rows / weights of Y0, Y1; second-stage weights

__global__ void __launch_bounds__(kFwdCols) postprocess_fwd_kernel(PostArgs a, int tiles_y, int tiles_x) {
    __shared__ RowTaps rows[kFwdRows];
    const long long n_tiles = (long long)a.n_maps * tiles_y * tiles_x;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long t2 = tile / tiles_x;
        const int ty = (int)(t2 % tiles_y);
        const long long m = t2 / tiles_y;
        const int oy0 = ty * kFwdRows, ox = tx * kFwdCols + threadIdx.x;
        const int nrow = min(kFwdRows, a.oh - oy0);
        __syncthreads();  // the previous tile's readers are done
        if (threadIdx.x < nrow) {
            RowTaps r;
            int Y[2];
            half_pixel(a.s2y, oy0 + threadIdx.x, a.rh, Y[0], Y[1], r.ly[0], r.ly[1]);
#pragma unroll
            for (int k = 0; k < 2; ++k) half_pixel(a.s1y, Y[k], a.Hs, r.a[k][0], r.a[k][1], r.la[k][0], r.la[k][1]);
            rows[threadIdx.x] = r;
        }
        __syncthreads();
        if (ox >= a.ow) continue;
        int X[2], bc[2][2];
        float lx[2], lb[2][2];
        half_pixel(a.s2x, ox, a.rw, X[0], X[1], lx[0], lx[1]);
#pragma unroll
        for (int k = 0; k < 2; ++k) half_pixel(a.s1x, X[k], a.Ws, bc[k][0], bc[k][1], lb[k][0], lb[k][1]);
        const float* p = a.in + m * (long long)a.Hs * a.Ws;
        float* out = a.out + (m * a.oh + oy0) * (long long)a.ow + ox;
        for (int i = 0; i < nrow; ++i) {
            const RowTaps r = rows[i];
            float I[2][2];
#pragma unroll
            for (int y = 0; y < 2; ++y) {
                const float* r0 = p + r.a[y][0] * a.Ws;
                const float* r1 = p + r.a[y][1] * a.Ws;
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    const float v00 = __ldg(r0 + bc[x][0]), v01 = __ldg(r0 + bc[x][1]);
                    const float v10 = __ldg(r1 + bc[x][0]), v11 = __ldg(r1 + bc[x][1]);
                    I[y][x] = r.la[y][0] * (lb[x][0] * v00 + lb[x][1] * v01) + r.la[y][1] * (lb[x][0] * v10 + lb[x][1] * v11);
                }
            }
            out[(long long)i * a.ow] = r.ly[0] * (lx[0] * I[0][0] + lx[1] * I[0][1]) + r.ly[1] * (lx[0] * I[1][0] + lx[1] * I[1][1]);
        }
    }
}

// Backward: one CTA per (map, tile of kPostTileY x kPostTileX output pixels).  The tile's source footprint
// is a small window of the 256 x 256 map, accumulated in shared memory (shared-memory atomics) and
// flushed with one global atomicAdd per touched source pixel: ~50x fewer global atomics than scattering
// the 16 taps of every output pixel.
constexpr int kPostTileY = 16, kPostTileX = 64, kPostWin = 40 * 48;  // window capacity (rows x cols) in floats

__global__ void __launch_bounds__(256) postprocess_bwd_kernel(PostArgs a, int tiles_y, int tiles_x) {
    __shared__ float win[kPostWin];
    __shared__ int s_box[4];  // r_lo, c_lo, n_rows, n_cols of the source window
    const long long n_tiles = (long long)a.n_maps * tiles_y * tiles_x;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long t2 = tile / tiles_x;
        const int ty = (int)(t2 % tiles_y);
        const long long m = t2 / tiles_y;
        const int oy0 = ty * kPostTileY, ox0 = tx * kPostTileX;
        const int oy1 = min(a.oh, oy0 + kPostTileY) - 1, ox1 = min(a.ow, ox0 + kPostTileX) - 1;
        if (threadIdx.x == 0) {  // source window of the tile: both stages are monotone in the index
            int i0, i1; float l0, l1;
            int Ylo, Yhi, Xlo, Xhi, r_lo, r_hi, c_lo, c_hi;
            half_pixel(a.s2y, oy0, a.rh, Ylo, i1, l0, l1); half_pixel(a.s2y, oy1, a.rh, i0, Yhi, l0, l1);
            half_pixel(a.s2x, ox0, a.rw, Xlo, i1, l0, l1); half_pixel(a.s2x, ox1, a.rw, i0, Xhi, l0, l1);
            half_pixel(a.s1y, Ylo, a.Hs, r_lo, i1, l0, l1); half_pixel(a.s1y, Yhi, a.Hs, i0, r_hi, l0, l1);
            half_pixel(a.s1x, Xlo, a.Ws, c_lo, i1, l0, l1); half_pixel(a.s1x, Xhi, a.Ws, i0, c_hi, l0, l1);
            s_box[0] = r_lo; s_box[1] = c_lo; s_box[2] = r_hi - r_lo + 1; s_box[3] = c_hi - c_lo + 1;
        }
        __syncthreads();
        const int r_lo = s_box[0], c_lo = s_box[1], nr = s_box[2], nc = s_box[3];
        const bool fits = nr * nc <= kPostWin;
        float* gin = a.gin + m * (long long)a.Hs * a.Ws;
        if (fits) for (int i = threadIdx.x; i < nr * nc; i += blockDim.x) win[i] = 0.f;
        __syncthreads();
        const float* g = a.gout + m * (long long)a.oh * a.ow;
        const int tw = ox1 - ox0 + 1, th = oy1 - oy0 + 1;
        for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
            const int oy = oy0 + i / tw, ox = ox0 + i % tw;
            const float go = __ldg(g + (long long)oy * a.ow + ox);
            if (go == 0.f) continue;
            int Y[2], X[2];
            float ly[2], lx[2];
            half_pixel(a.s2y, oy, a.rh, Y[0], Y[1], ly[0], ly[1]);
            half_pixel(a.s2x, ox, a.rw, X[0], X[1], lx[0], lx[1]);
#pragma unroll
            for (int iy = 0; iy < 2; ++iy) {
                int ar[2]; float la[2];
                half_pixel(a.s1y, Y[iy], a.Hs, ar[0], ar[1], la[0], la[1]);
#pragma unroll
                for (int ix = 0; ix < 2; ++ix) {
                    int bc[2]; float lb[2];
                    half_pixel(a.s1x, X[ix], a.Ws, bc[0], bc[1], lb[0], lb[1]);
                    const float w = ly[iy] * lx[ix] * go;
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float v = w * la[p] * lb[q];
                            if (fits) atomicAdd(&win[(ar[p] - r_lo) * nc + (bc[q] - c_lo)], v);
                            else atomicAdd(gin + ar[p] * a.Ws + bc[q], v);
                        }
                }
            }
        }
        __syncthreads();
        if (fits)
            for (int i = threadIdx.x; i < nr * nc; i += blockDim.x) {
                const float v = win[i];
                if (v != 0.f) atomicAdd(gin + (r_lo + i / nc) * a.Ws + c_lo + i % nc, v);
            }
        __syncthreads();
    }
}

// Backward as a GATHER (no atomics, no zero fill): a CTA owns a kGT_Y x kGT_X tile of SOURCE pixels.  Both
// stages are separable and monotone in the index, so the output pixels that read source column s form one
// interval [lo(s), hi(s)] (found by bisection on the forward's own tap formula), likewise for rows; the tile's
// window of grad_out is staged in shared memory once, reduced along x with the composite column weights into
// T[row][source column], then along y.  Every source pixel is written exactly once.
// (Down-sampling to the original size, the reference's case: ~6 taps per axis and source pixel.  When the
// output is much larger than the source the taps per source pixel grow and the host keeps the scatter kernel.)
constexpr int kGT_Y = 8, kGT_X = 32, kGK = 16, kGWR = 40, kGWC = 104, kGRows = 256;
constexpr int kGTileRows = 2 * kGT_Y;  // source rows per tile: every thread produces two of them

// weight of source index s in output index o along one axis: two bilinear stages, ATen's align_corners=False
__device__ __forceinline__ float post_axis_weight(float s2, int mid_in, float s1, int src_in, int o, int s) {
    int Y0, Y1, a0, a1; float l0, l1, p0, p1;
    half_pixel(s2, o, mid_in, Y0, Y1, l0, l1);
    half_pixel(s1, Y0, src_in, a0, a1, p0, p1);
    float w = l0 * ((a0 == s ? p0 : 0.f) + (a1 == s ? p1 : 0.f));
    half_pixel(s1, Y1, src_in, a0, a1, p0, p1);
    return w + l1 * ((a0 == s ? p0 : 0.f) + (a1 == s ? p1 : 0.f));
}
// first / last source index an output index reads (monotone non-decreasing in o)
__device__ __forceinline__ int post_src_lo(float s2, int mid_in, float s1, int src_in, int o) {
    int Y0, Y1, a0, a1; float l0, l1;
    half_pixel(s2, o, mid_in, Y0, Y1, l0, l1);
    half_pixel(s1, Y0, src_in, a0, a1, l0, l1);
    return a0;
}
__device__ __forceinline__ int post_src_hi(float s2, int mid_in, float s1, int src_in, int o) {
    int Y0, Y1, a0, a1; float l0, l1;
    half_pixel(s2, o, mid_in, Y0, Y1, l0, l1);
    half_pixel(s1, Y1, src_in, a0, a1, l0, l1);
    return a1;
}
// [lo, hi] of the output indices that may read source index s (empty: hi < lo)
__device__ __forceinline__ void post_out_range(float s2, int mid_in, float s1, int src_in, int n_out, int s, int& lo, int& hi) {
    int a = 0, b = n_out;  // smallest o with src_hi(o) >= s
    while (a < b) { const int m = (a + b) >> 1; if (post_src_hi(s2, mid_in, s1, src_in, m) >= s) b = m; else a = m + 1; }
    lo = a;
    a = -1; b = n_out - 1;  // largest o with src_lo(o) <= s
    while (a < b) { const int m = (a + b + 1) >> 1; if (post_src_lo(s2, mid_in, s1, src_in, m) <= s) a = m; else b = m - 1; }
    hi = a;
}

// A job is a STRIP: one map, kGT_X source columns, up to kGRows source rows.  The strip's column ranges and
// weights and ALL its row ranges and weights are computed once (one thread per row / column), then the CTA walks
// down the strip in tiles of kGT_Y rows: stage the tile's window of grad_out, reduce along x, reduce along y.
__global__ void __launch_bounds__(kGT_Y * kGT_X) postprocess_bwd_gather_kernel(PostArgs a, int chunks_y, int tiles_x) {
    __shared__ float gw[kGWR][kGWC + 1];
    __shared__ float Tt[kGWR][kGT_X];
    __shared__ float wx[kGT_X][kGK], wy[kGRows][kGK];
    __shared__ int xlo[kGT_X], xn[kGT_X], ylo[kGRows], yn[kGRows];
    __shared__ int s_x[3];  // ox0, cols, columns fit
    const int tid = threadIdx.x, nt = kGT_Y * kGT_X;
    const int sy = tid / kGT_X, sx = tid - sy * kGT_X;
    // The ranges and weights depend on the geometry only, not on the map: a CTA keeps ONE strip position
    // (column tile, row chunk) and walks over the maps, so the set-up below runs once per CTA.
    const int n_classes = chunks_y * tiles_x;
    const int cls = blockIdx.x % n_classes, m_first = blockIdx.x / n_classes, m_step = max(1, (int)gridDim.x / n_classes);
    if (blockIdx.x >= (unsigned)(m_step * n_classes)) return;  // (grid is a multiple of the classes: never taken)
    const int tx = cls % tiles_x, cy = cls / tiles_x;
    const int sx0 = tx * kGT_X, row0 = cy * kGRows, n_rows = min(kGRows, a.Hs - row0);
    {
        if (tid < kGT_X) {
            int lo = 0, hi = -1;
            if (sx0 + tid < a.Ws) post_out_range(a.s2x, a.rw, a.s1x, a.Ws, a.ow, sx0 + tid, lo, hi);
            xlo[tid] = lo; xn[tid] = hi >= lo ? hi - lo + 1 : 0;
        }
        if (tid < n_rows) {
            int lo = 0, hi = -1;
            post_out_range(a.s2y, a.rh, a.s1y, a.Hs, a.oh, row0 + tid, lo, hi);
            ylo[tid] = lo; yn[tid] = hi >= lo ? hi - lo + 1 : 0;
        }
        __syncthreads();
        if (tid == 0) {
            int x0 = 1 << 30, x1 = -1, kmax = 0;
            for (int i = 0; i < kGT_X; ++i) if (xn[i]) { x0 = min(x0, xlo[i]); x1 = max(x1, xlo[i] + xn[i] - 1); kmax = max(kmax, xn[i]); }
            const int cols = x1 >= x0 ? x1 - x0 + 1 : 0;
            s_x[0] = cols ? x0 : 0; s_x[1] = cols; s_x[2] = cols <= kGWC && kmax <= kGK;
        }
        for (int i = tid; i < kGT_X * kGK; i += nt) {
            const int c = i / kGK, j = i - c * kGK;
            wx[c][j] = j < xn[c] ? post_axis_weight(a.s2x, a.rw, a.s1x, a.Ws, xlo[c] + j, sx0 + c) : 0.f;
        }
        for (int i = tid; i < n_rows * kGK; i += nt) {
            const int r = i / kGK, j = i - r * kGK;
            wy[r][j] = j < yn[r] ? post_axis_weight(a.s2y, a.rh, a.s1y, a.Hs, ylo[r] + j, row0 + r) : 0.f;
        }
        __syncthreads();
    }
    for (long long m = m_first; m < a.n_maps; m += m_step) {
        const float* g = a.gout + m * (long long)a.oh * a.ow;
        float* gin = a.gin + m * (long long)a.Hs * a.Ws;
        const int ox0 = s_x[0], cols = s_x[1];
        for (int r0 = 0; r0 < n_rows; r0 += kGTileRows) {
            // the tile's rows of grad_out: ranges are monotone in the source row
            const int r_last = min(n_rows, r0 + kGTileRows) - 1;
            int oy0 = 1 << 30, oy1 = -1, kmax = 0;
            for (int i = r0; i <= r_last; ++i) if (yn[i]) { oy0 = min(oy0, ylo[i]); oy1 = max(oy1, ylo[i] + yn[i] - 1); kmax = max(kmax, yn[i]); }
            const int rows = oy1 >= oy0 ? oy1 - oy0 + 1 : 0;
            const bool fits = s_x[2] && rows <= kGWR && kmax <= kGK;  // block-uniform
            float acc[2] = {0.f, 0.f};
            if (fits) {
                for (int r = tid >> 5; r < rows; r += nt >> 5)  // a warp per window row: coalesced, no divisions
                    for (int c = tid & 31; c < cols; c += 32) gw[r][c] = __ldg(g + (long long)(oy0 + r) * a.ow + ox0 + c);
                __syncthreads();
                for (int i = tid; i < rows * kGT_X; i += nt) {  // along x: T[row][source column]
                    const int r = i / kGT_X, c = i - r * kGT_X;
                    float t = 0.f;
                    const int base = xlo[c] - ox0, n = xn[c];
                    for (int j = 0; j < n; ++j) t += wx[c][j] * gw[r][base + j];
                    Tt[r][c] = t;
                }
                __syncthreads();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int my = r0 + sy + h * kGT_Y;
                    if (my <= r_last) {
                        const int base = ylo[my] - oy0, n = yn[my];
                        for (int j = 0; j < n; ++j) acc[h] += wy[my][j] * Tt[base + j][sx];
                    }
                }
                __syncthreads();  // Tt / gw are rewritten by the next tile
            } else {  // window or tap count beyond the staging buffers: straight from global
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int my = r0 + sy + h * kGT_Y;
                    if (my > r_last || sx0 + sx >= a.Ws) continue;
                    for (int jy = 0; jy < yn[my]; ++jy) {
                        const int oy = ylo[my] + jy;
                        const float wv = post_axis_weight(a.s2y, a.rh, a.s1y, a.Hs, oy, row0 + my);
                        if (wv == 0.f) continue;
                        float t = 0.f;
                        for (int jx = 0; jx < xn[sx]; ++jx) {
                            const int ox = xlo[sx] + jx;
                            t += post_axis_weight(a.s2x, a.rw, a.s1x, a.Ws, ox, sx0 + sx) * __ldg(g + (long long)oy * a.ow + ox);
                        }
                        acc[h] += wv * t;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int my = r0 + sy + h * kGT_Y;
                if (my <= r_last && sx0 + sx < a.Ws) gin[(long long)(row0 + my) * a.Ws + sx0 + sx] = acc[h];
            }
        }
    }
}

}  // namespace tl
