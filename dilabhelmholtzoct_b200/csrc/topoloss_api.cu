// C ABI of libtopoloss.so (see include/topoloss.h for the contract and the reference
// interface each entry point replaces).  Host side: argument checks, workspace carving and
// kernel launches on the caller's stream.  No device allocation, no host synchronisation.
#include "../../include/topoloss.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "match_kernel.cuh"
#include "ph_kernel.cuh"
#include "ph_small.cuh"
#include "resample_kernel.cuh"
#include "seg_sort.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define TL_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) return fail(TL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// Optional per-kernel timing (bench.py's roofline): when enabled on this thread, tl_forward and
// tl_backward bracket each kernel with cudaEventRecord on the caller's stream.  Events are created
// lazily and reused; nothing is synchronised until tl_timing_read.
constexpr int kStages = 6;        // ph, sort, match, loss, grad-zero, grad-scatter
constexpr int kTimingRing = 128;  // calls remembered between two reads
struct Timing {
    bool on = false;
    int n_fwd = 0, n_bwd = 0;
    cudaEvent_t ev[kTimingRing][kStages + 2] = {};
    bool have[kTimingRing] = {};
};
Timing g_timing;  // process-wide: tl_backward runs on autograd's worker thread

cudaEvent_t timing_event(int call, int idx) {
    Timing& t = g_timing;
    if (!t.have[call]) {
        for (int i = 0; i < kStages + 2; ++i) cudaEventCreate(&t.ev[call][i]);
        t.have[call] = true;
    }
    return t.ev[call][idx];
}
#define TL_MARK(call, idx, st) do { if (g_timing.on && (call) < kTimingRing) cudaEventRecord(timing_event((call), (idx)), (st)); } while (0)

constexpr int kPhSlots = 296;     // CTAs of the persistence kernel (2 per SM on a 148-SM B200)
constexpr int kSmallSlots = 160;   // >= SM count: one 1024-thread CTA per SM in ph_small_kernel
constexpr int kSortSlots = 296;
constexpr int kMatchSlots = 296;

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

int max_pairs(int H, int W, int dim) {
    // one pair per level-0 root at most; roots are strict local extrema or mutual picks (<= nodes/2 + 1)
    const long long nn = dim == 1 ? (long long)H * W + 1 : (long long)(H + 1) * (W + 1);
    return (int)(nn / 2 + 2);
}

struct Layout {
    int M, cap, n_nodes;
    size_t counter, counts[2], cost, tpers, coef, pairs[2], skeys[2], match1;
    size_t T, t_stride;
    int small;  // 1: shared-memory persistence kernel (<= 65535 nodes)
    size_t rootpix, zval, T2g, k_stride, elist, e_stride;
    size_t key_tmp, idx_a, idx_b, rec_tmp;
    size_t v, minv, u, way, pcol, used, stride_c, stride_r;
    size_t total;
};

// Carve the workspace.  n_sets = 2 for the loss (pred + truth), 1 for tl_persistence_pairs.
Layout make_layout(int M, int H, int W, int dim, int B) {
    Layout L;
    memset(&L, 0, sizeof(L));
    L.M = M;
    L.cap = max_pairs(H, W, dim);
    L.n_nodes = dim == 1 ? H * W + 1 : (H + 1) * (W + 1);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return at; };
    L.counter = take(256);
    for (int s = 0; s < 2; ++s) L.counts[s] = take(sizeof(int32_t) * (size_t)M);
    L.cost = take(sizeof(double) * (size_t)M);
    L.tpers = take(sizeof(double) * (size_t)M);
    L.coef = take(sizeof(double) * (size_t)(B > 0 ? B : 1));
    for (int s = 0; s < 2; ++s) L.pairs[s] = take(sizeof(tl::PairRec) * (size_t)M * L.cap);
    for (int s = 0; s < 2; ++s) L.skeys[s] = take(sizeof(uint64_t) * (size_t)M * L.cap);
    L.match1 = take(sizeof(int32_t) * (size_t)M * L.cap);
    // shared-memory kernel: processes the map in bands of whole node rows (<= 65535 nodes each)
    L.small = (W + 1) <= tl::kSmallMaxRow && (long long)L.n_nodes < (1ll << 22);
    if (L.small) {
        L.k_stride = align_up((size_t)L.cap + 2, 64);
        L.rootpix = take(sizeof(uint32_t) * L.k_stride * kSmallSlots);
        L.zval = take(sizeof(uint32_t) * L.k_stride * kSmallSlots);
        L.T2g = take(sizeof(tl::TEntry) * L.k_stride * kSmallSlots);
        L.e_stride = align_up((size_t)H * (W + 1) + (size_t)(H + 1) * W, 64);
        L.elist = take(sizeof(tl::CrossEdge) * L.e_stride * kSmallSlots);
    }
    {
        const char* fg = getenv("TL_FORCE_GLOBAL");
        if (!L.small || (fg && fg[0] == '1')) {
            L.t_stride = align_up(sizeof(uint64_t) * (size_t)L.n_nodes) / sizeof(uint64_t);
            L.T = take(sizeof(uint64_t) * L.t_stride * kPhSlots);
        }
    }
    L.key_tmp = take(sizeof(uint64_t) * (size_t)L.cap * kSortSlots);
    L.idx_a = take(sizeof(uint32_t) * (size_t)L.cap * kSortSlots);
    L.idx_b = take(sizeof(uint32_t) * (size_t)L.cap * kSortSlots);
    L.rec_tmp = take(sizeof(tl::PairRec) * (size_t)L.cap * kSortSlots);
    L.stride_c = align_up((size_t)2 * L.cap + 2, 32);
    L.stride_r = align_up((size_t)L.cap + 2, 32);
    L.v = take(sizeof(double) * L.stride_c * kMatchSlots);
    L.minv = take(sizeof(double) * L.stride_c * kMatchSlots);
    L.u = take(sizeof(double) * L.stride_r * kMatchSlots);
    L.way = take(sizeof(int32_t) * L.stride_c * kMatchSlots);
    L.pcol = take(sizeof(int32_t) * L.stride_c * kMatchSlots);
    L.used = take(L.stride_c * kMatchSlots);
    L.total = o;
    return L;
}

int check_shape(int B, int C, int H, int W, int feat_d) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(TL_ERR_ARG, "bad shape [%d,%d,%d,%d]", B, C, H, W);
    if (feat_d != 0 && feat_d != 1) return fail(TL_ERR_ARG, "feat_d must be 0 or 1 on 2-D maps, got %d", feat_d);
    if (H != W) return fail(TL_ERR_ARG, "non-square maps are not supported (H=%d, W=%d)", H, W);
    if ((long long)(2 * H + 1) * (2 * W + 1) >= 0x7FFFFFF0LL) return fail(TL_ERR_ARG, "map too large");
    if ((long long)B * C > (1 << 24)) return fail(TL_ERR_ARG, "too many maps");
    return TL_OK;
}

template <typename T>
T* at(void* ws, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(ws) + off); }

int launch_ph(const float* m0, const float* m1, int n_sets, const Layout& L, int H, int W, int dim,
              void* ws, cudaStream_t st, bool want_sort_keys) {
    tl::PhArgs a;
    a.maps[0] = m0; a.maps[1] = m1;
    for (int s = 0; s < 2; ++s) {
        a.pairs[s] = at<tl::PairRec>(ws, L.pairs[s]);
        a.skeys[s] = want_sort_keys ? at<uint64_t>(ws, L.skeys[s]) : nullptr;
        a.counts[s] = at<int32_t>(ws, L.counts[s]);
    }
    a.n_sets = n_sets; a.n_maps = L.M; a.H = H; a.W = W; a.cap = L.cap;
    a.magic_W = tl::FastDiv::magic_of((uint32_t)W); a.magic_GW = tl::FastDiv::magic_of((uint32_t)(2 * W + 1));
    a.magic_VW = tl::FastDiv::magic_of((uint32_t)(W + 1));
    a.T = at<uint64_t>(ws, L.T); a.t_stride = L.t_stride;
    a.job_counter = at<unsigned int>(ws, L.counter);
    TL_CUDA(cudaMemsetAsync(a.job_counter, 0, 256, st));
    long long jobs = (long long)n_sets * L.M;
    int grid = (int)(jobs < kPhSlots ? jobs : kPhSlots);
    // TL_FORCE_GLOBAL=1 (tests) routes every shape through the global-memory kernel, which otherwise only
    // serves maps wider than kSmallMaxRow
    const char* fg = getenv("TL_FORCE_GLOBAL");
    const bool use_small = L.small && !(fg && fg[0] == '1');
    if (use_small) {
        static int n_sm = 0;
        if (n_sm == 0) {
            int dev = 0, v = 0;
            TL_CUDA(cudaGetDevice(&dev));
            TL_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
            n_sm = v > 0 ? v : 148;
        }
        if (grid > n_sm) grid = n_sm;  // one 1024-thread CTA per SM, work handed out dynamically
        if (grid > kSmallSlots) grid = kSmallSlots;
        tl::PhSmallArgs sa;
        sa.base = a;
        sa.rootpix = at<uint32_t>(ws, L.rootpix); sa.zval = at<uint32_t>(ws, L.zval);
        sa.T2g = at<tl::TEntry>(ws, L.T2g);
        sa.k_stride = L.k_stride;
        sa.elist = at<tl::CrossEdge>(ws, L.elist); sa.e_stride = L.e_stride;
        const char* pe = getenv("TL_PROFILE");
        sa.prof = (pe && pe[0] == '1') ? at<unsigned long long>(ws, L.counter) + 8 : nullptr;
        const char* nb = getenv("TL_NO_BINARY");
        sa.binary_path = !(nb && nb[0] == '1');
        if (dim == 1) {
            TL_CUDA(cudaFuncSetAttribute(tl::ph_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tl::kSmallSmemBytes));
            tl::ph_small_kernel<1><<<grid, tl::kPhThreads, tl::kSmallSmemBytes, st>>>(sa);
        } else {
            TL_CUDA(cudaFuncSetAttribute(tl::ph_small_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tl::kSmallSmemBytes));
            tl::ph_small_kernel<0><<<grid, tl::kPhThreads, tl::kSmallSmemBytes, st>>>(sa);
        }
    } else if (dim == 1) tl::ph_kernel<1><<<grid, tl::kPhThreads, 0, st>>>(a);
    else tl::ph_kernel<0><<<grid, tl::kPhThreads, 0, st>>>(a);
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

int launch_sort(int n_sets, const Layout& L, void* ws, cudaStream_t st) {
    tl::SortArgs a;
    for (int s = 0; s < 2; ++s) {
        a.pairs[s] = at<tl::PairRec>(ws, L.pairs[s]);
        a.skeys[s] = at<uint64_t>(ws, L.skeys[s]);
        a.counts[s] = at<int32_t>(ws, L.counts[s]);
    }
    a.n_sets = n_sets; a.n_maps = L.M; a.cap = L.cap;
    a.key_tmp = at<uint64_t>(ws, L.key_tmp);
    a.idx_a = at<uint32_t>(ws, L.idx_a);
    a.idx_b = at<uint32_t>(ws, L.idx_b);
    a.rec_tmp = at<tl::PairRec>(ws, L.rec_tmp);
    long long jobs = (long long)n_sets * L.M;
    int grid = (int)(jobs < kSortSlots ? jobs : kSortSlots);
    tl::seg_sort_kernel<<<grid, tl::kSortThreads, 0, st>>>(a);
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

void fill_match_scratch(tl::MatchArgs& a, void* ws, size_t v, size_t minv, size_t u, size_t way, size_t pcol,
                        size_t used, size_t stride_c, size_t stride_r) {
    a.v = at<double>(ws, v); a.minv = at<double>(ws, minv); a.u = at<double>(ws, u);
    a.way = at<int32_t>(ws, way); a.pcol = at<int32_t>(ws, pcol); a.used = at<uint8_t>(ws, used);
    a.stride_c = stride_c; a.stride_r = stride_r;
}

}  // namespace

extern "C" {

int tl_version(void) { return TL_ABI_VERSION; }

const char* tl_last_error(void) { return g_err; }

/* Debug aid (not part of the hot path): copies the 8 phase cycle counters that the persistence
 * kernel accumulates when the environment has TL_PROFILE=1.  Synchronises the device. */
int tl_debug_profile(const void* ws, unsigned long long* host_out8) {
    if (!ws || !host_out8) return fail(TL_ERR_ARG, "null pointer");
    TL_CUDA(cudaDeviceSynchronize());
    TL_CUDA(cudaMemcpy(host_out8, static_cast<const char*>(ws) + 64, 64, cudaMemcpyDeviceToHost));
    return TL_OK;
}

int tl_timing_enable(int on) {
    g_timing.on = on != 0;
    g_timing.n_fwd = g_timing.n_bwd = 0;
    return TL_OK;
}

int tl_timing_read(float* ms_sum6, int* n_calls) {
    if (!ms_sum6 || !n_calls) return fail(TL_ERR_ARG, "null pointer");
    Timing& t = g_timing;
    for (int i = 0; i < kStages; ++i) ms_sum6[i] = 0.f;
    const int nf = t.n_fwd < kTimingRing ? t.n_fwd : kTimingRing;
    const int nb = t.n_bwd < kTimingRing ? t.n_bwd : kTimingRing;
    for (int c = 0; c < nf; ++c) {
        TL_CUDA(cudaEventSynchronize(t.ev[c][4]));
        for (int i = 0; i < 4; ++i) { float ms = 0.f; TL_CUDA(cudaEventElapsedTime(&ms, t.ev[c][i], t.ev[c][i + 1])); ms_sum6[i] += ms; }
    }
    for (int c = 0; c < nb; ++c) {
        TL_CUDA(cudaEventSynchronize(t.ev[c][7]));
        for (int i = 4; i < 6; ++i) { float ms = 0.f; TL_CUDA(cudaEventElapsedTime(&ms, t.ev[c][i + 1], t.ev[c][i + 2])); ms_sum6[i] += ms; }
    }
    n_calls[0] = nf; n_calls[1] = nb;
    t.n_fwd = t.n_bwd = 0;
    return TL_OK;
}

#ifdef TL_STATS
int tl_debug_stats(unsigned long long* host_out8, int reset) {
    TL_CUDA(cudaDeviceSynchronize());
    TL_CUDA(cudaMemcpyFromSymbol(host_out8, tl::g_stats, 64));
    if (reset) { unsigned long long z[8] = {0}; TL_CUDA(cudaMemcpyToSymbol(tl::g_stats, z, 64)); }
    return TL_OK;
}
#endif

int tl_max_pairs(int H, int W, int dim) {
    if (H <= 0 || W <= 0 || (dim != 0 && dim != 1)) return fail(TL_ERR_ARG, "bad arguments");
    return max_pairs(H, W, dim);
}

int tl_workspace_bytes(int B, int C, int H, int W, int feat_d, size_t* bytes) {
    if (!bytes) return fail(TL_ERR_ARG, "bytes is null");
    int rc = check_shape(B, C, H, W, feat_d);
    if (rc != TL_OK) return rc;
    *bytes = make_layout(B * C, H, W, feat_d, B).total;
    return TL_OK;
}

int tl_forward(const float* pred, const float* truth, int B, int C, int H, int W, int feat_d, float q,
               float lamda, int loss_r, int B_global, void* ws, size_t ws_bytes, float* loss_out,
               void* stream) {
    int rc = check_shape(B, C, H, W, feat_d);
    if (rc != TL_OK) return rc;
    if (!pred || !truth || !ws || !loss_out) return fail(TL_ERR_ARG, "null pointer");
    if (!(q > 0.f)) return fail(TL_ERR_ARG, "loss_q must be positive");
    if (B_global <= 0) B_global = B;
    const Layout L = make_layout(B * C, H, W, feat_d, B);
    if (ws_bytes < L.total) return fail(TL_ERR_WORKSPACE, "workspace %zu < %zu bytes", ws_bytes, L.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const int call = g_timing.n_fwd;
    TL_MARK(call, 0, st);
    // pairs leave the persistence kernel in a deterministic (raster) order, so the matching needs no
    // sort; the segmented sort only serves tl_persistence_pairs (gudhi's emission order)
    rc = launch_ph(pred, truth, 2, L, H, W, feat_d, ws, st, false);
    if (rc != TL_OK) return rc;
    TL_MARK(call, 1, st);
    TL_MARK(call, 2, st);

    tl::MatchArgs m;
    m.d1 = tl::Diagrams{reinterpret_cast<const char*>(at<tl::PairRec>(ws, L.pairs[0])) + offsetof(tl::PairRec, b),
                        (int)sizeof(tl::PairRec), nullptr, at<int32_t>(ws, L.counts[0]), L.cap};
    m.d2 = tl::Diagrams{reinterpret_cast<const char*>(at<tl::PairRec>(ws, L.pairs[1])) + offsetof(tl::PairRec, b),
                        (int)sizeof(tl::PairRec), nullptr, at<int32_t>(ws, L.counts[1]), L.cap};
    m.n_diag = L.M; m.q = q; m.loss_r = loss_r;
    m.cost = at<double>(ws, L.cost); m.tpers = at<double>(ws, L.tpers);
    m.match1 = at<int32_t>(ws, L.match1);
    m.fill1 = at<tl::PairRec>(ws, L.pairs[0]);
    fill_match_scratch(m, ws, L.v, L.minv, L.u, L.way, L.pcol, L.used, L.stride_c, L.stride_r);
    m.counter = at<unsigned int>(ws, L.counter) + 32;  // byte 128 of the counter block launch_ph zeroed
    tl::match_kernel<<<L.M < kMatchSlots ? L.M : kMatchSlots, tl::kMatchThreads, 0, st>>>(m);
    TL_CUDA(cudaGetLastError());
    TL_MARK(call, 3, st);

    tl::LossArgs la;
    la.cost = m.cost; la.tpers = m.tpers; la.B = B; la.C = C; la.B_global = B_global; la.loss_r = loss_r;
    la.q = q; la.lamda = lamda; la.loss_out = loss_out; la.coef = at<double>(ws, L.coef);
    tl::loss_kernel<<<1, 256, 0, st>>>(la);
    TL_CUDA(cudaGetLastError());
    TL_MARK(call, 4, st);
    if (g_timing.on) ++g_timing.n_fwd;
    return TL_OK;
}

int tl_backward(const float* grad_loss, const void* ws, size_t ws_bytes, int B, int C, int H, int W,
                   int feat_d, float q, float lamda, int loss_r, int B_global, float* grad_pred, void* stream) {
    int rc = check_shape(B, C, H, W, feat_d);
    if (rc != TL_OK) return rc;
    if (!ws || !grad_pred) return fail(TL_ERR_ARG, "null pointer");
    if (B_global <= 0) B_global = B;
    const Layout L = make_layout(B * C, H, W, feat_d, B);
    if (ws_bytes < L.total) return fail(TL_ERR_WORKSPACE, "workspace %zu < %zu bytes", ws_bytes, L.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* w = const_cast<void*>(ws);
    const int call = g_timing.n_bwd;
    TL_MARK(call, 5, st);
    TL_CUDA(cudaMemsetAsync(grad_pred, 0, sizeof(float) * (size_t)B * C * H * W, st));
    TL_MARK(call, 6, st);
    tl::GradArgs g;
    g.pairs = at<tl::PairRec>(w, L.pairs[0]); g.counts = at<int32_t>(w, L.counts[0]);
    g.coef = at<double>(w, L.coef); g.grad_loss = grad_loss;
    g.M = L.M; g.C = C; g.cap = L.cap; g.N = H * W; g.B_global = B_global; g.loss_r = loss_r;
    g.q = q; g.lamda = lamda; g.grad_pred = grad_pred;
    tl::grad_kernel<<<L.M < 1184 ? L.M : 1184, 256, 0, st>>>(g);
    TL_CUDA(cudaGetLastError());
    TL_MARK(call, 7, st);
    if (g_timing.on) ++g_timing.n_bwd;
    return TL_OK;
}

}  // extern "C"

namespace {

__global__ void export_pairs_kernel(const tl::PairRec* recs, const int32_t* cnt, int n_maps, int cap_in,
                                    int32_t* pairs, int cap_out, int32_t* counts) {
    for (int map = blockIdx.x; map < n_maps; map += gridDim.x) {
        const int n = cnt[map];
        if (threadIdx.x == 0) counts[map] = n;
        const int lim = min(min(n, cap_in), cap_out);
        for (int i = threadIdx.x; i < lim; i += blockDim.x) {
            const tl::PairRec r = recs[(size_t)map * cap_in + i];
            pairs[((size_t)map * cap_out + i) * 2] = r.cre;
            pairs[((size_t)map * cap_out + i) * 2 + 1] = r.des;
        }
    }
}

}  // namespace

extern "C" {

int tl_persistence_pairs(const float* maps, int n_maps, int H, int W, int dim, void* ws, size_t ws_bytes,
                         int32_t* pairs, int cap, int32_t* counts, void* stream) {
    int rc = check_shape(n_maps, 1, H, W, dim);
    if (rc != TL_OK) return rc;
    if (!maps || !ws || !pairs || !counts || cap <= 0) return fail(TL_ERR_ARG, "null pointer or cap <= 0");
    const Layout L = make_layout(n_maps, H, W, dim, n_maps);
    if (ws_bytes < L.total) return fail(TL_ERR_WORKSPACE, "workspace %zu < %zu bytes", ws_bytes, L.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_ph(maps, nullptr, 1, L, H, W, dim, ws, st, true);
    if (rc != TL_OK) return rc;
    rc = launch_sort(1, L, ws, st);
    if (rc != TL_OK) return rc;
    export_pairs_kernel<<<n_maps < 1184 ? n_maps : 1184, 256, 0, st>>>(
        at<tl::PairRec>(ws, L.pairs[0]), at<int32_t>(ws, L.counts[0]), n_maps, L.cap, pairs, cap, counts);
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

int tl_wasserstein_workspace_bytes(int n_diag, int max_rows1, int max_rows2, size_t* bytes) {
    if (!bytes || n_diag <= 0 || max_rows1 < 0 || max_rows2 < 0) return fail(TL_ERR_ARG, "bad arguments");
    const size_t sc = align_up((size_t)max_rows1 + max_rows2 + 2, 32);
    const size_t sr = align_up((size_t)(max_rows1 < max_rows2 ? max_rows1 : max_rows2) + 2, 32);
    const int slots = n_diag < kMatchSlots ? n_diag : kMatchSlots;
    *bytes = align_up(sizeof(double) * sc * slots) * 2 + align_up(sizeof(double) * sr * slots) +
             align_up(sizeof(int32_t) * sc * slots) * 2 + align_up(sc * slots);
    return TL_OK;
}

int tl_wasserstein(const float* D1, const int32_t* off1, const float* D2, const int32_t* off2, int n_diag,
                   int max_rows1, int max_rows2, float q, void* ws, size_t ws_bytes, double* cost,
                   int32_t* match1, void* stream) {
    size_t need = 0;
    int rc = tl_wasserstein_workspace_bytes(n_diag, max_rows1, max_rows2, &need);
    if (rc != TL_OK) return rc;
    if (!D1 || !off1 || !D2 || !off2 || !ws || !cost || !match1) return fail(TL_ERR_ARG, "null pointer");
    if (!(q > 0.f)) return fail(TL_ERR_ARG, "q must be positive");
    if (ws_bytes < need) return fail(TL_ERR_WORKSPACE, "workspace %zu < %zu bytes", ws_bytes, need);
    const size_t sc = align_up((size_t)max_rows1 + max_rows2 + 2, 32);
    const size_t sr = align_up((size_t)(max_rows1 < max_rows2 ? max_rows1 : max_rows2) + 2, 32);
    const int slots = n_diag < kMatchSlots ? n_diag : kMatchSlots;
    size_t o = 0;
    auto take = [&](size_t b) { size_t a = o; o += align_up(b); return a; };
    const size_t ov = take(sizeof(double) * sc * slots), ominv = take(sizeof(double) * sc * slots);
    const size_t ou = take(sizeof(double) * sr * slots);
    const size_t oway = take(sizeof(int32_t) * sc * slots), opcol = take(sizeof(int32_t) * sc * slots);
    const size_t oused = take(sc * slots);
    tl::MatchArgs m;
    m.d1 = tl::Diagrams{reinterpret_cast<const char*>(D1), 8, off1, nullptr, 0};
    m.d2 = tl::Diagrams{reinterpret_cast<const char*>(D2), 8, off2, nullptr, 0};
    m.n_diag = n_diag; m.q = q; m.loss_r = 0; m.cost = cost; m.tpers = nullptr; m.match1 = match1; m.fill1 = nullptr;
    m.counter = nullptr;
    fill_match_scratch(m, ws, ov, ominv, ou, oway, opcol, oused, sc, sr);
    tl::match_kernel<<<slots, tl::kMatchThreads, 0, static_cast<cudaStream_t>(stream)>>>(m);
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

}  // extern "C"

extern "C" {

namespace {
int resample_args(tl::ResampleArgs& a, const float* in, int n_maps, int H, int W, int S, int apply_sigmoid) {
    if (!in) return fail(TL_ERR_ARG, "null pointer");
    if (n_maps <= 0 || H <= 0 || W <= 0 || S <= 0) return fail(TL_ERR_ARG, "bad shape [%d,%d,%d] -> %d", n_maps, H, W, S);
    if ((long long)n_maps * H * W >= (1ll << 40) || (long long)n_maps * S * S >= (1ll << 40)) return fail(TL_ERR_ARG, "too large");
    a.in = in; a.out = nullptr; a.gout = nullptr; a.gin = nullptr;
    a.n_maps = n_maps; a.H = H; a.W = W; a.S = S; a.apply_sigmoid = apply_sigmoid != 0;
    a.sy = S > 1 ? (float)(H - 1) / (float)(S - 1) : 0.f;
    a.sx = S > 1 ? (float)(W - 1) / (float)(S - 1) : 0.f;
    return TL_OK;
}
int grid_for(long long work, int block) {
    long long g = (work + block - 1) / block;
    const long long cap = 148ll * 16;  // a few waves of resident CTAs on a 148-SM B200, grid-stride beyond
    return (int)(g < 1 ? 1 : g > cap ? cap : g);
}
}  // namespace

int tl_resample_forward(const float* in, int n_maps, int H, int W, int S, int apply_sigmoid, float* out, void* stream) {
    tl::ResampleArgs a;
    int rc = resample_args(a, in, n_maps, H, W, S, apply_sigmoid);
    if (rc != TL_OK) return rc;
    if (!out) return fail(TL_ERR_ARG, "null pointer");
    a.out = out;
    tl::resample_fwd_kernel<<<grid_for((long long)n_maps * S * S, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

int tl_resample_backward(const float* grad_out, const float* in, int n_maps, int H, int W, int S, int apply_sigmoid,
                         float* grad_in, void* stream) {
    tl::ResampleArgs a;
    int rc = resample_args(a, in, n_maps, H, W, S, apply_sigmoid);
    if (rc != TL_OK) return rc;
    if (!grad_out || !grad_in) return fail(TL_ERR_ARG, "null pointer");
    if (reinterpret_cast<uintptr_t>(grad_in) & 15) return fail(TL_ERR_ARG, "grad_in must be 16-byte aligned");
    a.gout = grad_out; a.gin = grad_in;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n = (long long)n_maps * H * W;
    tl::zero_fill_kernel<<<grid_for(n >> 2, 256), 256, 0, st>>>(grad_in, n);
    TL_CUDA(cudaGetLastError());
    tl::resample_bwd_kernel<<<grid_for((long long)n_maps * S * S, 256), 256, 0, st>>>(a);
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

}  // extern "C"

extern "C" {

namespace {
int post_args(tl::PostArgs& a, int n_maps, int Hs, int Ws, int T, int rh, int rw, int oh, int ow) {
    if (n_maps <= 0 || Hs <= 0 || Ws <= 0 || T <= 0 || oh <= 0 || ow <= 0) return fail(TL_ERR_ARG, "bad shape");
    if (rh <= 0 || rw <= 0 || rh > T || rw > T) return fail(TL_ERR_ARG, "crop [%d,%d] outside the %dx%d intermediate", rh, rw, T, T);
    if ((long long)n_maps * oh * ow >= (1ll << 40)) return fail(TL_ERR_ARG, "too large");
    a.in = nullptr; a.out = nullptr; a.gout = nullptr; a.gin = nullptr;
    a.n_maps = n_maps; a.Hs = Hs; a.Ws = Ws; a.T = T; a.rh = rh; a.rw = rw; a.oh = oh; a.ow = ow;
    a.s1y = (float)Hs / (float)T; a.s1x = (float)Ws / (float)T;
    a.s2y = (float)rh / (float)oh; a.s2x = (float)rw / (float)ow;
    return TL_OK;
}
}  // namespace

int tl_postprocess_forward(const float* in, int n_maps, int Hs, int Ws, int T, int rh, int rw, int oh, int ow,
                           float* out, void* stream) {
    tl::PostArgs a;
    int rc = post_args(a, n_maps, Hs, Ws, T, rh, rw, oh, ow);
    if (rc != TL_OK) return rc;
    if (!in || !out) return fail(TL_ERR_ARG, "null pointer");
    a.in = in; a.out = out;
    const int ftx = (ow + tl::kFwdCols - 1) / tl::kFwdCols, fty = (oh + tl::kFwdRows - 1) / tl::kFwdRows;
    const long long f_tiles = (long long)n_maps * fty * ftx, f_cap = 148ll * 64;
    tl::postprocess_fwd_kernel<<<(int)(f_tiles < f_cap ? f_tiles : f_cap), tl::kFwdCols, 0, static_cast<cudaStream_t>(stream)>>>(a, fty, ftx);
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

int tl_postprocess_backward(const float* grad_out, int n_maps, int Hs, int Ws, int T, int rh, int rw, int oh, int ow,
                            float* grad_in, void* stream) {
    tl::PostArgs a;
    int rc = post_args(a, n_maps, Hs, Ws, T, rh, rw, oh, ow);
    if (rc != TL_OK) return rc;
    if (!grad_out || !grad_in) return fail(TL_ERR_ARG, "null pointer");
    if (reinterpret_cast<uintptr_t>(grad_in) & 15) return fail(TL_ERR_ARG, "grad_in must be 16-byte aligned");
    a.gout = grad_out; a.gin = grad_in;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n = (long long)n_maps * Hs * Ws;
    tl::zero_fill_kernel<<<grid_for(n >> 2, 256), 256, 0, st>>>(grad_in, n);
    TL_CUDA(cudaGetLastError());
    const int tiles_y = (oh + tl::kPostTileY - 1) / tl::kPostTileY, tiles_x = (ow + tl::kPostTileX - 1) / tl::kPostTileX;
    const long long n_tiles = (long long)n_maps * tiles_y * tiles_x;
    const long long cap = 148ll * 32;
    tl::postprocess_bwd_kernel<<<(int)(n_tiles < cap ? n_tiles : cap), 256, 0, st>>>(a, tiles_y, tiles_x);
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

}  // extern "C"
