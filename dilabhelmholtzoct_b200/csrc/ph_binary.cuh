// Two-valued maps (one-hot ground truth, thresholded masks): H1 pairs without a merge tree.
//
// For a map that takes exactly two values lo < hi the gudhi pairing that the generic path computes
// (sublevel filtration, T-construction, cell order (value, dim, position), strict positive
// persistence) collapses to a connected-component problem:
//
//   * the H1 classes with positive persistence are the 4-connected components ("blobs") of hi
//     pixels that do not touch the image border (a blob on the border merges into OUTSIDE through a
//     boundary edge of value hi: zero persistence);
//   * in the descending dual scan every hi-hi edge comes before every lo edge, so a blob is complete
//     before its first lo edge arrives, and its root (eldest square) is its LAST pixel in raster order;
//   * lo edges are scanned by descending bitmap position, i.e. bottom-up: the first lo edge that
//     joins the blob to an elder component is the h-edge under that last pixel (the pixel below is lo,
//     and straight down from it the scan has already connected everything to OUTSIDE or to a blob with a
//     later last pixel).  Its coface walk lands on the lower pixel (the upper one does not attain lo).
//
// Hence  destroyer = last raster pixel p of the blob,  creator = p + W,  emitted in raster order of p
// -- verified against the oracle on random two-valued maps (tests/test_oracle.py) and on the GPU
// (tests/test_gpu_parity.py).  The blobs are labelled on a 1-bit-per-pixel mask with a run-based
// union-find (one 16-bit entry per horizontal run); anything that does not fit (more than two values,
// too many runs, mask larger than shared memory) returns 0 and the caller takes the generic path.
#pragma once
#include "tl_common.cuh"

namespace tl {

struct BinView {
    const uint32_t* m;   // [H][nw] hi-pixel mask
    int nw;
    // run ENDS inside word (r, k): a hi pixel whose right neighbour is not hi
    __device__ __forceinline__ uint32_t ends(int wi, int k) const {
        const uint32_t w = m[wi];
        const uint32_t nxt = (k + 1 < nw) ? (m[wi + 1] & 1u) : 0u;
        return w & ~((w >> 1) | (nxt << 31));
    }
};

__device__ __forceinline__ uint32_t bin_find(volatile uint16_t* par, uint32_t x) {
    uint32_t px = par[x];
    while (px != x) {
        const uint32_t gp = par[px];
        if (gp != px) par[x] = (uint16_t)gp;  // path halving: entries only ever hold ancestors
        x = px; px = gp;
    }
    return x;
}

// All threads of the CTA call this (block-uniform).  Returns 1 when the map was handled (pairs and
// count written), 0 when the caller must run the generic path (nothing has been written then).
__device__ __noinline__ int binary_h1_pairs(const float* __restrict__ f, int H, int W, unsigned char* smem, int smem_bytes,
                                            const PairStore& ps, int set, int map, unsigned long long* prof) {
    __shared__ unsigned int sb_omin, sb_omax;
    __shared__ int sb_wsum[32];
    __shared__ int sb_total, sb_avail;
    __shared__ unsigned long long sb_base;
    __shared__ double sb_red[32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int nw = (W + 31) >> 5, n_words = H * nw;
    // layout: mask u32[n_words] | pre u16[n_words] | par u16[R_cap] | border u8[R_cap]
    const int fixed = ((n_words * 6 + 15) & ~15);
    if (fixed + 3 * 64 > smem_bytes || n_words >= (1 << 22)) return 0;
    int r_cap = (smem_bytes - fixed) / 3;
    if (r_cap > 65535) r_cap = 65535;
    uint32_t* m = reinterpret_cast<uint32_t*>(smem);
    uint16_t* pre = reinterpret_cast<uint16_t*>(smem + 4 * (size_t)n_words);
    uint16_t* par = reinterpret_cast<uint16_t*>(smem + fixed);
    unsigned char* border = smem + fixed + 2 * (size_t)r_cap;
    const FastDiv divNW((uint32_t)nw);
    long long t0 = prof ? clock64() : 0;
#define TLB_PROF(slot) do { if (prof && tid == 0) { const long long t1 = clock64(); atomicAdd(prof + (slot), (unsigned long long)(t1 - t0)); t0 = t1; } } while (0)

    if (tid == 0) { sb_omin = 0xFFFFFFFFu; sb_omax = 0u; }
    __syncthreads();
    // ---- mask of the pixels that differ from pixel 0, and the min / max of those "other" values.  One
    //      word per warp per trip, 4 trips in flight; after the first 8 trips a block vote throws out
    //      maps with more than two values (every prediction map) before the rest of the map is read
    const uint32_t ref = mono32(__ldg(f));
    uint32_t omin = 0xFFFFFFFFu, omax = 0u;
    // vector path (rows of whole 128-pixel groups, 16-byte aligned map): a lane reads 4 pixels per trip, a warp
    // 128 pixels = 4 mask words; otherwise one pixel per lane, one word per warp per trip.  The loads of a
    // group of trips are issued unconditionally (clamped addresses) BEFORE any of them is used, so a
    // thread has 4 requests in flight instead of one.
    const bool vec = (W & 127) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0;
    const int per_trip = vec ? 4 : 1;                       // words per warp per trip
    const int n_it = (n_words + nwarp * per_trip - 1) / (nwarp * per_trip);
    auto scan_words = [&](int it0, int it1) {
        if (vec) {
            const float4* f4 = reinterpret_cast<const float4*>(f);
            const int n_quads = n_words * 8;
            for (int it = it0; it < it1; it += 4) {
                float4 raw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = (warp + (it + u) * nwarp) * 32 + lane;
                    raw[u] = __ldg(f4 + min(q, n_quads - 1));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int wi = (warp + (it + u) * nwarp) * 4;
                    const bool ok = it + u < it1 && wi < n_words;  // warp-uniform
                    const uint32_t v0 = mono32(raw[u].x), v1 = mono32(raw[u].y), v2 = mono32(raw[u].z), v3 = mono32(raw[u].w);
                    if (ok && ((raw[u].x != raw[u].x) | (raw[u].y != raw[u].y) | (raw[u].z != raw[u].z) | (raw[u].w != raw[u].w))) { omin = 0u; omax = 0xFFFFFFFFu; }  // NaN: leave it to the generic path
                    unsigned nib = 0u;
                    if (ok) {
                        if (v0 != ref) { nib |= 1u; omin = min(omin, v0); omax = max(omax, v0); }
                        if (v1 != ref) { nib |= 2u; omin = min(omin, v1); omax = max(omax, v1); }
                        if (v2 != ref) { nib |= 4u; omin = min(omin, v2); omax = max(omax, v2); }
                        if (v3 != ref) { nib |= 8u; omin = min(omin, v3); omax = max(omax, v3); }
                    }
                    unsigned w = nib << (4 * (lane & 7));
                    w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
                    w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
                    w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
                    if (ok && (lane & 7) == 0) m[wi + (lane >> 3)] = w;
                }
            }
            return;
        }
        for (int it = it0; it < it1; it += 8) {
            float raw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int wi = min(warp + (it + u) * nwarp, n_words - 1);
                const int r = (int)divNW.div((uint32_t)wi), k = wi - r * nw;
                raw[u] = __ldg(f + (size_t)r * W + min(32 * k + lane, W - 1));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int wi = warp + (it + u) * nwarp;
                const bool okw = it + u < it1 && wi < n_words;  // warp-uniform
                const int r = (int)divNW.div((uint32_t)min(wi, n_words - 1)), k = min(wi, n_words - 1) - r * nw;
                const uint32_t v = mono32(raw[u]);
                const bool other = okw && 32 * k + lane < W && v != ref;
                if (other) { omin = min(omin, v); omax = max(omax, v); }
                if (okw && 32 * k + lane < W && raw[u] != raw[u]) { omin = 0u; omax = 0xFFFFFFFFu; }  // NaN: leave it to the generic path
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, other);
                if (lane == 0 && okw) m[wi] = bal;
            }
        }
    };
    auto vote = [&]() {
        const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, omin), hi = __reduce_max_sync(0xFFFFFFFFu, omax);
        if (lane == 0 && lo <= hi) { atomicMin(&sb_omin, lo); atomicMax(&sb_omax, hi); }
        __syncthreads();
        const bool many = sb_omin < sb_omax;  // at least two values besides pixel 0's
        __syncthreads();                      // block-uniform: nobody updates the pair before everyone has read it
        return many;
    };
    const int it_probe = n_it < (vec ? 4 : 8) ? n_it : (vec ? 4 : 8);
    scan_words(0, it_probe);
    if (vote()) return 0;
    if (it_probe < n_it) {
        scan_words(it_probe, n_it);
        if (vote()) return 0;
    }
    TLB_PROF(1);  // mask scan
    const uint32_t other = sb_omin;
    if (other > sb_omax) {  // no other value: constant map, no finite H1 pair
        if (tid == 0) { int av; ps_reserve(ps, set, map, 0, &av); ps.dsum[set][map] = 0.0; }
        return 1;
    }
    if (other < ref) {  // pixel 0 carries hi: flip the mask (bits past column W stay clear)
        for (int wi = tid; wi < n_words; wi += nt) {
            const int r = (int)divNW.div((uint32_t)wi), k = wi - r * nw;
            const uint32_t valid = (32 * k + 32 <= W) ? 0xFFFFFFFFu : ((1u << (W - 32 * k)) - 1u);
            m[wi] = ~m[wi] & valid;
        }
        __syncthreads();
    }
    BinView V{m, nw};

    // ---- runs in raster order: pre[word] = runs that END before this word
    const int cw = (n_words + nt - 1) / nt;  // contiguous words per thread
    const int w_beg = min(n_words, tid * cw), w_end = min(n_words, w_beg + cw);
    auto block_excl_scan = [&](int mine, int& total) {
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        __syncthreads();  // earlier readers of sb_wsum / sb_total are done
        if (lane == 31) sb_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = lane < nwarp ? sb_wsum[lane] : 0;
            int s = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += t; }
            sb_wsum[lane] = s - v;
            if (lane == 31) sb_total = s;
        }
        __syncthreads();
        total = sb_total;
        return sb_wsum[warp] + incl - mine;
    };
    int n_runs = 0;
    {
        int mine = 0;
        for (int wi = w_beg; wi < w_end; ++wi) {
            const int r = (int)divNW.div((uint32_t)wi), k = wi - r * nw;
            mine += __popc(V.ends(wi, k));
        }
        int run = block_excl_scan(mine, n_runs);
        if (n_runs > r_cap) return 0;  // block-uniform; pre[] not written yet, nothing to undo
        for (int wi = w_beg; wi < w_end; ++wi) {
            const int r = (int)divNW.div((uint32_t)wi), k = wi - r * nw;
            pre[wi] = (uint16_t)run;
            run += __popc(V.ends(wi, k));
        }
    }
    TLB_PROF(2);  // runs
    for (int x = tid; x < n_runs; x += nt) { par[x] = (uint16_t)x; border[x] = 0; }
    __syncthreads();
    auto run_of = [&](int wi, int k, int bit) {  // run that contains hi pixel (word wi, bit)
        return (uint32_t)pre[wi] + (uint32_t)__popc(V.ends(wi, k) & ((1u << bit) - 1u));
    };

    // ---- vertical contacts: one union per maximal horizontal stretch of hi-over-hi pixels; the root of a
    //      blob is its LARGEST run index, the run that holds its last raster pixel
    for (int wi = w_beg; wi < w_end; ++wi) {
        const int r = (int)divNW.div((uint32_t)wi), k = wi - r * nw;
        if (r == 0) continue;
        const uint32_t v = m[wi] & m[wi - nw];
        if (!v) continue;
        const uint32_t carry = k > 0 ? ((m[wi - 1] & m[wi - 1 - nw]) >> 31) : 0u;
        uint32_t s = v & ~((v << 1) | carry);
        while (s) {
            const int bit = __ffs(s) - 1;
            s &= s - 1;
            uint32_t a = run_of(wi, k, bit), b = run_of(wi - nw, k, bit);
            for (;;) {
                uint32_t ra = bin_find(par, a), rb = bin_find(par, b);
                if (ra == rb) break;
                if (ra < rb) { const uint32_t t = ra; ra = rb; rb = t; }
                const unsigned short old = atomicCAS(reinterpret_cast<unsigned short*>(par + rb), (unsigned short)rb, (unsigned short)ra);
                if (old == (unsigned short)rb) break;
                a = ra; b = rb;
            }
        }
    }
    __syncthreads();
    TLB_PROF(3);  // unions
    // ---- blobs on the image border: first / last row (every run), first / last column
    for (int wi = tid; wi < 2 * nw; wi += nt) {
        const int k = wi < nw ? wi : wi - nw, w2 = wi < nw ? wi : (H - 1) * nw + k;
        if (wi >= nw && H == 1) continue;
        uint32_t e = V.ends(w2, k);
        uint32_t x = pre[w2];
        while (e) { e &= e - 1; border[bin_find(par, x)] = 1; ++x; }
    }
    for (int r = tid; r < H; r += nt) {
        const int w0 = r * nw, w1 = r * nw + nw - 1, lb = (W - 1) & 31;
        if (m[w0] & 1u) border[bin_find(par, run_of(w0, 0, 0))] = 1;
        if ((m[w1] >> lb) & 1u) border[bin_find(par, run_of(w1, nw - 1, lb))] = 1;
    }
    __syncthreads();
    TLB_PROF(4);  // border
    // ---- emit: one pair per root run that is not on the border, in raster order of its last pixel
    int total = 0;
    {
        int mine = 0;
        for (int wi = w_beg; wi < w_end; ++wi) {
            const int r = (int)divNW.div((uint32_t)wi), k = wi - r * nw;
            uint32_t e = V.ends(wi, k);
            uint32_t x = pre[wi];
            while (e) { e &= e - 1; if (par[x] == (uint16_t)x && !border[x]) ++mine; ++x; }
        }
        int slot = block_excl_scan(mine, total);
        if (tid == 0) { int av; sb_base = ps_reserve(ps, set, map, total, &av); sb_avail = av; }
        __syncthreads();
        PairRec* out = ps.arena + sb_base;
        uint64_t* skeys = ps.skeys ? ps.skeys + sb_base : nullptr;
        const int avail = sb_avail;
        double dacc = 0.0;
        if (mine) {
            for (int wi = w_beg; wi < w_end; ++wi) {
                const int r = (int)divNW.div((uint32_t)wi), k = wi - r * nw;
                uint32_t e = V.ends(wi, k);
                uint32_t x = pre[wi];
                while (e) {
                    const int bit = __ffs(e) - 1;
                    e &= e - 1;
                    if (par[x] == (uint16_t)x && !border[x]) {
                        if (slot < avail) {
                            PairRec rec;
                            rec.des = r * W + 32 * k + bit;
                            rec.cre = rec.des + W;  // a blob off the border never reaches the last row
                            rec.b = __ldg(f + rec.cre); rec.d = __ldg(f + rec.des);
                            rec.tb = rec.td = __int_as_float(0x7FC00000);
                            out[slot] = rec;
                            if (skeys) skeys[slot] = ((uint64_t)mono32(rec.d) << 32) | (uint32_t)rec.des;
                            dacc += (double)cost_diag(rec.b, rec.d, ps.q);
                        }
                        ++slot;
                    }
                    ++x;
                }
            }
        }
        dacc = block_sum(dacc, sb_red);
        if (tid == 0) ps.dsum[set][map] = dacc;
    }
    TLB_PROF(5);  // emit
#undef TLB_PROF
    return 1;
}

}  // namespace tl
