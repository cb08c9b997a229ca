/*
 * Oracle F -- fast CPU restatement of the reference's topological loss.
 *
 * TEST INFRASTRUCTURE ONLY.  The product package must never link, load or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may, as the checker or the timed CPU arm.
 *
 * PARITY UNPINNED: the arithmetic of the reference path lives in third-party packages
 * that are neither vendored, pinned nor installed (torch_topological -> gudhi, POT;
 * SURVEY.md section 8c); the reference has no tests or golden vectors.  This file
 * restates their published algorithms and is pinned against oracle_literal.py (a
 * literal boundary-matrix reduction of the gudhi cell complex) and hand-derived KATs.
 * to_wasserstein and to_topo_loss (everything downstream of the pairs) ARE pinned against
 * the unmodified reference file run under torch autograd: tests/golden/orchestration_vectors.json
 * (tests/golden/make_golden_orchestration.py); the pairs of to_cubical_pairs are not.
 *
 * What each function follows:
 *   to_cubical_pairs      gudhi Bitmap_cubical_complex (T-construction, order
 *                         (value, dim, position)), Persistent_cohomology (strict
 *                         positive persistence), cofaces_of_persistence_pairs, as used by
 *                         torch_topological CubicalComplex._forward, called from
 *                         /root/reference/octsam/models/topological_loss.py:55-63.
 *                         Formulated as two Kruskal passes.
 *   to_wasserstein        torch_topological WassersteinDistance._make_distance_matrix +
 *                         POT ot.emd2 (exact optimum; gradient = optimal plan), called
 *                         from topological_loss.py:78-82.
 *   to_topo_loss          topological_loss.py:11-96 (orchestration, mean, lamda, loss_r)
 *                         plus the autograd backward of training_utils.py:66.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- keys */
static inline uint32_t mono32(float f) {
    uint32_t u;
    f = f + 0.0f; /* -0.0 -> +0.0: gudhi compares doubles, where they are equal */
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

typedef struct { uint64_t key; int32_t a, b; } edge_t; /* key = mono(value)<<32 | pos */

static int cmp_edge_asc(const void* x, const void* y) {
    uint64_t a = ((const edge_t*)x)->key, b = ((const edge_t*)y)->key;
    return a < b ? -1 : (a > b ? 1 : 0);
}

static int32_t uf_find(int32_t* p, int32_t x) {
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
}

static inline float fmin2(float a, float b) { return a < b ? a : b; }

/* value of the v-edge(i,j) between pixels (i,j-1),(i,j); j in [0,W] */
static inline float vedge_val(const float* f, int W, int i, int j) {
    if (j == 0) return f[i * W];
    if (j == W) return f[i * W + W - 1];
    return fmin2(f[i * W + j - 1], f[i * W + j]);
}
/* value of the h-edge(i,j) between pixels (i-1,j),(i,j); i in [0,H] */
static inline float hedge_val(const float* f, int H, int W, int i, int j) {
    if (i == 0) return f[j];
    if (i == H) return f[(H - 1) * W + j];
    return fmin2(f[(i - 1) * W + j], f[i * W + j]);
}
/* top(edge): left/upper adjacent pixel if it attains the edge value, else the other */
static inline int32_t vedge_top(const float* f, int W, int i, int j) {
    if (j == 0) return i * W;
    if (j == W) return i * W + W - 1;
    float v = fmin2(f[i * W + j - 1], f[i * W + j]);
    return f[i * W + j - 1] == v ? i * W + j - 1 : i * W + j;
}
static inline int32_t hedge_top(const float* f, int H, int W, int i, int j) {
    if (i == 0) return j;
    if (i == H) return (H - 1) * W + j;
    float v = fmin2(f[(i - 1) * W + j], f[i * W + j]);
    return f[(i - 1) * W + j] == v ? (i - 1) * W + j : i * W + j;
}
/* vertex (i,j), i in [0,H], j in [0,W]: value and first raster pixel attaining it */
static inline float vertex_val_top(const float* f, int H, int W, int i, int j, int32_t* top) {
    float best = INFINITY; int32_t bi = -1;
    for (int di = -1; di <= 0; ++di)
        for (int dj = -1; dj <= 0; ++dj) {
            int r = i + di, c = j + dj;
            if (r < 0 || r >= H || c < 0 || c >= W) continue;
            float v = f[r * W + c];
            if (bi < 0 || v < best) { best = v; bi = r * W + c; }
        }
    if (top) *top = bi;
    return best;
}

typedef struct { int32_t cre, des; uint64_t dkey; } pair_t;
static int cmp_pair(const void* x, const void* y) {
    uint64_t a = ((const pair_t*)x)->dkey, b = ((const pair_t*)y)->dkey;
    return a < b ? -1 : (a > b ? 1 : 0);
}

/*
 * Persistence pairs of one HxW map in homology dimension dim (0 or 1), as
 * (creator pixel, destroyer pixel) flat C-order indices, in gudhi's emission order
 * (filtration order of the death cell).  For dim 0 the essential class is appended
 * last, paired with argmax(f) (torch_topological's fake destroyer).
 * Returns the number of pairs, or -1 if cap is too small / bad arguments.
 */
int to_cubical_pairs(const float* f, int H, int W, int dim, int32_t* pairs, int cap) {
    if (H <= 0 || W <= 0 || (dim != 0 && dim != 1)) return -1;
    const int GW = 2 * W + 1;
    const int64_t nE = (int64_t)H * (W + 1) + (int64_t)(H + 1) * W;
    edge_t* E = (edge_t*)malloc(sizeof(edge_t) * (size_t)nE);
    pair_t* out = (pair_t*)malloc(sizeof(pair_t) * (size_t)(H * W + 1));
    int np = 0, ne = 0;
    if (dim == 0) {
        /* nodes = vertices (i,j) id i*(W+1)+j; v-edge joins (i,j),(i+1,j); h-edge joins (i,j),(i,j+1) */
        const int VW = W + 1, nV = (H + 1) * (W + 1);
        for (int i = 0; i < H; ++i)
            for (int j = 0; j <= W; ++j) {
                E[ne].key = ((uint64_t)mono32(vedge_val(f, W, i, j)) << 32) | (uint32_t)(2 * j + (2 * i + 1) * GW);
                E[ne].a = i * VW + j; E[ne].b = (i + 1) * VW + j; ++ne;
            }
        for (int i = 0; i <= H; ++i)
            for (int j = 0; j < W; ++j) {
                E[ne].key = ((uint64_t)mono32(hedge_val(f, H, W, i, j)) << 32) | (uint32_t)(2 * j + 1 + (2 * i) * GW);
                E[ne].a = i * VW + j; E[ne].b = i * VW + j + 1; ++ne;
            }
        qsort(E, (size_t)ne, sizeof(edge_t), cmp_edge_asc);
        int32_t* par = (int32_t*)malloc(sizeof(int32_t) * (size_t)nV);
        uint64_t* vkey = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nV);
        for (int i = 0; i <= H; ++i)
            for (int j = 0; j <= W; ++j) {
                int id = i * VW + j; par[id] = id;
                vkey[id] = ((uint64_t)mono32(vertex_val_top(f, H, W, i, j, NULL)) << 32) | (uint32_t)(2 * j + 2 * i * GW);
            }
        for (int e = 0; e < ne; ++e) {
            int32_t ra = uf_find(par, E[e].a), rb = uf_find(par, E[e].b);
            if (ra == rb) continue;
            int32_t dead = vkey[ra] > vkey[rb] ? ra : rb, live = dead == ra ? rb : ra;
            par[dead] = live; /* root = elder (smaller key) vertex */
            if ((E[e].key >> 32) > (vkey[dead] >> 32)) {
                int32_t pc, pd; uint32_t pos = (uint32_t)E[e].key;
                int Y = (int)(pos / GW), X = (int)(pos % GW);
                vertex_val_top(f, H, W, dead / VW, dead % VW, &pc);
                pd = (Y & 1) ? vedge_top(f, W, (Y - 1) / 2, X / 2) : hedge_top(f, H, W, Y / 2, (X - 1) / 2);
                out[np].cre = pc; out[np].des = pd; out[np].dkey = E[e].key; ++np;
            }
        }
        /* already in death-cell (edge) order.  Essential class: surviving root. */
        int32_t root = uf_find(par, 0), pc, amax = 0;
        vertex_val_top(f, H, W, root / VW, root % VW, &pc);
        for (int k = 1; k < H * W; ++k) if (f[k] > f[amax]) amax = k;
        out[np].cre = pc; out[np].des = amax; out[np].dkey = ~0ull; ++np;
        free(par); free(vkey);
    } else {
        /* nodes = squares id r*W+c, OUTSIDE = H*W (eldest); scan edges descending */
        const int OUT = H * W;
        for (int i = 0; i < H; ++i)
            for (int j = 0; j <= W; ++j) {
                E[ne].key = ((uint64_t)mono32(vedge_val(f, W, i, j)) << 32) | (uint32_t)(2 * j + (2 * i + 1) * GW);
                E[ne].a = j == 0 ? OUT : i * W + j - 1; E[ne].b = j == W ? OUT : i * W + j; ++ne;
            }
        for (int i = 0; i <= H; ++i)
            for (int j = 0; j < W; ++j) {
                E[ne].key = ((uint64_t)mono32(hedge_val(f, H, W, i, j)) << 32) | (uint32_t)(2 * j + 1 + (2 * i) * GW);
                E[ne].a = i == 0 ? OUT : (i - 1) * W + j; E[ne].b = i == H ? OUT : i * W + j; ++ne;
            }
        qsort(E, (size_t)ne, sizeof(edge_t), cmp_edge_asc);
        int32_t* par = (int32_t*)malloc(sizeof(int32_t) * (size_t)(OUT + 1));
        for (int k = 0; k <= OUT; ++k) par[k] = k;
        /* square key (mono(f), raster index): raster index is monotone in the bitmap position */
        #define SQKEY(k) ((k) == OUT ? ~0ull : (((uint64_t)mono32(f[k]) << 32) | (uint32_t)(k)))
        for (int e = ne - 1; e >= 0; --e) {
            int32_t ra = uf_find(par, E[e].a), rb = uf_find(par, E[e].b);
            if (ra == rb) continue;
            int32_t dead = SQKEY(ra) < SQKEY(rb) ? ra : rb, live = dead == ra ? rb : ra;
            par[dead] = live; /* root = elder (larger key) square */
            if ((uint32_t)(SQKEY(dead) >> 32) > (uint32_t)(E[e].key >> 32)) {
                uint32_t pos = (uint32_t)E[e].key;
                int Y = (int)(pos / GW), X = (int)(pos % GW);
                int32_t pc = (Y & 1) ? vedge_top(f, W, (Y - 1) / 2, X / 2) : hedge_top(f, H, W, Y / 2, (X - 1) / 2);
                out[np].cre = pc; out[np].des = dead;
                /* death cell = the square `dead`: bitmap position order == raster order */
                out[np].dkey = SQKEY(dead); ++np;
            }
        }
        #undef SQKEY
        qsort(out, (size_t)np, sizeof(pair_t), cmp_pair);
        free(par);
    }
    int rc = np;
    if (np > cap) rc = -1;
    else for (int k = 0; k < np; ++k) { pairs[2 * k] = out[k].cre; pairs[2 * k + 1] = out[k].des; }
    free(E); free(out);
    return rc;
}

/* ---------------------------------------------------------------- Wasserstein */
/* torch: vector_norm(D - 0.5*(x+y), inf) in fp32 */
static inline float diag_dist(float b, float d) {
    float h = 0.5f * (b + d);
    float v0 = fabsf(b - h), v1 = fabsf(d - h);
    return v0 > v1 ? v0 : v1;
}
static inline float powq(float x, double q) { return q == 2.0 ? x * x : powf(x, (float)q); }
static inline float linf(float b, float d, float b2, float d2) {
    float x = fabsf(b - b2), y = fabsf(d - d2);
    return x > y ? x : y;
}

/*
 * Exact optimal partial matching between D1 (n points, rows (b,d)) and D2 (m points) with
 * the diagonal, cost = POT emd2 on the (n+1)x(m+1) matrix of torch_topological.
 * match1[i] = index of the D2 point matched to D1[i], or -1 (diagonal).
 * Returns the optimal cost (fp64 accumulation of the fp32 matrix entries, as POT does).
 */
double to_wasserstein(const float* D1, int n, const float* D2, int m, double q, int32_t* match1) {
    float* a = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    float* t = (float*)malloc(sizeof(float) * (size_t)(m + 1));
    double total = 0.0;
    for (int i = 0; i < n; ++i) a[i] = powq(diag_dist(D1[2 * i], D1[2 * i + 1]), q);
    for (int j = 0; j < m; ++j) t[j] = powq(diag_dist(D2[2 * j], D2[2 * j + 1]), q);
    for (int i = 0; i < n; ++i) match1[i] = -1;
    /* rows = the smaller diagram; columns = the larger one plus one private diagonal slot per row */
    const int swap = n < m;
    const int R = swap ? n : m, Cn = swap ? m : n;
    const float* DR = swap ? D1 : D2; const float* DC = swap ? D2 : D1;
    const float* dr = swap ? a : t;   const float* dc = swap ? t : a;
    for (int k = 0; k < Cn; ++k) total += dc[k];
    if (R > 0) {
        const int NC = Cn + R;
        /* Hungarian (shortest augmenting paths) on R x NC, cost(r,k<Cn) = c(r,k) - dc[k]; cost(r,Cn+s) = dr[r] */
        double* u = (double*)calloc((size_t)R + 1, sizeof(double));
        double* v = (double*)calloc((size_t)NC + 1, sizeof(double));
        double* minv = (double*)malloc(sizeof(double) * ((size_t)NC + 1));
        int* p = (int*)calloc((size_t)NC + 1, sizeof(int));
        int* way = (int*)calloc((size_t)NC + 1, sizeof(int));
        char* used = (char*)malloc((size_t)NC + 1);
        for (int r = 1; r <= R; ++r) {
            p[0] = r; int j0 = 0;
            for (int k = 0; k <= NC; ++k) { minv[k] = INFINITY; used[k] = 0; }
            do {
                used[j0] = 1;
                int i0 = p[j0], j1 = 0; double delta = INFINITY;
                const float rb = DR[2 * (i0 - 1)], rd = DR[2 * (i0 - 1) + 1];
                for (int k = 1; k <= NC; ++k) {
                    if (used[k]) continue;
                    double c;
                    if (k <= Cn) c = (double)powq(linf(rb, rd, DC[2 * (k - 1)], DC[2 * (k - 1) + 1]), q) - (double)dc[k - 1];
                    else c = (double)dr[i0 - 1];
                    double cur = c - u[i0] - v[k];
                    if (cur < minv[k]) { minv[k] = cur; way[k] = j0; }
                    if (minv[k] < delta) { delta = minv[k]; j1 = k; }
                }
                for (int k = 0; k <= NC; ++k) {
                    if (used[k]) { u[p[k]] += delta; v[k] -= delta; }
                    else minv[k] -= delta;
                }
                j0 = j1;
            } while (p[j0] != 0);
            do { int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
        }
        for (int k = 1; k <= NC; ++k) {
            if (!p[k]) continue;
            int r = p[k] - 1;
            if (k <= Cn) {
                int c = k - 1;
                total += (double)powq(linf(DR[2 * r], DR[2 * r + 1], DC[2 * c], DC[2 * c + 1]), q) - (double)dc[c];
                if (swap) match1[r] = c; else match1[c] = r;
            } else total += (double)dr[r];
        }
        free(u); free(v); free(minv); free(p); free(way); free(used);
    }
    free(a); free(t);
    return total;
}

/* d cost / d (b,d) of one D1 point given its match (torch cdist(p=inf)/vector_norm/pow backward) */
static void point_grad(float b, float d, int matched, float b2, float d2, double q, double* gb, double* gd) {
    if (!matched) {
        float x = diag_dist(b, d);
        double g = x > 0.f ? q * pow((double)x, q - 1.0) : (q == 1.0 ? 1.0 : 0.0);
        /* vector_norm(inf) of (x - proj): derivative (-1/2,+1/2)*sign(d-b); zero vector -> 0 */
        double s = d > b ? 1.0 : (d < b ? -1.0 : 0.0);
        *gb = -0.5 * g * s; *gd = 0.5 * g * s;
        return;
    }
    float xb = b - b2, xd = d - d2, ab = fabsf(xb), ad = fabsf(xd);
    float dist = ab > ad ? ab : ad;
    double g = dist > 0.f ? q * pow((double)dist, q - 1.0) : 0.0;
    *gb = (ab == dist) ? g * (xb > 0 ? 1.0 : (xb < 0 ? -1.0 : 0.0)) : 0.0;
    *gd = (ad == dist) ? g * (xd > 0 ? 1.0 : (xd < 0 ? -1.0 : 0.0)) : 0.0;
}

/*
 * Whole loss, forward + backward, mirroring topological_loss.py:11-96 for interp = 0.
 * pred/truth: [B,C,H,W] fp32.  grad_pred (may be NULL): d loss / d pred, fully written.
 * pair_counts (may be NULL): [B*C*2] number of selected-dimension pairs (pred, truth).
 * nthreads <= 1 -> sequential over maps (the reference's Python loop); else OpenMP over maps.
 * Returns 0, or -1 on bad arguments.
 */
int to_topo_loss(const float* pred, const float* truth, int B, int C, int H, int W, int feat_d,
                 double q, double lamda, int loss_r, int nthreads,
                 float* loss_out, float* grad_pred, int32_t* pair_counts) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (feat_d != 0 && feat_d != 1)) return -1;
    const int N = H * W, M = B * C, cap = N + 1;
    double* cost = (double*)malloc(sizeof(double) * (size_t)M);
    double* tpers = (double*)calloc((size_t)M, sizeof(double));
    int32_t** P1 = (int32_t**)calloc((size_t)M, sizeof(int32_t*));
    int32_t** MT = (int32_t**)calloc((size_t)M, sizeof(int32_t*));
    int32_t** P2 = (int32_t**)calloc((size_t)M, sizeof(int32_t*));
    int* n1 = (int*)calloc((size_t)M, sizeof(int));
    int* n2 = (int*)calloc((size_t)M, sizeof(int));
    (void)nthreads;
#ifdef _OPENMP
    #pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 1 ? nthreads : 1)
#endif
    for (int k = 0; k < M; ++k) {
        const float* fp = pred + (size_t)k * N; const float* ft = truth + (size_t)k * N;
        int32_t* p1 = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)cap);
        int32_t* p2 = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)cap);
        int a = to_cubical_pairs(fp, H, W, feat_d, p1, cap);
        int b = to_cubical_pairs(ft, H, W, feat_d, p2, cap);
        float* D1 = (float*)malloc(sizeof(float) * 2 * (size_t)(a + 1));
        float* D2 = (float*)malloc(sizeof(float) * 2 * (size_t)(b + 1));
        for (int i = 0; i < a; ++i) { D1[2 * i] = fp[p1[2 * i]]; D1[2 * i + 1] = fp[p1[2 * i + 1]]; }
        for (int j = 0; j < b; ++j) { D2[2 * j] = ft[p2[2 * j]]; D2[2 * j + 1] = ft[p2[2 * j + 1]]; }
        int32_t* mt = (int32_t*)malloc(sizeof(int32_t) * (size_t)(a + 1));
        cost[k] = to_wasserstein(D1, a, D2, b, q, mt);
        if (loss_r) for (int i = 0; i < a; ++i) tpers[k] += pow(fabs((double)(D1[2 * i + 1] - D1[2 * i])), q);
        free(D1); free(D2);
        P1[k] = p1; P2[k] = p2; MT[k] = mt; n1[k] = a; n2[k] = b;
        if (pair_counts) { pair_counts[2 * k] = a; pair_counts[2 * k + 1] = b; }
    }
    /* per image: S_b = sum_c fp32(cost_c) in fp32 (emd2 returns the dtype of M), W_b = S_b^(1/q) */
    float loss = 0.f;
    float* Sb = (float*)malloc(sizeof(float) * (size_t)B);
    for (int b = 0; b < B; ++b) {
        float S = 0.f;
        for (int c = 0; c < C; ++c) S += (float)cost[b * C + c];
        Sb[b] = S;
        loss += q == 2.0 ? sqrtf(S) : powf(S, (float)(1.0 / q));
    }
    loss /= (float)B;
    if (loss_r) {
        float reg = 0.f;
        for (int k = 0; k < M; ++k) reg += (float)tpers[k];
        loss += reg / (float)M;
    }
    *loss_out = (float)lamda * loss;
    if (grad_pred) {
        memset(grad_pred, 0, sizeof(float) * (size_t)M * N);
        for (int k = 0; k < M; ++k) {
            const int b = k / C;
            const float* fp = pred + (size_t)k * N; const float* ft = truth + (size_t)k * N;
            float* g = grad_pred + (size_t)k * N;
            /* d loss / d S_b; S_b == 0 gives inf * 0 = NaN on every diagram entry, as autograd does */
            double gS = lamda / B * (1.0 / q) * pow((double)Sb[b], 1.0 / q - 1.0);
            const int degenerate = !(Sb[b] > 0.f);
            for (int i = 0; i < n1[k]; ++i) {
                int pc = P1[k][2 * i], pd = P1[k][2 * i + 1];
                float bb = fp[pc], dd = fp[pd];
                double gb, gd;
                int j = MT[k][i];
                if (j >= 0) point_grad(bb, dd, 1, ft[P2[k][2 * j]], ft[P2[k][2 * j + 1]], q, &gb, &gd);
                else point_grad(bb, dd, 0, 0.f, 0.f, q, &gb, &gd);
                if (degenerate) { gb = NAN; gd = NAN; } else { gb *= gS; gd *= gS; }
                if (loss_r) {
                    double pers = (double)dd - (double)bb, ap = fabs(pers);
                    double gr = ap > 0 ? q * pow(ap, q - 1.0) * (pers > 0 ? 1.0 : -1.0) : 0.0;
                    gr *= lamda / M;
                    gb -= gr; gd += gr;
                }
                g[pc] += (float)gb; g[pd] += (float)gd;
            }
        }
    }
    for (int k = 0; k < M; ++k) { free(P1[k]); free(P2[k]); free(MT[k]); }
    free(P1); free(P2); free(MT); free(n1); free(n2); free(cost); free(tpers); free(Sb);
    return 0;
}

int to_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
