/*
 * oracle.c — CPU restatement of the reference BiCGSTAB hot path under the shared
 * arithmetic spec (see oracle.h header and DESIGN.md §3).  TEST INFRASTRUCTURE ONLY.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile). Contraction is
 * off so every rounding below is exactly what is written; fma() is an explicit call.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#define ORC_CLONES __attribute__((target_clones("fma", "default")))
#else
#define ORC_CLONES
#endif

static inline double FMA(double a, double b, double c) { return __builtin_fma(a, b, c); }

/* ---------------------------------------------------------------------------------
 * R(): the lane-strided reduction used at every level of the tree.
 * 32 lane accumulators; lane l adds values l, l+32, ... in order; then a butterfly
 * (xor 16,8,4,2,1) in which every lane computes acc[l] + acc[l^s]; result = lane 0.
 * ------------------------------------------------------------------------------- */
static double butterfly(double acc[32]) {
    double t[32];
    for (int s = 16; s >= 1; s >>= 1) {
        for (int l = 0; l < 32; ++l) t[l] = acc[l] + acc[l ^ s];
        memcpy(acc, t, sizeof t);
    }
    return acc[0];
}

ORC_CLONES
double orc_reduce_values(int64_t m, const double *v) {
    double acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0;
    for (int64_t j = 0; j < m; ++j) acc[j & 31] = acc[j & 31] + v[j];
    return butterfly(acc);
}

/* A slab is 32 consecutive elements (one per lane): products are rounded, then butterflied.
 * A tile is 64 consecutive slabs (2048 elements): R() over its slab sums.
 * A group is 1024 consecutive tiles (2 Mi elements): R() over its tile sums.
 * The result is R() over the group sums.  Slab/tile/group boundaries are multiples of
 * 32/2048/2Mi in the GLOBAL index, so the tree is independent of grid and of row sharding. */
ORC_CLONES
static double dot_slab(int64_t m, const double *a, const double *b) {
    double acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = (l < m) ? a[l] * b[l] : 0.0;
    return butterfly(acc);
}

ORC_CLONES
static double dot_tile(int64_t m, const double *a, const double *b) {
    double ss[ORC_TILE_SLABS];
    int ns = (int)((m + 31) / 32);
    for (int q = 0; q < ns; ++q) {
        int64_t lo = (int64_t)q * 32, c = m - lo < 32 ? m - lo : 32;
        ss[q] = dot_slab(c, a + lo, b + lo);
    }
    return orc_reduce_values(ns, ss);
}

void orc_dot_tiles(int64_t n, const double *a, const double *b, double *out) {
    int64_t nt = (n + ORC_TILE - 1) / ORC_TILE;
    for (int64_t q = 0; q < nt; ++q) {
        int64_t lo = q * ORC_TILE, m = n - lo < ORC_TILE ? n - lo : ORC_TILE;
        out[q] = dot_tile(m, a + lo, b + lo);
    }
}

double orc_combine_tiles(int64_t ntile, const double *tile) {
    if (ntile <= 0) return 0.0;
    int64_t ngroup = (ntile + ORC_GROUP - 1) / ORC_GROUP;
    double *g = (double *)malloc(sizeof(double) * (size_t)ngroup);
    for (int64_t q = 0; q < ngroup; ++q) {
        int64_t lo = q * ORC_GROUP, m = ntile - lo < ORC_GROUP ? ntile - lo : ORC_GROUP;
        g[q] = orc_reduce_values(m, tile + lo);
    }
    double r = orc_reduce_values(ngroup, g);
    free(g);
    return r;
}

double orc_dot(int64_t n, const double *a, const double *b) {
    if (n <= 0) return 0.0;
    int64_t nt = (n + ORC_TILE - 1) / ORC_TILE;
    double *tile = (double *)malloc(sizeof(double) * (size_t)nt);
    orc_dot_tiles(n, a, b, tile);
    double r = orc_combine_tiles(nt, tile);
    free(tile);
    return r;
}

/* ---------------------------------------------------------------------------------
 * Row sum spec.  len <= ORC_LONG_ROW: one sequential FMA chain from +0.0 in storage
 * (ascending column) order.  Longer rows: 32 interleaved chains + butterfly.
 * ------------------------------------------------------------------------------- */
ORC_CLONES
static double rowsum(const double *a, const int *ja, int len, int base, const double *x) {
    if (len <= ORC_LONG_ROW) {
        double acc = 0.0;
        for (int k = 0; k < len; ++k) acc = FMA(a[k], x[ja[k] - base], acc);
        return acc;
    }
    double acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0;
    for (int k = 0; k < len; ++k) acc[k & 31] = FMA(a[k], x[ja[k] - base], acc[k & 31]);
    return butterfly(acc);
}

ORC_CLONES
void orc_spmv(int n, const int *ia, const int *ja, const double *a,
              const double *x, const double *d, double *y) {
    int base = ia[0];
    for (int i = 0; i < n; ++i) {
        int s = ia[i] - base, e = ia[i + 1] - base;
        double v = rowsum(a + s, ja + s, e - s, base, x);
        if (d) v = v + d[i] * x[i];      /* fl(rowsum + fl(d*x)): mult_spec then csrmv beta=1 */
        y[i] = v;
    }
}

/* ---------------------------------------------------------------------------------
 * ILU(0), IKJ order, no pivoting, pattern of A, unit-L / U-with-diagonal in one array.
 * ------------------------------------------------------------------------------- */
ORC_CLONES
int orc_ilu0(int n, const int *ia, const int *ja, const double *a, double *M) {
    int base = ia[0];
    int64_t nnz = ia[n] - base;
    memcpy(M, a, sizeof(double) * (size_t)nnz);
    int *diag = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int status = 0;
    for (int i = 0; i < n; ++i) {
        diag[i] = -1;
        for (int p = ia[i] - base; p < ia[i + 1] - base; ++p)
            if (ja[p] - base == i) { diag[i] = p; break; }
        if (diag[i] < 0 && status == 0) status = 1 + i;
    }
    if (status > 0) { free(diag); return status; }
    for (int i = 0; i < n; ++i) {
        int rs = ia[i] - base, re = ia[i + 1] - base;
        for (int p = rs; p < re; ++p) {
            int k = ja[p] - base;
            if (k >= i) break;
            double piv = M[diag[k]];
            double l = M[p] / piv;
            M[p] = l;
            /* row i -= l * (upper part of row k), restricted to row i's pattern */
            int q = p + 1;
            for (int pk = diag[k] + 1; pk < ia[k + 1] - base; ++pk) {
                int j = ja[pk];
                while (q < re && ja[q] < j) ++q;
                if (q >= re) break;
                if (ja[q] == j) M[q] = FMA(-l, M[pk], M[q]);
            }
        }
        if (M[diag[i]] == 0.0 && status == 0) status = -(1 + i);
    }
    free(diag);
    return status;
}

ORC_CLONES
void orc_sptrsv_lower_unit(int n, const int *ia, const int *ja, const double *M,
                           const double *rhs, double *out) {
    int base = ia[0];
    for (int i = 0; i < n; ++i) {
        double acc = rhs[i];
        for (int p = ia[i] - base; p < ia[i + 1] - base; ++p) {
            int k = ja[p] - base;
            if (k >= i) break;
            acc = FMA(-M[p], out[k], acc);
        }
        out[i] = acc;
    }
}

ORC_CLONES
void orc_sptrsv_upper(int n, const int *ia, const int *ja, const double *M,
                      const double *rhs, double *out) {
    int base = ia[0];
    for (int i = n - 1; i >= 0; --i) {
        double acc = rhs[i];
        double dg = 0.0;
        for (int p = ia[i] - base; p < ia[i + 1] - base; ++p) {
            int k = ja[p] - base;
            if (k < i) continue;
            if (k == i) { dg = M[p]; continue; }
            acc = FMA(-M[p], out[k], acc);
        }
        out[i] = acc / dg;
    }
}

int orc_levels(int n, const int *ia, const int *ja, int upper, int *level) {
    int base = ia[0], nl = 0;
    if (!upper) {
        for (int i = 0; i < n; ++i) {
            int lv = 0;
            for (int p = ia[i] - base; p < ia[i + 1] - base; ++p) {
                int k = ja[p] - base;
                if (k >= i) break;
                if (level[k] + 1 > lv) lv = level[k] + 1;
            }
            level[i] = lv;
            if (lv + 1 > nl) nl = lv + 1;
        }
    } else {
        for (int i = n - 1; i >= 0; --i) {
            int lv = 0;
            for (int p = ia[i + 1] - base - 1; p >= ia[i] - base; --p) {
                int k = ja[p] - base;
                if (k <= i) break;
                if (level[k] + 1 > lv) lv = level[k] + 1;
            }
            level[i] = lv;
            if (lv + 1 > nl) nl = lv + 1;
        }
    }
    return nl;
}

/* ---------------------------------------------------------------------------------
 * Unpreconditioned BiCGSTAB — restates gpu_pbicgstab2 (shifted) pbicgstab.cu:581-754.
 * Element-wise forms are the reference's scal/axpy chains: every product and every sum
 * is rounded separately (axpy with alpha=1 is a plain add).
 * ------------------------------------------------------------------------------- */
ORC_CLONES
int orc_bicgstab_unprec(int n, const int *ia, const int *ja, const double *a,
                        const double *d, const double *x0in, const double *b,
                        int maxit, double tol, double *x, orc_stats *st,
                        double *hist, int hist_cap) {
    size_t N = (size_t)(n > 0 ? n : 1);
    double *w = (double *)calloc(N * 8, sizeof(double));
    double *r0 = w, *r = w + N, *v = w + 2 * N, *p = w + 3 * N, *s = w + 4 * N,
           *t = w + 5 * N, *x0 = w + 6 * N, *h = w + 7 * N;
    double omega = 1, alpha = 1, beta, rho = 1, rho_;
    int nh = 0, ret = 0;
    memset(st, 0, sizeof *st);
    for (int i = 0; i < n; ++i) { x0[i] = x0in ? x0in[i] : 1.0; x[i] = 0.0; }
    /* r = b - (A0 + diag d) x0   (:645-649);  r0 = r (:652) */
    orc_spmv(n, ia, ja, a, x0, d, r);
    for (int i = 0; i < n; ++i) { r[i] = b[i] - r[i]; r0[i] = r[i]; }
    double norm0 = sqrt(orc_dot(n, r, r));                      /* :655 */
    double norm = norm0;
    st->nrm_r0 = norm0;
    if (hist && nh < hist_cap) hist[nh++] = norm0;
    st->breakdown = 3;
    int it = 0;
    for (it = 0; it < maxit; ++it) {
        rho_ = orc_dot(n, r0, r);                               /* :665 */
        beta = (rho_ / rho) * (alpha / omega);                  /* :666 */
        double momega = -omega;
        for (int i = 0; i < n; ++i) {                           /* :668-672 */
            double q = momega * v[i];
            q = p[i] + q;
            q = beta * q;
            p[i] = r[i] + q;
        }
        orc_spmv(n, ia, ja, a, p, d, v);                        /* :675-676 */
        double rv = orc_dot(n, r0, v);                          /* :688 */
        alpha = rho_ / rv;
        double malpha = -alpha;
        for (int i = 0; i < n; ++i) {                           /* :694-700 */
            double q = alpha * p[i];
            h[i] = x0[i] + q;
            q = malpha * v[i];
            s[i] = r[i] + q;
        }
        orc_spmv(n, ia, ja, a, s, d, t);                        /* :703-704 */
        double num = orc_dot(n, t, s), den = orc_dot(n, t, t);  /* :708-709 */
        omega = num / den;
        momega = -omega;
        for (int i = 0; i < n; ++i) {                           /* :714-720 */
            double q = omega * s[i];
            x[i] = h[i] + q;
            q = momega * t[i];
            r[i] = s[i] + q;                                    /* r_ ; rotated at :744 */
        }
        norm = sqrt(orc_dot(n, r, r));                          /* :723 */
        if (hist && nh < hist_cap) hist[nh++] = norm;
        if (norm < tol * norm0) { st->converged = 1; st->breakdown = 0; ++it; ret = 1; break; }
        if (fabs(omega) < 1e-5 || isnan(omega)) {               /* :735 */
            st->breakdown = isnan(omega) ? 2 : 1; ++it; break;
        }
        for (int i = 0; i < n; ++i) x0[i] = x[i];               /* :747 */
        rho = rho_;
    }
    st->iterations = it;
    st->half_steps = nh;
    st->nrm_r = norm;
    free(w);
    return ret;
}

/* ---------------------------------------------------------------------------------
 * ILU0 right-preconditioned BiCGSTAB — restates gpu_pbicgstab pbicgstab.cu:45-154,
 * cuBLAS semantics: axpy = one FMA per element, scal = one multiply.
 * ------------------------------------------------------------------------------- */
/* The loop with the preconditioner built from its own CSR (pia, pja, pa): the matrix itself for the reference's algorithm,
 * its block-diagonal part for the block-Jacobi ILU(0) of row-sharded handles (orc_bicgstab_ilu0_blocks). */
ORC_CLONES
static int ilu0_loop(int n, const int *ia, const int *ja, const double *a,
                     const int *pia, const int *pja, const double *pa,
                     const double *b, int maxit, double tol, double *x,
                     orc_stats *st, double *hist, int hist_cap) {
    size_t N = (size_t)(n > 0 ? n : 1);
    int64_t nnz = pia[n] - pia[0];
    double *M = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    memset(st, 0, sizeof *st);
    int fs = orc_ilu0(n, pia, pja, pa, M);
    if (fs > 0) { free(M); st->breakdown = 4; return 0; }
    double *w = (double *)calloc(N * 7, sizeof(double));
    double *r = w, *rw = w + N, *p = w + 2 * N, *pw = w + 3 * N, *s = w + 4 * N,
           *t = w + 5 * N, *v = w + 6 * N;
    double rho = 0.0, rhop, beta, alpha = 0.0, omega = 0.0, nrmr, nrmr0;
    int nh = 0;
    for (int i = 0; i < n; ++i) x[i] = 1.0;                     /* :306-308 */
    orc_spmv(n, ia, ja, a, x, NULL, r);                         /* :67 */
    for (int i = 0; i < n; ++i) {                               /* :69-73 */
        double q = -1.0 * r[i];
        r[i] = FMA(1.0, b[i], q);
        rw[i] = r[i];
        p[i] = r[i];
    }
    nrmr0 = sqrt(orc_dot(n, r, r));                             /* :74 */
    nrmr = nrmr0;
    st->nrm_r0 = nrmr0;
    if (hist && nh < hist_cap) hist[nh++] = nrmr0;
    st->breakdown = 3;
    int i = 0;
    for (i = 0; i < maxit;) {
        rhop = rho;
        rho = orc_dot(n, rw, r);                                /* :81 */
        if (i > 0) {
            beta = (rho / rhop) * (alpha / omega);              /* :84 */
            double nomega = -omega;
            for (int k = 0; k < n; ++k) {                       /* :86-88 */
                double q = FMA(nomega, v[k], p[k]);
                q = beta * q;
                p[k] = FMA(1.0, r[k], q);
            }
        }
        orc_sptrsv_lower_unit(n, pia, pja, M, p, t);            /* :92-94 */
        orc_sptrsv_upper(n, pia, pja, M, t, pw);                /* :96-98 */
        orc_spmv(n, ia, ja, a, pw, NULL, v);                    /* :104 */
        double temp = orc_dot(n, rw, v);                        /* :106 */
        alpha = rho / temp;
        double nalpha = -alpha;
        for (int k = 0; k < n; ++k) {                           /* :109-110 */
            r[k] = FMA(nalpha, v[k], r[k]);
            x[k] = FMA(alpha, pw[k], x[k]);
        }
        nrmr = sqrt(orc_dot(n, r, r));                          /* :111 */
        if (hist && nh < hist_cap) hist[nh++] = nrmr;
        if (nrmr < tol * nrmr0) { st->converged = 1; st->breakdown = 0; break; }   /* :116 */
        orc_sptrsv_lower_unit(n, pia, pja, M, r, t);            /* :121-123 */
        orc_sptrsv_upper(n, pia, pja, M, t, s);                 /* :125-127 */
        orc_spmv(n, ia, ja, a, s, NULL, t);                     /* :132 */
        temp = orc_dot(n, t, r);                                /* :135 */
        double temp2 = orc_dot(n, t, t);                        /* :136 */
        omega = temp / temp2;
        double nomega = -omega;
        for (int k = 0; k < n; ++k) {                           /* :139-140 */
            x[k] = FMA(omega, s[k], x[k]);
            r[k] = FMA(nomega, t[k], r[k]);
        }
        nrmr = sqrt(orc_dot(n, r, r));                          /* :142 */
        if (hist && nh < hist_cap) hist[nh++] = nrmr;
        if (nrmr < tol * nrmr0) { st->converged = 1; st->breakdown = 0; i++; break; } /* :147 */
        i++;
    }
    st->iterations = i;
    st->half_steps = nh;
    st->nrm_r = nrmr;
    free(w);
    free(M);
    return st->converged;
}

int orc_bicgstab_ilu0(int n, const int *ia, const int *ja, const double *a,
                      const double *b, int maxit, double tol, double *x,
                      orc_stats *st, double *hist, int hist_cap) {
    return ilu0_loop(n, ia, ja, a, ia, ja, a, b, maxit, tol, x, st, hist, hist_cap);
}

/* Block-Jacobi ILU(0): what a row-sharded handle runs (csrc/ilu0.cu build_local_block; the reference has no multi-GPU path).
 * Rows [row_start[k], row_start[k+1]) form block k; the preconditioner is ILU(0) of the matrix with every entry outside the
 * diagonal blocks dropped (= independent ILU(0) factors of the blocks); SpMV, dots and updates act on the whole system.
 * The reduction tree is partition-independent, so the sharded GPU run must give these bits. */
int orc_bicgstab_ilu0_blocks(int n, const int *ia, const int *ja, const double *a,
                             int nblk, const int64_t *row_start,
                             const double *b, int maxit, double tol, double *x,
                             orc_stats *st, double *hist, int hist_cap) {
    const int base = ia[0];
    int64_t nnz = ia[n] - base;
    int *pia = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *pja = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    double *pa = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    int q = 0;
    pia[0] = base;
    for (int k = 0; k < nblk; ++k) {
        const int64_t r0 = row_start[k], r1 = row_start[k + 1];
        for (int64_t i = r0; i < r1; ++i) {
            for (int p = ia[i] - base; p < ia[i + 1] - base; ++p) {
                const int64_t c = (int64_t)ja[p] - base;
                if (c >= r0 && c < r1) { pja[q] = ja[p]; pa[q] = a[p]; ++q; }
            }
            pia[i + 1] = q + base;
        }
    }
    int ret = ilu0_loop(n, ia, ja, a, pia, pja, pa, b, maxit, tol, x, st, hist, hist_cap);
    free(pia); free(pja); free(pa);
    return ret;
}

/* ---------------------------------------------------------------------------------
 * Generators
 * ------------------------------------------------------------------------------- */
int64_t orc_poisson3d(int N, int64_t row0, int64_t row1, int *ia, int *ja, double *a) {
    int64_t nn = (int64_t)N * N, cnt = 0;
    for (int64_t r = row0; r < row1; ++r) {
        int i = (int)(r % N), j = (int)((r / N) % N), k = (int)(r / nn);
        if (ia) ia[r - row0] = (int)cnt;
        int64_t cols[7]; double vals[7]; int m = 0;
        if (k > 0)     { cols[m] = r - nn; vals[m++] = -1.0; }
        if (j > 0)     { cols[m] = r - N;  vals[m++] = -1.0; }
        if (i > 0)     { cols[m] = r - 1;  vals[m++] = -1.0; }
        cols[m] = r; vals[m++] = 6.0;
        if (i < N - 1) { cols[m] = r + 1;  vals[m++] = -1.0; }
        if (j < N - 1) { cols[m] = r + N;  vals[m++] = -1.0; }
        if (k < N - 1) { cols[m] = r + nn; vals[m++] = -1.0; }
        if (ja) for (int q = 0; q < m; ++q) { ja[cnt + q] = (int)cols[q]; a[cnt + q] = vals[q]; }
        cnt += m;
    }
    if (ia) ia[row1 - row0] = (int)cnt;
    return cnt;
}

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t hash2(uint64_t seed, uint64_t a) {
    return mix64(seed + 0x9E3779B97F4A7C15ULL * (a + 1));
}
static inline uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
    return mix64(hash2(seed, a) + 0x9E3779B97F4A7C15ULL * (b + 1));
}
static inline double u01(uint64_t z) { return (double)(z >> 11) * 0x1.0p-53; }

void orc_xtrue(uint64_t seed, int64_t i0, int64_t cnt, double *out) {
    for (int64_t q = 0; q < cnt; ++q) out[q] = 2.0 * u01(hash2(seed, (uint64_t)(i0 + q))) - 1.0;
}

/* Row-length law: class by hash -> 90% Binomial(12,1/2), 9% Binomial(48,1/2),
 * 1% Binomial(192,1/2) candidate off-diagonal columns (means 6/24/96), drawn uniformly
 * from [0,n) \ {i}, duplicates dropped, sorted ascending; values U(-10,10);
 * diagonal = sum|offdiag| (storage order) + U(1,10).  Integer-only sampling so CPU and
 * GPU generate identical bits. */
static int rdd_row(int n, uint64_t seed, int i, int *cols, double *vals) {
    uint64_t hc = hash3(seed, (uint64_t)i, 0);
    double uc = u01(hc);
    int words = uc < 0.90 ? 0 : (uc < 0.99 ? 1 : 3);
    int k;
    if (words == 0) k = __builtin_popcountll(hash3(seed, (uint64_t)i, 1) & 0xFFFULL);
    else if (words == 1) k = __builtin_popcountll(hash3(seed, (uint64_t)i, 1) & 0xFFFFFFFFFFFFULL);
    else k = __builtin_popcountll(hash3(seed, (uint64_t)i, 1)) + __builtin_popcountll(hash3(seed, (uint64_t)i, 2))
           + __builtin_popcountll(hash3(seed, (uint64_t)i, 3));
    if (k > n - 1) k = n - 1;
    int m = 0;
    for (int s = 0; s < k; ++s) {
        uint64_t hz = hash3(seed, (uint64_t)i, 16 + (uint64_t)s);
        int c = (int)(((unsigned __int128)hz * (uint64_t)(n - 1)) >> 64);
        if (c >= i) c += 1;                                   /* skip the diagonal */
        /* sorted insert, drop duplicates */
        int pos = m;
        while (pos > 0 && cols[pos - 1] > c) --pos;
        if (pos > 0 && cols[pos - 1] == c) continue;
        for (int q = m; q > pos; --q) cols[q] = cols[q - 1];
        cols[pos] = c;
        ++m;
    }
    /* insert the diagonal */
    int pos = m;
    while (pos > 0 && cols[pos - 1] > i) --pos;
    for (int q = m; q > pos; --q) cols[q] = cols[q - 1];
    cols[pos] = i;
    ++m;
    if (vals) {
        double sum = 0.0;
        for (int q = 0; q < m; ++q) {
            if (cols[q] == i) continue;
            double u = u01(hash3(seed, (uint64_t)i, 0x100000000ULL + (uint64_t)cols[q]));
            double v = u * 20.0;
            v = v - 10.0;
            vals[q] = v;
            sum = sum + fabs(v);
        }
        double ud = u01(hash3(seed, (uint64_t)i, 4));
        double dv = ud * 9.0;
        dv = dv + 1.0;
        vals[pos] = sum + dv;
    }
    return m;
}

int64_t orc_random_dd(int n, uint64_t seed, int *ia, int *ja, double *a) {
    int cols[200]; double vals[200];
    int64_t cnt = 0;
    for (int i = 0; i < n; ++i) {
        if (ja == NULL) {
            ia[i] = (int)cnt;
            cnt += rdd_row(n, seed, i, cols, NULL);
        } else {
            int m = rdd_row(n, seed, i, cols, vals);
            for (int q = 0; q < m; ++q) { ja[cnt + q] = cols[q]; a[cnt + q] = vals[q]; }
            cnt += m;
        }
    }
    if (ja == NULL) ia[n] = (int)cnt;
    return cnt;
}

void orc_glibc_rand_vector(int n, double p_zero, double vmin, double vmax, double *out) {
    srand(1);   /* glibc default state: the reference never calls srand */
    for (int i = 0; i < n; ++i) {
        double u = (double)rand() / (double)RAND_MAX;           /* pbicgstab.cu:413-417 */
        if (u <= p_zero) { out[i] = 0.0; continue; }
        double w = (double)rand() / (double)RAND_MAX;           /* pbicgstab.cu:419-423 */
        out[i] = w * (vmax - vmin) + vmin;
    }
}

void orc_to_dense_vector(int n, const double *A, const int *IA, double *out) {
    int sum = IA[0], count = 0;
    for (int i = 0; i < n; ++i) {
        if (IA[i + 1] - sum > 0) { out[i] = A[count++]; sum = IA[i + 1]; }
        else out[i] = 0.0;
    }
}
