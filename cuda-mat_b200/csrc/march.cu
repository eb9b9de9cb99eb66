// march.cu — SpMV variant MARCH: plane-marching stencil SpMV with the BiCGSTAB vector updates folded in.
//
// Replaces, per iteration of the unpreconditioned loop (pbicgstab.cu:662-749):
//     cudaMemcpy/Dscal/Daxpy chain for p' (:668-672) + mult_spec + cusparseDcsrmv (:675-676) + cublasDdot (:688)   -> ONE kernel (MAKE_P)
//     cudaMemcpy/Dscal/Daxpy chain for s  (:698-700) + mult_spec + cusparseDcsrmv (:703-704) + 2 cublasDdot (:708-709) -> ONE kernel (MAKE_S)
// and cusparseDcsrmv alone (:67,104,132,646) as MODE LOAD_X.
//
// Applies when the TILED analysis (rowclass.cu) found a superset pattern whose column offsets split into three groups
//     {o : |o + D| <= H}, {o : |o| <= H}, {o : |o - D| <= H}          (7-point N^3 Poisson: D = N^2, H = N)
// with D a multiple of the 2048-row tile and n a multiple of D ("planes").  Then the x window a tile needs at offset +D
// is the tile's own rows one plane up: a persistent CTA that walks a column of tiles plane by plane keeps the last three
// planes' tile (+-H halo) in a shared-memory ring of 4 buffers and fetches only ONE new buffer per tile — 1 + 2H/2048
// x-reads per row instead of the 3.25 of the TILED kernel, which stages all three windows per tile.
//
// Folding the updates: the ring holds the SpMV operand, so it may as well be COMPUTED on the way in.  MAKE_P loads
// r, p, v of the incoming tile (+halo), forms p' = fl(r + fl(beta * fl(p + fl(-omega * v)))) element-wise — the reference's
// scal/axpy chain, every product and sum rounded — stores it into the ring, and writes the centre rows of p' to global
// memory once.  The +-H halo rows are formed redundantly by the neighbouring tile's CTA (same inputs, same operations,
// same bits), which is why p' and v' are written to a SECOND buffer (ping-pong): an in-place update would race with the
// neighbours' halo reads.  MAKE_S does the same for s = fl(r + fl(-alpha * v)).  The fused dots deposit slab sums exactly
// like every other SpMV variant (internal.cuh), so the reduction tree and all results stay bit-identical.
//
// Loads are register-staged and software-pipelined: the global loads of plane k+2 are issued, plane k is multiplied
// out of the ring while they are in flight, then they are converted and stored into the ring slot that plane k-2 left.
// One CTA barrier per tile.
#include "solver.h"
#include <algorithm>
#include <cstring>

namespace cudamat {

enum : int { MARCH_LOAD_X = 0, MARCH_MAKE_P = 1, MARCH_MAKE_S = 2 };

struct MarchArgs {
    int n;
    const double *in0;      // LOAD_X: x           MAKE_P: r       MAKE_S: r
    const double *in1;      //                     MAKE_P: p (old) MAKE_S: v
    const double *in2;      //                     MAKE_P: v (old)
    double *xout;           // MAKE_P: p' / MAKE_S: s  (centre rows, written once)
    double *y;              // y = A * operand
    const double *u;        // dot operand of red0 = y.u; nullptr = the operand itself (taken from the ring)
    const double *d;        // optional diagonal shift
    const unsigned char *tmask;
    RedCtx rc; DevScalars *sc; int check_status;
};

template <int EPT>
struct MarchStage { double a[EPT], b[EPT], c[EPT]; };

// Work split: linear order q = column * P + plane over the S * P tiles; CTA b owns q in [T b / G, T (b+1) / G).
template <int MODE, int NDOT, bool HAS_D, int SL, int EPT>
__global__ void __launch_bounds__(kCtaThreads, 2) k_spmv_march(const MarchArgs a, const __grid_constant__ MarchPlan M) {
    extern __shared__ __align__(16) double ring[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BUF = M.buf_elems, H = M.H, S = M.S, P = M.P;
    const long long T = (long long)S * P;
    long long q = T * blockIdx.x / gridDim.x;
    const long long q_end = T * (blockIdx.x + 1) / gridDim.x;
    pdl_sync();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    double c1 = 0.0, c2 = 0.0;                                     // MAKE_P: beta, -omega; MAKE_S: -alpha
    if (MODE == MARCH_MAKE_P) { c1 = a.sc->beta; c2 = -a.sc->omega; }
    if (MODE == MARCH_MAKE_S) { c2 = -a.sc->alpha; }

    // global loads of tile (col, pl) with its halo -> registers
    auto issue = [&](int col, int pl, double (&ra)[EPT], double (&rb)[EPT], double (&rcv)[EPT]) {
        const long long g0 = ((long long)pl * S + col) * kTile - H;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int e = tid + j * kCtaThreads;
            const long long g = g0 + e;
            const bool ok = e < BUF && g >= 0 && g < a.n;
            ra[j] = ok ? __ldg(a.in0 + g) : 0.0;
            if (MODE != MARCH_LOAD_X) rb[j] = ok ? __ldg(a.in1 + g) : 0.0;
            if (MODE == MARCH_MAKE_P) rcv[j] = ok ? __ldg(a.in2 + g) : 0.0;
        }
    };
    // registers -> operand values -> ring slot pl & 3; centre rows of an owned tile also go to global memory
    auto convert = [&](int col, int pl, bool owned, const double (&ra)[EPT], const double (&rb)[EPT], const double (&rcv)[EPT]) {
        const long long g0 = ((long long)pl * S + col) * kTile - H;
        double *slot = ring + (pl & 3) * BUF;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int e = tid + j * kCtaThreads;
            if (e < BUF) {
                double v;
                if (MODE == MARCH_LOAD_X) v = ra[j];
                else if (MODE == MARCH_MAKE_P) {                   // pbicgstab.cu:668-672
                    v = __dmul_rn(c2, rcv[j]);
                    v = __dadd_rn(rb[j], v);
                    v = __dmul_rn(c1, v);
                    v = __dadd_rn(ra[j], v);
                } else {                                           // pbicgstab.cu:698-700
                    v = __dadd_rn(ra[j], __dmul_rn(c2, rb[j]));
                }
                slot[e] = v;
                if (MODE != MARCH_LOAD_X && owned && e >= H && e < H + kTile) a.xout[g0 + e] = v;
            }
        }
    };

    while (q < q_end) {
        const int col = (int)(q / P), k0 = (int)(q % P);
        const int k1 = (int)min((long long)P - 1, k0 + (q_end - q) - 1);
        q += k1 - k0 + 1;
        __syncthreads();                                           // previous segment's last reads of the ring are done
        {   // fill the ring: planes k0-1, k0, k0+1
            double ra[EPT], rb[EPT], rcv[EPT];
#pragma unroll 1
            for (int pl = k0 - 1; pl <= k0 + 1; ++pl) {
                if (pl < 0 || pl >= P) continue;
                issue(col, pl, ra, rb, rcv);
                convert(col, pl, pl >= k0 && pl <= k1, ra, rb, rcv);
            }
        }
#pragma unroll 1
        for (int k = k0; k <= k1; ++k) {
            const int tile = k * S + col;
            const int row_base = tile * kTile;
            // ---- issue: plane k+2 (consumed after the multiply), masks / dot operand / shift of plane k ----
            double ra[EPT], rb[EPT], rcv[EPT];
            const bool pre = k + 2 <= min(k1 + 1, P - 1);
            if (pre) issue(col, k + 2, ra, rb, rcv);
            unsigned mk[kSlabsPerWarp];
            double uu[kSlabsPerWarp], dd[kSlabsPerWarp];
#pragma unroll
            for (int j = 0; j < kSlabsPerWarp; ++j) {
                const int row = row_base + (j * kCtaWarps + warp) * kSlab + lane;
                mk[j] = __ldg(a.tmask + row);
                if (NDOT >= 1 && a.u) uu[j] = __ldg(a.u + row);
                if (HAS_D) dd[j] = __ldg(a.d + row);
            }
            __syncthreads();                                       // ring stores of the previous step are visible
            // ---- multiply plane k out of the ring ----
            int eoff[SL];
#pragma unroll
            for (int t = 0; t < SL; ++t) eoff[t] = ((k + M.dz[t]) & 3) * BUF + H + M.loff[t];
            const double *cen = ring + (k & 3) * BUF + H;
            double pp[NDOT > 0 ? NDOT : 1], w[NDOT > 0 ? NDOT * 2 : 1];
#pragma unroll
            for (int j = 0; j < kSlabsPerWarp; ++j) {
                const int i = (j * kCtaWarps + warp) * kSlab + lane;
                const unsigned mask = mk[j];
                double xv[SL];
#pragma unroll
                for (int t = 0; t < SL; ++t) xv[t] = ring[eoff[t] + i];
                double sum = 0.0;
                if (__all_sync(0xffffffffu, mask == (1u << SL) - 1u)) {
#pragma unroll
                    for (int t = 0; t < SL; ++t) sum = __fma_rn(M.val[t], xv[t], sum);
                } else {
#pragma unroll
                    for (int t = 0; t < SL; ++t)
                        if (mask & (1u << t)) sum = __fma_rn(M.val[t], xv[t], sum);
                }
                const double xc = (HAS_D || (NDOT >= 1)) ? cen[i] : 0.0;
                if (HAS_D) sum = __dadd_rn(sum, __dmul_rn(dd[j], xc));
                a.y[row_base + i] = sum;
                double p0 = 0.0, p1 = 0.0;
                if (NDOT >= 1) p0 = __dmul_rn(sum, a.u ? uu[j] : xc);
                if (NDOT >= 2) p1 = __dmul_rn(sum, sum);
                if constexpr (NDOT >= 1) {
                    if (j & 1) {
                        w[j >> 1] = packed_pair(pp[0], p0, 16, lane);
                        if constexpr (NDOT >= 2) w[2 + (j >> 1)] = packed_pair(pp[1], p1, 16, lane);
                    } else {
                        pp[0] = p0;
                        if constexpr (NDOT >= 2) pp[1] = p1;
                    }
                }
            }
            if constexpr (NDOT >= 1) {                             // same packed butterfly as k_spmv_tiled (bit-identical slab sums)
                double z = packed_pair(w[0], w[1], 8, lane);
                if constexpr (NDOT >= 2) z = packed_pair(z, packed_pair(w[2], w[3], 8, lane), 4, lane);
                else z = __dadd_rn(z, __shfl_xor_sync(0xffffffffu, z, 4));
                z = __dadd_rn(z, __shfl_xor_sync(0xffffffffu, z, 2));
                z = __dadd_rn(z, __shfl_xor_sync(0xffffffffu, z, 1));
                const int idx = ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1) + (NDOT >= 2 ? 4 * ((lane >> 2) & 1) : 0);
                const int j = idx & 3, qq = idx >> 2;
                if ((lane & (NDOT >= 2 ? 3 : 7)) == 0)
                    __stcg(a.rc.slab_part + (size_t)qq * a.rc.slab_stride + (size_t)tile * kTileSlabs + j * kCtaWarps + warp, z);
            }
            // ---- convert + store plane k+2 into the slot plane k-2 left ----
            if (pre) convert(col, k + 2, k + 2 <= k1, ra, rb, rcv);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// MARCH plan from the TILED plan's superset pattern (pure host code)
bool march_plan_host(const TiledDict &T, long long n, MarchPlan &M) {
    memset(&M, 0, sizeof M);
    if (T.sup_len <= 0 || T.sup_len > 8 || n <= 0) return false;
    constexpr int kHmax = 512;
    long long D = 0;
    for (int q = 0; q < T.sup_len; ++q) {
        const long long o = T.sup_off[q] < 0 ? -(long long)T.sup_off[q] : T.sup_off[q];
        if (o > kHmax && (D == 0 || o < D)) D = o;
    }
    if (D == 0) return false;                                      // a 1-plane (2-D) pattern: nothing to march along
    // D is the smallest far offset; the plane stride is the nearest multiple of the tile such that every far offset is
    // within H of +-D
    D = (D + kHmax) / kTile * kTile;
    if (D < kTile || n % D != 0 || n / D < 1 || n / kTile > 0x3fffffffLL) return false;
    int H = 0;
    for (int q = 0; q < 8; ++q) {
        if (q >= T.sup_len) { M.dz[q] = 0; M.loff[q] = 0; M.val[q] = 0.0; continue; }
        const long long o = T.sup_off[q];
        int dz = 0;
        if (o > kHmax) dz = 1; else if (o < -kHmax) dz = -1;
        const long long lo = o - dz * D;
        if (lo > kHmax || lo < -kHmax) return false;
        M.dz[q] = dz; M.loff[q] = (int)lo; M.val[q] = T.sup_val[q];
        H = std::max(H, (int)(lo < 0 ? -lo : lo));
    }
    H = (H + 31) / 32 * 32;
    M.D = (int)D; M.H = H; M.S = (int)(D / kTile); M.P = (int)(n / D);
    M.buf_elems = kTile + 2 * H; M.len = T.sup_len;
    return true;
}

template <int MODE, int NDOT, bool HAS_D, int SL, int EPT>
static int launch_march_t(cudamat_solver *s, const MarchArgs &a) {
    const MarchPlan &M = *s->march;
    const void *kern = (const void *)k_spmv_march<MODE, NDOT, HAS_D, SL, EPT>;
    const size_t smem = sizeof(double) * 4 * (size_t)M.buf_elems;
    static bool attr_set[64] = {};
    const int dv = s->device & 63;
    if (!attr_set[dv]) { CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024)); attr_set[dv] = true; }
    const long long T = (long long)M.S * M.P;
    const int grid = (int)std::min<long long>(T, (long long)s->march_grid);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kCtaThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    void *args[] = {(void *)&a, (void *)s->march};
    CM_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
    s->launches++;
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}
template <int MODE, int NDOT, bool HAS_D>
static int launch_march_m(cudamat_solver *s, const MarchArgs &a) {
    const MarchPlan &M = *s->march;
    const bool seven = M.len == 7;
    if (M.H <= 256) return seven ? launch_march_t<MODE, NDOT, HAS_D, 7, 5>(s, a) : launch_march_t<MODE, NDOT, HAS_D, 8, 5>(s, a);
    return seven ? launch_march_t<MODE, NDOT, HAS_D, 7, 6>(s, a) : launch_march_t<MODE, NDOT, HAS_D, 8, 6>(s, a);
}

bool march_available(const cudamat_solver *s) { return s->march != nullptr && s->comm == nullptr; }

// y = A x (+ d.x) with ndot fused dots against u (nullptr: against x itself)
int launch_march_spmv(cudamat_solver *s, const SpmvArgs &sa) {
    MarchArgs a{};
    a.n = sa.n; a.in0 = sa.x; a.y = sa.y; a.u = (sa.u == sa.x) ? nullptr : sa.u; a.d = sa.d;
    a.tmask = s->cls[1].d_tmask; a.rc = sa.rc; a.sc = sa.sc; a.check_status = sa.check_status;
    const bool hd = sa.d != nullptr;
    switch (sa.ndot) {
    case 0: return hd ? launch_march_m<MARCH_LOAD_X, 0, true>(s, a) : launch_march_m<MARCH_LOAD_X, 0, false>(s, a);
    case 1: return hd ? launch_march_m<MARCH_LOAD_X, 1, true>(s, a) : launch_march_m<MARCH_LOAD_X, 1, false>(s, a);
    default: return hd ? launch_march_m<MARCH_LOAD_X, 2, true>(s, a) : launch_march_m<MARCH_LOAD_X, 2, false>(s, a);
    }
}
// p' = r + beta (p - omega v) [unprec form]; v' = (A + diag d) p'; red0 = rhat . v'        (pbicgstab.cu:668-689)
int launch_march_make_p(cudamat_solver *s, const double *r, const double *p_old, const double *v_old, double *p_new, double *v_new,
                        const double *rhat, const double *d, const RedCtx &rc) {
    MarchArgs a{};
    a.n = s->n; a.in0 = r; a.in1 = p_old; a.in2 = v_old; a.xout = p_new; a.y = v_new; a.u = rhat; a.d = d;
    a.tmask = s->cls[1].d_tmask; a.rc = rc; a.sc = s->d_sc; a.check_status = 1;
    return d ? launch_march_m<MARCH_MAKE_P, 1, true>(s, a) : launch_march_m<MARCH_MAKE_P, 1, false>(s, a);
}
// s = r - alpha v; t = (A + diag d) s; red0 = t . s, red1 = t . t                            (pbicgstab.cu:698-709)
int launch_march_make_s(cudamat_solver *s, const double *r, const double *v, double *sv, double *t, const double *d, const RedCtx &rc) {
    MarchArgs a{};
    a.n = s->n; a.in0 = r; a.in1 = v; a.xout = sv; a.y = t; a.u = nullptr; a.d = d;
    a.tmask = s->cls[1].d_tmask; a.rc = rc; a.sc = s->d_sc; a.check_status = 1;
    return d ? launch_march_m<MARCH_MAKE_S, 2, true>(s, a) : launch_march_m<MARCH_MAKE_S, 2, false>(s, a);
}

}  // namespace cudamat
