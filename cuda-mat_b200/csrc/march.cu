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
    HaloWait hw;            // sharded handles (LOAD_X only): flags of the neighbours' boundary planes
};

// Work split: items (z-chunk, column) in z-chunk-major order, item = blockIdx.x (+ k * gridDim.x).  Neighbouring CTAs walk
// neighbouring columns through the SAME planes at about the same time, so the +-H halo lines a tile shares with the
// tiles next to it are still in L2 when the second reader arrives.
//
// Thread mapping.  Staging: a thread moves PAIRS of consecutive elements (16-byte loads / stores); pair tid + 512 j.
// Multiply: lane l of warp w owns rows 64 g + 2 l and 64 g + 2 l + 1 of the row groups g = w and g = w + 16: an aligned
// pair of x feeds two rows with one 16-byte shared-memory load, and every row's FMA chain still runs in storage order.
// Slab sums: rows 32 s .. 32 s + 31 sit in one half-warp, two rows per lane; the spec's butterfly (partner row ^ 16, 8,
// 4, 2, 1 — internal.cuh) becomes lane ^ 8, 4, 2, 1 on each of the two row parities followed by the in-thread sum of the
// two: the same tree, `own + partner` at every node, hence the same bits.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds64(uint32_t ad) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(ad)); return v; }
__device__ __forceinline__ double2 lds128(uint32_t ad) {
    double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ad)); return v;
}
// 16-byte asynchronous global -> shared copy (LDGSTS): a register-free prefetch into a slot only this thread reads
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t ad, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ad), "d"(v.x), "d"(v.y) : "memory"); }

// SHAPE 1: the 7-entry pattern (-D, -a, -1, 0, +1, +a, +D) with a even — every 7-point grid stencil with an even line
//          length: planes and the 16-byte alignment of every tap are compile-time facts (5 aligned pair loads, 2 x 2
//          single loads per two rows).
// SHAPE 0: any other pattern of <= 8 entries: planes / offsets are kernel parameters, every tap is two 8-byte loads.
template <int MODE, int NDOT, bool HAS_D, bool U_RING, int SHAPE, int HB, bool SHARD_S = false>
__global__ void __launch_bounds__(kCtaThreads, 2) k_spmv_march(const MarchArgs a, const __grid_constant__ MarchPlan M) {
    constexpr int SL = SHAPE == 1 ? 7 : 8;
    constexpr int H = HB * 256;
    constexpr int BUF = kTile + 2 * H;                             // elements per ring slot
    constexpr int PAIRS = BUF / 2;
    constexpr int PPT = (PAIRS + kCtaThreads - 1) / kCtaThreads;   // pairs per thread: 3 (H = 256: the last one half populated)
    constexpr int NV = MODE == MARCH_MAKE_P ? 3 : MODE == MARCH_MAKE_S ? 2 : 1;
    constexpr int DEPTH = 1;                                       // planes in flight ahead of the multiply (2 measured slower: spills)
    extern __shared__ __align__(16) double ring[];                 // 4 slots
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = M.S, P = M.P, Zc = M.Zc, n = a.n;
    // sharded handles: LOAD_X reads the operand incl. its halo region; MAKE_S forms s = r - alpha v for the shard's own planes and takes
    // the neighbours' planes of s (pushed by k_update_s_boundary right before) from the halo region of the s vector (a.xout).
    // That MAKE_S form is its own instantiation (SHARD_S): the run-time plane test costs the single-GPU kernel its registers
    // (156 bytes of spills, 107 -> 142-154 us at 256^3).
    constexpr bool HALO_S = MODE == MARCH_MAKE_S && SHARD_S;
    constexpr bool HALO_OK = MODE == MARCH_LOAD_X || HALO_S;
    const int nlim = MODE == MARCH_LOAD_X ? M.n_tot : n;
    const int lo_base = HALO_OK ? M.lo_base : -1, hi_base = HALO_OK ? M.hi_base : -1;
    pdl_sync();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    double c1 = 0.0, c2 = 0.0;                                     // MAKE_P: beta, -omega; MAKE_S: -alpha
    if (MODE == MARCH_MAKE_P) { c1 = a.sc->beta; c2 = -a.sc->omega; }
    if (MODE == MARCH_MAKE_S) { c2 = -a.sc->alpha; }
    const double *const in[3] = {a.in0, a.in1, a.in2};
    const uint32_t ring_s = smem_addr(ring);
    const uint32_t side_s = ring_s + 4u * BUF * 8u;                // [u: 2 x 512 x 16 B][d: 2 x 512 x 16 B] thread-private slots
    // byte offset of this thread's first row (of its first row group) inside a ring slot, as seen through every tap
    uint32_t toff[SL];
#pragma unroll
    for (int t = 0; t < SL; ++t) toff[t] = (uint32_t)(H + M.loff[t] + warp * 64 + 2 * lane) * 8u;
    const uint32_t coff = (uint32_t)(H + warp * 64 + 2 * lane) * 8u;

    // global loads of the tile whose buffer starts at element g0 = tile base - H -> registers
    auto issue = [&](int g0, double2 (&st)[NV][PPT], bool halo = false) {
        if (HALO_S && halo) {                                      // a neighbour's plane of s: one vector, taken as it is
            const int lim = M.n_tot;
            if (g0 >= 0 && g0 + BUF <= lim) {
                const double2 *src = reinterpret_cast<const double2 *>(a.xout + g0) + tid;
#pragma unroll
                for (int j = 0; j < PPT; ++j)
                    if ((j + 1) * kCtaThreads <= PAIRS || tid + j * kCtaThreads < PAIRS) st[0][j] = __ldcg(src + j * kCtaThreads);
            } else {
#pragma unroll
                for (int j = 0; j < PPT; ++j) {
                    const int g = g0 + 2 * (tid + j * kCtaThreads);
                    const bool inb = tid + j * kCtaThreads < PAIRS;
                    st[0][j].x = (inb && g >= 0 && g < lim) ? __ldcg(a.xout + g) : 0.0;
                    st[0][j].y = (inb && g + 1 >= 0 && g + 1 < lim) ? __ldcg(a.xout + g + 1) : 0.0;
                }
            }
            return;
        }
        if (g0 >= 0 && g0 + BUF <= nlim) {                         // CTA-uniform: everything but the two ends of the vector
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const double2 *src = reinterpret_cast<const double2 *>(in[v] + g0) + tid;
#pragma unroll
                for (int j = 0; j < PPT; ++j)
                    if ((j + 1) * kCtaThreads <= PAIRS || tid + j * kCtaThreads < PAIRS) st[v][j] = __ldg(src + j * kCtaThreads);
            }
        } else {
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int j = 0; j < PPT; ++j) {
                    const int g = g0 + 2 * (tid + j * kCtaThreads);
                    const bool inb = tid + j * kCtaThreads < PAIRS;
                    st[v][j].x = (inb && g >= 0 && g < nlim) ? __ldg(in[v] + g) : 0.0;
                    st[v][j].y = (inb && g + 1 >= 0 && g + 1 < nlim) ? __ldg(in[v] + g + 1) : 0.0;
                }
        }
    };
    auto make = [&](double r, double p, double v) -> double {
        if (MODE == MARCH_LOAD_X) return r;
        if (MODE == MARCH_MAKE_P) {                                // pbicgstab.cu:668-672: scal, axpy, scal, axpy
            double q = __dmul_rn(c2, v);
            q = __dadd_rn(p, q);
            q = __dmul_rn(c1, q);
            return __dadd_rn(r, q);
        }
        return __dadd_rn(r, __dmul_rn(c2, p));                     // pbicgstab.cu:698-700 (p = v here)
    };
    // registers -> operand values -> ring slot; centre pairs of an owned tile also go to global memory (written once)
    auto convert = [&](int g0, int slot, bool owned, const double2 (&st)[NV][PPT], bool halo = false) {
        const uint32_t dst = ring_s + (uint32_t)(slot * BUF + 2 * tid) * 8u;
        double2 *gout = reinterpret_cast<double2 *>(a.xout + g0) + tid;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int pi = tid + j * kCtaThreads;
            if ((j + 1) * kCtaThreads <= PAIRS || pi < PAIRS) {
                double2 o;
                if (HALO_S && halo) o = st[0][j];
                else {
                    o.x = make(st[0][j].x, st[NV > 1 ? 1 : 0][j].x, st[NV > 2 ? 2 : 0][j].x);
                    o.y = make(st[0][j].y, st[NV > 1 ? 1 : 0][j].y, st[NV > 2 ? 2 : 0][j].y);
                }
                sts128(dst + j * kCtaThreads * 16, o);
                if (MODE != MARCH_LOAD_X && owned && 2 * pi >= H && 2 * pi < H + kTile) gout[j * kCtaThreads] = o;
            }
        }
    };

    const int items = Zc * S;
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int zc = item / S, col = item - zc * S;
        const int k0 = (int)((long long)P * zc / Zc), k1 = (int)((long long)P * (zc + 1) / Zc) - 1;
        if (k1 < k0) continue;
        __syncthreads();                                           // the previous item's last reads of the ring are done
        // first element of plane pl's buffer: planes -1 / P of a shard are the neighbours' planes in the halo region
        auto gofs = [&](int pl) -> int {
            return (pl < 0 ? lo_base : pl >= P ? hi_base : pl * S * kTile) + col * kTile - H;
        };
        if (HALO_OK && a.hw.nsrc > 0) {                            // the neighbours' rows of this exchange have arrived?
            if (k0 == 0 && lo_base >= 0) halo_wait(a.hw, col, a.sc ? &a.sc->status : nullptr);
            if (k1 == P - 1 && hi_base >= 0) halo_wait(a.hw, (P - 1) * S + col, a.sc ? &a.sc->status : nullptr);
        }
        {   // fill the ring: planes k0-1, k0, k0+1
            double2 st[NV][PPT];
#pragma unroll 1
            for (int pl = k0 - 1; pl <= k0 + 1; ++pl) {
                if ((pl < 0 && lo_base < 0) || (pl >= P && hi_base < 0) || pl > P) continue;
                const int g0 = gofs(pl);
                const bool halo = pl < 0 || pl >= P;
                issue(g0, st, halo);
                convert(g0, pl & 3, pl >= k0 && pl <= k1, st, halo);
            }
        }
        // Per-row side inputs of a plane — presence masks, dot operand u, shift d — are fetched one step AHEAD, right after
        // the previous plane's products are formed: the masks into two registers, u and d with 16-byte asynchronous copies
        // (cp.async) into shared-memory slots that only the issuing thread reads back, so they cost no registers while
        // in flight and need no barrier.
        unsigned mk[2];
        auto fetch_side = [&](int k) {
            const int row0 = (k * S + col) * kTile + warp * 64 + 2 * lane;
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                const int r0 = row0 + gi * kCtaWarps * 64;
                mk[gi] = __ldg(reinterpret_cast<const unsigned short *>(a.tmask + r0));
                if (NDOT >= 1 && !U_RING) cp_async16(side_s + (uint32_t)(gi * kCtaThreads + tid) * 16u, a.u + r0);
                if (HAS_D) cp_async16(side_s + (uint32_t)((2 + gi) * kCtaThreads + tid) * 16u, a.d + r0);
            }
            if ((NDOT >= 1 && !U_RING) || HAS_D) cp_async_commit();
        };
        fetch_side(k0);
        // Software pipeline: DEPTH planes ahead of the multiply are in flight in registers (LOAD_X: 2 x 1 vector, MAKE_*: 1 x
        // 2-3 vectors).  One step = barrier, multiply plane k out of the ring, convert + store plane k+2 into the slot plane
        // k-2 left, issue the loads of plane k+2+DEPTH into the registers that just became free.
        const int lim = hi_base >= 0 ? k1 + 1 : min(k1 + 1, P - 1);
        auto step = [&](int k, double2 (&st)[NV][PPT]) {
            const int tile = k * S + col;
            const int row0 = tile * kTile + warp * 64 + 2 * lane;  // this thread's first row (first row group)
            __syncthreads();                                       // ring stores of the previous step are visible
            // ---- multiply plane k out of the ring ----
            uint32_t sb[3];                                        // slots of planes k-1, k, k+1 (shared-memory byte addresses)
#pragma unroll
            for (int d = 0; d < 3; ++d) sb[d] = ring_s + (uint32_t)(((k + d - 1) & 3) * BUF) * 8u;
            constexpr unsigned FULL2 = ((1u << SL) - 1u) * 0x101u;
            constexpr int GB = kCtaWarps * 64 * 8;                 // byte distance of the warp's second row group
            double2 *yp = reinterpret_cast<double2 *>(a.y + row0);
            if ((NDOT >= 1 && !U_RING) || HAS_D) cp_async_wait_all();   // this thread's u / d of plane k (issued a step ago)
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                double s0 = 0.0, s1 = 0.0;
                double x0[SL], x1[SL];                             // the two rows' operand through every tap
                double2 xc = make_double2(0.0, 0.0);               // the operand at the rows themselves
                if (SHAPE == 1) {
                    // taps (-D, -a, -1, 0, +1, +a, +D): five aligned pairs; the +-1 taps come from the neighbouring lanes
                    // (the two edge lanes of the warp read theirs)
#pragma unroll
                    for (int t = 0; t < SL; ++t) {
                        if (t == 2 || t == 4) continue;
                        const double2 xx = lds128(sb[t == 0 ? 0 : t == 6 ? 2 : 1] + toff[t] + gi * GB);
                        x0[t] = xx.x; x1[t] = xx.y;
                    }
                    xc = make_double2(x0[3], x1[3]);
                    double lft = __shfl_up_sync(0xffffffffu, xc.y, 1), rgt = __shfl_down_sync(0xffffffffu, xc.x, 1);
                    if (lane == 0) lft = lds64(sb[1] + coff + gi * GB - 8);
                    if (lane == 31) rgt = lds64(sb[1] + coff + gi * GB + 16);
                    x0[2] = lft; x1[2] = xc.x;
                    x0[4] = xc.y; x1[4] = rgt;
                } else {
#pragma unroll
                    for (int t = 0; t < SL; ++t) {
                        const uint32_t ad = sb[M.dz[t] + 1] + toff[t] + gi * GB;
                        x0[t] = lds64(ad); x1[t] = lds64(ad + 8);
                    }
                    if (HAS_D || (NDOT >= 1 && U_RING)) xc = lds128(sb[1] + coff + gi * GB);
                }
                if (__all_sync(0xffffffffu, mk[gi] == FULL2)) {    // interior rows: the whole pattern, no predicates
#pragma unroll
                    for (int t = 0; t < SL; ++t) { s0 = __fma_rn(M.val[t], x0[t], s0); s1 = __fma_rn(M.val[t], x1[t], s1); }
                } else {
                    const unsigned m0 = mk[gi] & 0xffu, m1 = mk[gi] >> 8;
#pragma unroll
                    for (int t = 0; t < SL; ++t) {
                        if (m0 & (1u << t)) s0 = __fma_rn(M.val[t], x0[t], s0);
                        if (m1 & (1u << t)) s1 = __fma_rn(M.val[t], x1[t], s1);
                    }
                }
                if (HAS_D) {
                    const double2 dv = lds128(side_s + (uint32_t)((2 + gi) * kCtaThreads + tid) * 16u);
                    s0 = __dadd_rn(s0, __dmul_rn(dv.x, xc.x)); s1 = __dadd_rn(s1, __dmul_rn(dv.y, xc.y));
                }
                yp[gi * kCtaWarps * 32] = make_double2(s0, s1);
                if constexpr (NDOT >= 1) {
                    // slab sums of y.u [and y.y] of this row group's two slabs.  Spec tree of a slab (internal.cuh): partner row
                    // ^ 16, 8, 4, 2, 1 = lane ^ 8, 4, 2, 1 on each row parity, then parity 0 + parity 1.  With two dots the
                    // first level is PACKED (packed_pair): lanes with bit 3 clear carry y.u, the others y.y.
                    const double2 uv = U_RING ? xc : lds128(side_s + (uint32_t)(gi * kCtaThreads + tid) * 16u);
                    double e0 = __dmul_rn(s0, uv.x), e1 = __dmul_rn(s1, uv.y);
                    if constexpr (NDOT >= 2) {
                        e0 = packed_pair(e0, __dmul_rn(s0, s0), 8, lane);
                        e1 = packed_pair(e1, __dmul_rn(s1, s1), 8, lane);
                    } else {
                        e0 = __dadd_rn(e0, __shfl_xor_sync(0xffffffffu, e0, 8));
                        e1 = __dadd_rn(e1, __shfl_xor_sync(0xffffffffu, e1, 8));
                    }
#pragma unroll
                    for (int sh = 4; sh >= 1; sh >>= 1) {
                        e0 = __dadd_rn(e0, __shfl_xor_sync(0xffffffffu, e0, sh));
                        e1 = __dadd_rn(e1, __shfl_xor_sync(0xffffffffu, e1, sh));
                    }
                    const double z = __dadd_rn(e0, e1);
                    const int slab = tile * kTileSlabs + ((gi * kCtaWarps + warp) << 1) + (lane >> 4);
                    if ((lane & 7) == 0 && (NDOT >= 2 || (lane & 8) == 0))
                        __stcg(a.rc.slab_part + (size_t)(NDOT >= 2 ? ((lane >> 3) & 1) : 0) * a.rc.slab_stride + slab, z);
                }
                if (gi == 1 && k < k1) fetch_side(k + 1);
            }
            if (k + 2 <= lim) convert(((k + 2) * S + col) * kTile - H, (k + 2) & 3, k + 2 <= k1, st, k + 2 >= P);
            if (k + 2 + DEPTH <= lim) issue(gofs(k + 2 + DEPTH), st, k + 2 + DEPTH >= P);
        };
        double2 stA[NV][PPT], stB[DEPTH > 1 ? NV : 1][DEPTH > 1 ? PPT : 1];
        (void)stB;
        if (k0 + 2 <= lim) issue(gofs(k0 + 2), stA, k0 + 2 >= P);
        if constexpr (DEPTH > 1) { if (k0 + 3 <= lim) issue(gofs(k0 + 3), stB, k0 + 3 >= P); }
#pragma unroll 1
        for (int k = k0; k <= k1; k += DEPTH) {
            step(k, stA);
            if constexpr (DEPTH > 1) { if (k + 1 <= k1) step(k + 1, stB); }
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
    H = std::max(256, (H + 255) / 256 * 256);                      // 256 or 512: ring slots of 5 or 6 x 512 elements
    M.D = (int)D; M.H = H; M.S = (int)(D / kTile); M.P = (int)(n / D);
    M.buf_elems = kTile + 2 * H; M.len = T.sup_len;
    M.Zc = 1;
    // SHAPE 1: (-D, -a, -1, 0, +1, +a, +D), a even (the kernel then knows planes and alignment at compile time)
    M.shape = (M.len == 7 && M.dz[0] == -1 && M.dz[6] == 1 && M.loff[0] == 0 && M.loff[6] == 0 && M.dz[1] == 0 && M.dz[2] == 0 &&
               M.dz[3] == 0 && M.dz[4] == 0 && M.dz[5] == 0 && M.loff[2] == -1 && M.loff[3] == 0 && M.loff[4] == 1 &&
               (M.loff[1] & 1) == 0 && (M.loff[5] & 1) == 0) ? 1 : 0;
    return true;
}

// z-chunks of the plane range: as many as the CTA budget allows in ONE round of the G resident CTAs, but runs of at least 8
// planes (each run re-reads 2 extra planes).  256^3: S = 32, Zc = 9 (288 items for 296 CTAs); 512^3: S = 128, Zc = 2 (256
// items, 86 % of the slots).  Several rounds per CTA were tried for 512^3 (Zc = 9: 1152 items = 3.9 rounds, on paper 94 %
// instead of 86 %): 245 instead of 275 it/s — CTAs drift apart and the neighbouring columns' shared halo lines leave the L2.
int march_choose_zc(int S, int P, int G) { return std::max(1, std::min(G / std::max(1, S), std::max(1, P / 8))); }
template <int MODE, int NDOT, bool HAS_D, bool U_RING, int SHAPE, int HB, bool SHARD_S = false>
static int launch_march_t(cudamat_solver *s, const MarchArgs &a) {
    MarchPlan M = *s->march;
    const void *kern = (const void *)k_spmv_march<MODE, NDOT, HAS_D, U_RING, SHAPE, HB, SHARD_S>;
    const size_t smem = sizeof(double) * 4 * (size_t)(kTile + 2 * HB * 256) +
                        ((NDOT >= 1 && !U_RING) || HAS_D ? (HAS_D ? 4 : 2) * kCtaThreads * 16 : 0);
    static bool attr_set[64] = {};
    const int dv = s->device & 63;
    if (!attr_set[dv]) { CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr_set[dv] = true; }
    const int G = std::max(1, s->march_grid);
    M.Zc = march_choose_zc(M.S, M.P, G);
    const int grid = (int)std::min<long long>((long long)M.Zc * M.S, (long long)G);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kCtaThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    // Early (programmatic) launch only when the grid fills (nearly) every CTA slot of the GPU: CTAs that become resident while
    // the predecessor drains are placed wherever a slot frees up first, and a grid that underfills the GPU then ends up two to
    // an SM on some SMs and none on others (measured at 512^3: 256 CTAs for 296 slots, 4.20 ms per iteration with the early
    // launch, 3.81 ms without).  The kernel still releases ITS dependents at its top.
    cfg.attrs = at; cfg.numAttrs = (pdl_enabled() && grid * 20 >= G * 19) ? 1 : 0;
    void *args[] = {(void *)&a, (void *)&M};
    CM_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
    s->launches++;
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}
template <int MODE, int NDOT, bool HAS_D, bool U_RING, bool SHARD_S = false>
static int launch_march_m(cudamat_solver *s, const MarchArgs &a) {
    const MarchPlan &M = *s->march;
    if (M.H <= 256) return M.shape == 1 ? launch_march_t<MODE, NDOT, HAS_D, U_RING, 1, 1, SHARD_S>(s, a) : launch_march_t<MODE, NDOT, HAS_D, U_RING, 0, 1, SHARD_S>(s, a);
    return M.shape == 1 ? launch_march_t<MODE, NDOT, HAS_D, U_RING, 1, 2, SHARD_S>(s, a) : launch_march_t<MODE, NDOT, HAS_D, U_RING, 0, 2, SHARD_S>(s, a);
}
static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool march_available(const cudamat_solver *s) { return s->march != nullptr; }     // sharded handles: LOAD_X only (solver.cu keeps fuse = 0)

// y = A x (+ d.x) with ndot fused dots (red0 = y.u, red1 = y.y); u == x: the operand comes out of the ring.
// Staging and stores move 16-byte pairs: operands that are not 16-byte aligned are left to the TILED kernel (returns 1).
bool march_spmv_usable(const cudamat_solver *s, const SpmvArgs &sa) {
    return march_available(s) && sa.n > 0 && sa.x != sa.y && al16(sa.x) && al16(sa.y) && (sa.ndot == 0 || al16(sa.u)) && al16(sa.d);
}
int launch_march_spmv(cudamat_solver *s, const SpmvArgs &sa) {
    MarchArgs a{};
    const bool ring = sa.ndot >= 1 && sa.u == sa.x;
    a.n = sa.n; a.in0 = sa.x; a.y = sa.y; a.u = ring ? nullptr : sa.u; a.d = sa.d;
    a.tmask = s->march_tmask; a.rc = sa.rc; a.sc = sa.sc; a.check_status = sa.check_status; a.hw = sa.hw;
    const bool hd = sa.d != nullptr;
    if (sa.ndot == 0) return hd ? launch_march_m<MARCH_LOAD_X, 0, true, false>(s, a) : launch_march_m<MARCH_LOAD_X, 0, false, false>(s, a);
    if (sa.ndot == 1) {
        if (ring) return hd ? launch_march_m<MARCH_LOAD_X, 1, true, true>(s, a) : launch_march_m<MARCH_LOAD_X, 1, false, true>(s, a);
        return hd ? launch_march_m<MARCH_LOAD_X, 1, true, false>(s, a) : launch_march_m<MARCH_LOAD_X, 1, false, false>(s, a);
    }
    if (ring) return hd ? launch_march_m<MARCH_LOAD_X, 2, true, true>(s, a) : launch_march_m<MARCH_LOAD_X, 2, false, true>(s, a);
    return hd ? launch_march_m<MARCH_LOAD_X, 2, true, false>(s, a) : launch_march_m<MARCH_LOAD_X, 2, false, false>(s, a);
}
// p' = r + beta (p - omega v) [unprec form]; v' = (A + diag d) p'; red0 = rhat . v'        (pbicgstab.cu:668-689)
int launch_march_make_p(cudamat_solver *s, const double *r, const double *p_old, const double *v_old, double *p_new, double *v_new,
                        const double *rhat, const double *d, const RedCtx &rc) {
    MarchArgs a{};
    a.n = s->n; a.in0 = r; a.in1 = p_old; a.in2 = v_old; a.xout = p_new; a.y = v_new; a.u = rhat; a.d = d;
    a.tmask = s->march_tmask; a.rc = rc; a.sc = s->d_sc; a.check_status = 1;
    int e = ev_mark(s, true);
    if (e) return e;
    if ((e = d ? launch_march_m<MARCH_MAKE_P, 1, true, false>(s, a) : launch_march_m<MARCH_MAKE_P, 1, false, false>(s, a))) return e;
    return ev_mark(s, false);
}
// s = r - alpha v; t = (A + diag d) s; red0 = t . s, red1 = t . t                            (pbicgstab.cu:698-709)
int launch_march_make_s(cudamat_solver *s, const double *r, const double *v, double *sv, double *t, const double *d, const RedCtx &rc,
                        const HaloWait *hw) {
    MarchArgs a{};
    if (hw) a.hw = *hw;
    a.n = s->n; a.in0 = r; a.in1 = v; a.xout = sv; a.y = t; a.u = nullptr; a.d = d;
    a.tmask = s->march_tmask; a.rc = rc; a.sc = s->d_sc; a.check_status = 1;
    int e = ev_mark(s, true);
    if (e) return e;
    // slab shards (planes of s behind a shard face come from the neighbours): the SHARD_S instantiation
    const bool shard = s->march->lo_base >= 0 || s->march->hi_base >= 0;
    if (shard) e = d ? launch_march_m<MARCH_MAKE_S, 2, true, true, true>(s, a) : launch_march_m<MARCH_MAKE_S, 2, false, true, true>(s, a);
    else       e = d ? launch_march_m<MARCH_MAKE_S, 2, true, true>(s, a) : launch_march_m<MARCH_MAKE_S, 2, false, true>(s, a);
    if (e) return e;
    return ev_mark(s, false);
}

}  // namespace cudamat

// host planner of the MARCH variant, exported for the CPU tests (no device needed)
extern "C" int cudamat_march_plan_host(int sup_len, const int *sup_off, const double *sup_val, long long n, int *ok, int *D, int *H,
                                       int *S, int *P, int *dz, int *loff) {
    using namespace cudamat;
    if (sup_len < 0 || sup_len > 8 || !sup_off || !ok) { set_error("march_plan_host: invalid argument"); return CUDAMAT_E_INVALID; }
    TiledDict *T = new TiledDict();
    memset(T, 0, sizeof(TiledDict));
    T->sup_len = sup_len;
    for (int q = 0; q < sup_len; ++q) { T->sup_off[q] = sup_off[q]; T->sup_val[q] = sup_val ? sup_val[q] : 1.0; }
    MarchPlan M;
    *ok = march_plan_host(*T, n, M) ? 1 : 0;
    delete T;
    if (*ok) {
        if (D) *D = M.D; if (H) *H = M.H; if (S) *S = M.S; if (P) *P = M.P;
        for (int q = 0; q < 8; ++q) { if (dz) dz[q] = M.dz[q]; if (loff) loff[q] = M.loff[q]; }
    }
    return CUDAMAT_OK;
}
