// kernels.cu — hand-written sm_100a kernels of the BiCGSTAB iteration:
//   CSR SpMV with fused dot epilogues (replaces cusparseDcsrmv + cublasDdot, pbicgstab.cu:67,104-106,
//   132-136,646,676-688,704-709 and mult_spec :36-42), the fused vector updates with fused
//   dot/nrm2 reductions (replaces the cublasDaxpy/Dscal/Dcopy/Dnrm2 + cudaMemcpy D2D chains
//   :69-74,86-88,109-111,139-142,668-672,694-700,714-723,744-747), and the synthetic generators.
// Every kernel that reduces runs one CTA per 2048-row tile so the reduction tree of the spec
// (internal.cuh) is independent of scheduling; scalars never leave the device.
#include "solver.h"
#include "rowfuncs.cuh"
#include <cub/device/device_scan.cuh>
#include <cstdlib>
#include <algorithm>

namespace cudamat {

// ------------------------------------------------------------------------------------------
// SpMV variant ROWLANE: one row per lane, direct global loads, 8-deep predicated batches.
// ------------------------------------------------------------------------------------------
#ifndef CUDAMAT_ROWLANE_PREFETCH
#define CUDAMAT_ROWLANE_PREFETCH 0
#endif
#ifndef CUDAMAT_ROWLANE_MINB
#define CUDAMAT_ROWLANE_MINB 4
#endif
template <bool HAS_D, int NDOT>
__global__ void __launch_bounds__(kCtaThreads, CUDAMAT_ROWLANE_MINB) k_spmv_rowlane(const SpmvArgs a) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_base = blockIdx.x * kTile;
    // The CSR arrays are constant during a solve, so while the predecessor kernel drains (before the PDL
    // wait) this CTA already pulls the row bounds of its first slabs and the matrix lines behind them towards
    // L2/L1; only the x gathers and the dot operand depend on the predecessor.
    {
        const int row = row_base + warp * kSlab + lane;
        if (row < a.n) {
            const int s0 = __ldg(a.ia + row), e0 = __ldg(a.ia + row + 1);
            if (e0 > s0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.val + s0));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.val + e0 - 1));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ja + s0));
            }
        }
    }
    pdl_sync();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    halo_wait(a.hw, blockIdx.x, a.sc ? &a.sc->status : nullptr);
    // fused-dot products of the previous slab: their butterflies are issued right after the next slab's
    // row-pointer loads, so the shuffle latency hides under that memory round trip
    double pp0 = 0.0, pp1 = 0.0;
    int pslab = -1;
#if CUDAMAT_ROWLANE_PREFETCH
    // row bounds and dot operand of the NEXT slab are requested one slab ahead
    int ns = 0, ne = 0; double nu = 0.0;
    {
        const int row = row_base + warp * kSlab + lane;
        if (row < a.n) { ns = __ldg(a.ia + row); ne = __ldg(a.ia + row + 1); if (NDOT >= 1) nu = __ldg(a.u + row); }
    }
#endif
#pragma unroll 1
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int slab = j * kCtaWarps + warp;
        const int row0 = row_base + slab * kSlab;
        if (row0 >= a.n) break;                                   // warp-uniform
        const int row = row0 + lane;
        const bool active = row < a.n;
        int s = 0, e = 0;
        double uval = 0.0;
#if CUDAMAT_ROWLANE_PREFETCH
        s = ns; e = ne; uval = nu;
        {
            const int nrow = row + kCtaWarps * kSlab;
            ns = 0; ne = 0; nu = 0.0;
            if (j + 1 < kSlabsPerWarp && nrow < a.n) { ns = __ldg(a.ia + nrow); ne = __ldg(a.ia + nrow + 1); if (NDOT >= 1) nu = __ldg(a.u + nrow); }
        }
#else
        // everything that does not depend on the row's entries is requested up front
        if (active) { s = __ldg(a.ia + row); e = __ldg(a.ia + row + 1); if (NDOT >= 1) uval = __ldg(a.u + row); }
#endif
        if (NDOT >= 1 && pslab >= 0) {
            slab_deposit(a.rc, 0, blockIdx.x * kTileSlabs + (pslab), pp0, lane);
            if (NDOT >= 2) slab_deposit(a.rc, 1, blockIdx.x * kTileSlabs + (pslab), pp1, lane);
        }
        const int len = e - s;
        const int shortlen = (len <= kLongRow) ? len : 0;
        const int maxlen = __reduce_max_sync(0xffffffffu, shortlen);
        double sum = 0.0;
        for (int k0 = 0; k0 < maxlen; k0 += 8) {
            int cj[8]; double av[8], xv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const bool p = (k0 + q) < shortlen;
                cj[q] = p ? __ldg(a.ja + s + k0 + q) : -1;
                av[q] = p ? __ldg(a.val + s + k0 + q) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) xv[q] = (cj[q] >= 0) ? __ldg(a.x + cj[q]) : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((k0 + q) < shortlen) sum = __fma_rn(av[q], xv[q], sum);
        }
        unsigned lm = __ballot_sync(0xffffffffu, len > kLongRow);
        while (lm) {
            const int src = __ffs(lm) - 1;
            lm &= lm - 1;
            const int ss = __shfl_sync(0xffffffffu, s, src), ee = __shfl_sync(0xffffffffu, e, src);
            const double acc = rowsum_long(a, ss, ee, lane);
            if (lane == src) sum = acc;
        }
        if (HAS_D) { if (active) sum = __dadd_rn(sum, __dmul_rn(__ldg(a.d + row), __ldg(a.x + row))); }
        if (active) a.y[row] = sum;
        if (NDOT >= 1) pp0 = active ? __dmul_rn(sum, uval) : 0.0;
        if (NDOT >= 2) pp1 = active ? __dmul_rn(sum, sum) : 0.0;
        pslab = slab;
    }
    if (NDOT >= 1) {
        if (pslab >= 0) {
            slab_deposit(a.rc, 0, blockIdx.x * kTileSlabs + (pslab), pp0, lane);
            if (NDOT >= 2) slab_deposit(a.rc, 1, blockIdx.x * kTileSlabs + (pslab), pp1, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------
// SpMV variants PATTERN / CLASS: one row per lane like ROWLANE, but the column indices (and for CLASS the
// values) of a row come from its class in a small dictionary held in shared memory: one byte per row is read
// instead of 4 (12) bytes per entry (rowclass.cu).  col = row + offset, so the x gathers of a warp whose rows
// share a class are perfectly coalesced.  Entry order = storage order: bit-identical to the CSR kernels.
// PATTERN finds the row's first value at ia[first row of the slab] + exclusive scan of the class lengths.
// ------------------------------------------------------------------------------------------
#ifndef CUDAMAT_CLASS_MINB
#define CUDAMAT_CLASS_MINB 4
#endif
template <bool HAS_D, int NDOT, bool CLS_VALS>
__global__ void __launch_bounds__(kCtaThreads, CUDAMAT_CLASS_MINB) k_spmv_class(const SpmvArgs a, const ClassArgs c,
                                                                                const __grid_constant__ DictParam D) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntile = (a.n + kTile - 1) / kTile;
    const int tile_end = min(ntile, (int)(blockIdx.x + 1) * c.tiles_per_cta);
    // The class ids are constant during a solve: those of the first tile are fetched while the predecessor kernel
    // drains.  There is no CTA-level step in this kernel (slab sums go straight to global memory), so a CTA can walk
    // several consecutive tiles: fewer, longer-lived CTAs.
    int cid[kSlabsPerWarp], nid[kSlabsPerWarp];
    auto load_ids = [&](int tile, int (&id)[kSlabsPerWarp]) {
#pragma unroll
        for (int j = 0; j < kSlabsPerWarp; ++j) {
            const int row = tile * kTile + (j * kCtaWarps + warp) * kSlab + lane;
            id[j] = (tile < tile_end && row < a.n) ? (int)__ldg(c.cls + row) : 0xff;
        }
    };
    load_ids(blockIdx.x * c.tiles_per_cta, cid);
    pdl_sync();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    for (int tile = blockIdx.x * c.tiles_per_cta; tile < tile_end; ++tile) {
        const int row_base = tile * kTile;
        halo_wait(a.hw, tile, a.sc ? &a.sc->status : nullptr);
        load_ids(tile + 1, nid);
#pragma unroll
        for (int j = 0; j < kSlabsPerWarp; ++j) {
            const int row0 = row_base + (j * kCtaWarps + warp) * kSlab;
            if (row0 >= a.n) continue;                                // warp-uniform
            const int row = row0 + lane;
            const bool active = row < a.n;
            double uval = 0.0;
            if (NDOT >= 1 && active) uval = __ldg(a.u + row);
            double sum;
            const int c0 = __shfl_sync(0xffffffffu, cid[j], 0);
            const bool uni = __all_sync(0xffffffffu, cid[j] == c0) && c0 != 0xff;
            const int len0 = uni ? D.len[c0] : 0, rp = uni ? D.run[c0] : -1;
            if ((len0 == 7 && rp == 2) || (len0 == 5 && rp == 1)) {
                const int *off = D.off + c0 * kDictLen;
                const double *dv = D.val + c0 * kDictLen;
                const double *vrow = CLS_VALS ? nullptr : a.val + (__ldg(a.ia + row0) + lane * len0);
                if (len0 == 7) sum = class_row_uniform<7, 2, CLS_VALS>(a.x + row, vrow, off, dv, lane);
                else           sum = class_row_uniform<5, 1, CLS_VALS>(a.x + row, vrow, off, dv, lane);
            } else {
                sum = class_row_general<CLS_VALS>(a.x, a.val, a.ia, active ? cid[j] : 0, row0, row, active, lane, D);
            }
            if (HAS_D) { if (active) sum = __dadd_rn(sum, __dmul_rn(__ldg(a.d + row), __ldg(a.x + row))); }
            if (active) a.y[row] = sum;
            const int slab = tile * kTileSlabs + j * kCtaWarps + warp;
            if (NDOT >= 1) slab_deposit(a.rc, 0, slab, active ? __dmul_rn(sum, uval) : 0.0, lane);
            if (NDOT >= 2) slab_deposit(a.rc, 1, slab, active ? __dmul_rn(sum, sum) : 0.0, lane);
        }
#pragma unroll
        for (int j = 0; j < kSlabsPerWarp; ++j) cid[j] = nid[j];
    }
}

// ------------------------------------------------------------------------------------------
// SpMV variant STAGED: each warp streams the contiguous (val, col) span of its 32-row slab into
// shared memory with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), a ring of
// `stages` slabs deep, then every lane walks its own row out of shared memory.  Global loads
// issued by threads are only the row pointers and the coalesced x gathers.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------
// SpMV variant TILED: CLASS + shared-memory staging of x.  The column offsets of the frequent classes cluster into
// a few windows (7-point stencil: {-N^2}, {-N..N}, {+N^2}); for a 2048-row tile each window is ONE contiguous range of
// x, copied into shared memory by a TMA bulk copy (cp.async.bulk + mbarrier complete_tx) — 3 copies per CTA replace
// 7 x 2048 per-thread gathers, every tap becomes a conflict-free shared-memory read at row + disp[class][entry], and
// the global traffic is full lines regardless of the L1 hit rate.  Tiles with a row whose class does not fit the
// windows (halo columns of a shard, irregular rows) take the gather path of the CLASS kernel inside the same launch.
// Entry order = storage order: bit-identical to every other variant.
// ------------------------------------------------------------------------------------------
// one row out of the staged windows: the lane's class record (shared memory) gives the byte offset of every entry
// relative to the row's own position (-1 = no entry) and, for the values dictionary, its value
// `minlen` = shortest class of the matrix: passes that lie below it need no per-entry predicate
template <bool CLS_VALS>
__device__ __forceinline__ double tiled_row(const char *xrow, const TiledSmemClass *rec, const double *vrow, int maxlen, int minlen) {
    double sum = 0.0;
#pragma unroll 1
    for (int h = 0; h < maxlen; h += 4) {                           // 4 entries per pass: 3 (1) 16-byte dictionary reads
        const int4 b = *reinterpret_cast<const int4 *>(rec->boff + h);
        const int bo[4] = {b.x, b.y, b.z, b.w};
        double av[4];
        if (CLS_VALS) {
            const double2 t0 = *reinterpret_cast<const double2 *>(rec->val + h), t1 = *reinterpret_cast<const double2 *>(rec->val + h + 2);
            av[0] = t0.x; av[1] = t0.y; av[2] = t1.x; av[3] = t1.y;
        }
        if (h + 4 <= minlen) {                                      // kernel-uniform
            double xv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                xv[q] = *reinterpret_cast<const double *>(xrow + bo[q]);
                if (!CLS_VALS) av[q] = __ldg(vrow + h + q);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) sum = __fma_rn(av[q], xv[q], sum);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (bo[q] >= 0) {
                    const double xq = *reinterpret_cast<const double *>(xrow + bo[q]);
                    const double aq = CLS_VALS ? av[q] : __ldg(vrow + h + q);
                    sum = __fma_rn(aq, xq, sum);
                }
            }
        }
    }
    return sum;
}
// row = superset pattern + presence mask: byte offsets and values are kernel-wide constants at compile-time positions
// of the parameter block (uniform / constant-bank operands), entries in ascending-offset = storage order.  The x reads
// are unconditional (every pattern position of every row lies inside the staged windows); only the FMA is predicated.
template <bool CLS_VALS, int SL>
__device__ __forceinline__ double tiled_row_sup(const char *xrow, const TiledDict &D, unsigned mask, const double *vrow) {
    double sum = 0.0;
    if constexpr (!CLS_VALS) {                                      // values streamed from CSR: entry by entry (registers)
        int k = 0;
#pragma unroll
        for (int q = 0; q < SL; ++q) {
            if (mask & (1u << q)) {
                const double xq = *reinterpret_cast<const double *>(xrow + D.sup_boff[q]);
                sum = __fma_rn(__ldg(vrow + k), xq, sum);
                ++k;
            }
        }
    } else {
        double xv[SL];
#pragma unroll
        for (int q = 0; q < SL; ++q) xv[q] = *reinterpret_cast<const double *>(xrow + D.sup_boff[q]);
        if (__all_sync(0xffffffffu, mask == (1u << SL) - 1u)) {      // interior slab: every row holds the whole pattern
#pragma unroll
            for (int q = 0; q < SL; ++q) sum = __fma_rn(D.sup_val[q], xv[q], sum);
        } else {
#pragma unroll
            for (int q = 0; q < SL; ++q)
                if (mask & (1u << q)) sum = __fma_rn(D.sup_val[q], xv[q], sum);
        }
    }
    return sum;
}
template <bool CLS_VALS>
__device__ __forceinline__ double tiled_row_sup_any(const char *xrow, const TiledDict &D, unsigned mask, const double *vrow) {
    if (D.sup_len == 7) return tiled_row_sup<CLS_VALS, 7>(xrow, D, mask, vrow);          // kernel-uniform
    if (D.sup_len == 5) return tiled_row_sup<CLS_VALS, 5>(xrow, D, mask, vrow);
    return tiled_row_sup<CLS_VALS, 8>(xrow, D, mask, vrow);
}
// The 4 slabs of a warp.  FULL: every row of the tile exists and the tile is staged (no per-row guards).
// xs == nullptr: the tile is not eligible for the windows, rows take the gather path of the CLASS kernel.
template <bool HAS_D, int NDOT, bool CLS_VALS, bool FULL>
__device__ __forceinline__ void tiled_slabs(const SpmvArgs &a, const TiledDict &D, const double *xs, const TiledSmemClass *sdict,
                                            unsigned cids, bool u_staged, const double (&upre)[kSlabsPerWarp], int tile, int warp, int lane) {
    static_assert(kSlabsPerWarp == 4, "packed slab sums assume 4 slabs per warp");
    const int row_base = tile * kTile;
    const bool tiled = FULL || xs != nullptr;
    // slab sums of the fused dots: the first butterfly step (distance 16) is taken pairwise as soon as two slabs of the
    // warp are done (packed_pair), which halves the values carried through the row loop
    double pp[NDOT > 0 ? NDOT : 1], w[NDOT > 0 ? NDOT * 2 : 1];
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        double p0 = 0.0, p1 = 0.0;
        const int row0 = row_base + (j * kCtaWarps + warp) * kSlab;
        if (FULL || row0 < a.n) {                                 // warp-uniform
            const int row = row0 + lane;
            const bool active = FULL || row < a.n;
            const int cid = active ? (int)((cids >> (8 * j)) & 0xffu) : 0;
            double uval = 0.0;
            if (NDOT >= 1 && !u_staged) uval = CLS_VALS ? upre[j] : (active ? __ldg(a.u + row) : 0.0);
            double sum;
            if (tiled) {
                const double *xrow = xs + (row - row_base);
                if (NDOT >= 1 && u_staged) uval = xrow[D.disp0];
                const double *vrow = CLS_VALS ? nullptr : a.val + __ldg(a.ia + (active ? row : row0));
                if (D.sup_len > 0) sum = tiled_row_sup_any<CLS_VALS>(reinterpret_cast<const char *>(xrow), D, active ? (unsigned)cid : 0u, vrow);
                else
                sum = tiled_row<CLS_VALS>(reinterpret_cast<const char *>(xrow), sdict + cid, vrow, active ? D.maxlen : 0, D.minlen);
            } else {
                sum = class_row_general<CLS_VALS>(a.x, a.val, a.ia, cid, row0, row, active, lane, D);
            }
            if (HAS_D) { if (active) sum = __dadd_rn(sum, __dmul_rn(__ldg(a.d + row), __ldg(a.x + row))); }
            if (active) a.y[row] = sum;
            if (NDOT >= 1) p0 = active ? __dmul_rn(sum, uval) : 0.0;
            if (NDOT >= 2) p1 = active ? __dmul_rn(sum, sum) : 0.0;
        }
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
    if constexpr (NDOT >= 1) {
        // remaining steps; the sum of (dot q, slab j) ends in the lanes with (lane>>4 & 1) + 2 (lane>>3 & 1) [+ 4 (lane>>2 & 1)]
        // == q * 4 + j.  Bit-identical to one warp_butterfly per slab sum (see packed_butterfly).
        double z = packed_pair(w[0], w[1], 8, lane);
        if constexpr (NDOT >= 2) z = packed_pair(z, packed_pair(w[2], w[3], 8, lane), 4, lane);
        else z = __dadd_rn(z, __shfl_xor_sync(0xffffffffu, z, 4));
        z = __dadd_rn(z, __shfl_xor_sync(0xffffffffu, z, 2));
        z = __dadd_rn(z, __shfl_xor_sync(0xffffffffu, z, 1));
        const int idx = ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1) + (NDOT >= 2 ? 4 * ((lane >> 2) & 1) : 0);
        const int j = idx & 3, q = idx >> 2;
        const bool writer = (lane & (NDOT >= 2 ? 3 : 7)) == 0;
        if (writer && (FULL || row_base + (j * kCtaWarps + warp) * kSlab < a.n))
            __stcg(a.rc.slab_part + (size_t)q * a.rc.slab_stride + tile * kTileSlabs + j * kCtaWarps + warp, z);
    }
}
#ifndef CUDAMAT_TILED_MINB
#define CUDAMAT_TILED_MINB 3
#endif
template <bool HAS_D, int NDOT, bool CLS_VALS>
__global__ void __launch_bounds__(kCtaThreads, CUDAMAT_TILED_MINB) k_spmv_tiled(const SpmvArgs a, const TiledArgs c,
                                                                                const __grid_constant__ TiledDict D) {
    extern __shared__ __align__(128) double xs[];
    __shared__ uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int row_base = tile * kTile;
    const bool tiled = __ldg(c.tile_ok + tile) != 0;               // CTA-uniform, solve-constant
    // per row: the presence mask w.r.t. the superset pattern when the tile is staged and the pattern exists, else the class id
    const unsigned char *ids = (tiled && D.sup_len > 0) ? c.tmask : c.cls;
    int cid[kSlabsPerWarp];
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + (j * kCtaWarps + warp) * kSlab + lane;
        cid[j] = (row < a.n) ? (int)__ldg(ids + row) : 0xff;
    }
    const TiledSmemClass *sdict = reinterpret_cast<const TiledSmemClass *>(xs + D.sdict_base);
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_sync();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    halo_wait(a.hw, tile, a.sc ? &a.sc->status : nullptr);
    // SpMV 2 of the loop takes its dot operand from x itself (t.s): it is already in shared memory.  Any other operand
    // is fetched here, ahead of the wait for the windows, so that its DRAM latency overlaps the bulk copies (values
    // dictionary only: the kernel that streams the values from CSR has no registers to spare for it).
    const bool u_staged = NDOT >= 1 && tiled && a.u == a.x && D.disp0 >= 0;
    double upre[kSlabsPerWarp];
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + (j * kCtaWarps + warp) * kSlab + lane;
        upre[j] = (CLS_VALS && NDOT >= 1 && !u_staged && row < a.n) ? __ldg(a.u + row) : 0.0;
    }
    if (tiled) {
        if (warp == 0) {
            // lane g clips and issues window g (the windows are independent: no serial loop in one thread)
            uint32_t bytes = 0;
            long long gs = 0, cs = 0;
            if (lane < D.nseg) {
                gs = (long long)row_base + D.seg_lo[lane];
                cs = max(gs, 0LL);
                const long long ce = min(gs + D.seg_len[lane], (long long)c.nx), cnt = ce - cs;
                if (cnt > 0) {
                    bytes = (uint32_t)(cnt & ~1LL) * 8u;
                    if (cnt & 1) xs[D.seg_base[lane] + (int)(ce - 1 - gs)] = a.x[ce - 1];         // odd tail element
                }
            }
            const uint32_t dict_bytes = D.sup_len > 0 ? 0u : (uint32_t)sizeof(TiledSmemClass) * (uint32_t)c.ncls;
            const uint32_t total = __reduce_add_sync(0xffffffffu, bytes) + dict_bytes;
            if (lane == 0) {
                mbar_expect_tx(&s_bar, total);
                if (dict_bytes) tma_bulk_g2s((void *)sdict, c.sdict, dict_bytes, &s_bar);
            }
            __syncwarp();
            if (bytes) tma_bulk_g2s(xs + D.seg_base[lane] + (int)(cs - gs), a.x + cs, bytes, &s_bar);
        }
        if (warp == 0) {                                           // one warp probes the barrier, the others sleep at bar.sync
            unsigned spins = 0;
            while (!mbar_try_wait(&s_bar, 0)) { if (++spins > (1u << 26)) __trap(); }
        }
        __syncthreads();
    }
    const unsigned cids = (unsigned)(cid[0] & 0xff) | ((unsigned)(cid[1] & 0xff) << 8) | ((unsigned)(cid[2] & 0xff) << 16) | ((unsigned)cid[3] << 24);
    if (tiled && row_base + kTile <= a.n) tiled_slabs<HAS_D, NDOT, CLS_VALS, true>(a, D, xs, sdict, cids, u_staged, upre, tile, warp, lane);
    else tiled_slabs<HAS_D, NDOT, CLS_VALS, false>(a, D, tiled ? xs : nullptr, sdict, cids, u_staged, upre, tile, warp, lane);
}

struct StagedArgs {
    SpmvArgs a;
    int cap_nnz;      // smem capacity per slab stage in nnz, multiple of 16
    int stages;
    int64_t nnz;      // to bound aligned over-reads
};

// Per-warp stage layout (every region a multiple of 16 bytes):
//   double val[cap+2] | int col[cap+8] | int rowptr[36] | double u[32]
__host__ __device__ __forceinline__ int staged_stage_bytes(int cap) { return (cap + 2) * 8 + (cap + 8) * 4 + 144 + 256; }

// row sum of one lane's row; vb/cb are rebased so that the GLOBAL nnz index k addresses vb[k]/cb[k].
// The first 8 gathered x values (xv) were prefetched while the previous slab was being computed.
__device__ __forceinline__ double staged_rowsum(const double *vb, const int *cb, const double *x, int s, int e,
                                                const double (&xv)[8], int lane) {
    const int len = e - s;
    const int shortlen = (len <= kLongRow) ? len : 0;
    const int maxlen = __reduce_max_sync(0xffffffffu, shortlen);
    double sum = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (q < shortlen) sum = __fma_rn(vb[s + q], xv[q], sum);
    for (int k0 = 8; k0 < maxlen; k0 += 8) {
        int cj[8]; double xq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) cj[q] = ((k0 + q) < shortlen) ? cb[s + k0 + q] : -1;
#pragma unroll
        for (int q = 0; q < 8; ++q) xq[q] = (cj[q] >= 0) ? __ldg(x + cj[q]) : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if ((k0 + q) < shortlen) sum = __fma_rn(vb[s + k0 + q], xq[q], sum);
    }
    unsigned lm = __ballot_sync(0xffffffffu, len > kLongRow);
    while (lm) {
        const int src = __ffs(lm) - 1;
        lm &= lm - 1;
        const int ss = __shfl_sync(0xffffffffu, s, src), ee = __shfl_sync(0xffffffffu, e, src);
        double acc = 0.0;
        for (int k = ss + lane; k < ee; k += 32) acc = __fma_rn(vb[k], __ldg(x + cb[k]), acc);
        acc = warp_butterfly(acc);
        if (lane == src) sum = acc;
    }
    return sum;
}

template <bool HAS_D, int NDOT>
__global__ void __launch_bounds__(kCtaThreads, 1) k_spmv_staged(const StagedArgs g) {
    const SpmvArgs &a = g.a;
    pdl_prologue();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    halo_wait(a.hw, blockIdx.x, a.sc ? &a.sc->status : nullptr);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_bar[kCtaWarps][8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_base = blockIdx.x * kTile;
    const int S = g.stages;
    const int VB = (g.cap_nnz + 2) * 8, CB = (g.cap_nnz + 8) * 4;
    const int stage_bytes = staged_stage_bytes(g.cap_nnz);
    unsigned char *wbase = smem_raw + (size_t)warp * S * stage_bytes;
    if (lane == 0)
        for (int q = 0; q < S; ++q) mbar_init(&s_bar[warp][q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    const int rows_here = min(kTile, a.n - row_base);
    const int nslab_tile = (rows_here + kSlab - 1) / kSlab;
    const int nj = (nslab_tile > warp) ? (nslab_tile - warp + kCtaWarps - 1) / kCtaWarps : 0;
    // lane j (< 8) keeps the nnz span of slab j of this warp; "full" = 32 valid rows and a 36-int row
    // pointer copy that stays inside ia[0..n]
    int my_s = 0, my_e = 0, my_full = 0;
    if (lane < nj) {
        const int r0 = row_base + (lane * kCtaWarps + warp) * kSlab;
        my_s = __ldg(a.ia + r0);
        my_e = __ldg(a.ia + min(r0 + kSlab, a.n));
        my_full = (r0 + 35 <= a.n) ? 1 : 0;
    }
    unsigned phase_bits = 0, parity_mask = 0, staged_mask = 0;
    auto issue = [&](int j) {
        const int q = j % S;
        const int s0 = __shfl_sync(0xffffffffu, my_s, j), e0 = __shfl_sync(0xffffffffu, my_e, j);
        const int full = __shfl_sync(0xffffffffu, my_full, j);
        const int vs = s0 & ~1, ve = (e0 + 1) & ~1;
        const int cs = s0 & ~3, ce = (e0 + 3) & ~3;
        const bool fits = full && e0 > s0 && (e0 - s0) <= g.cap_nnz && (int64_t)ve <= (g.nnz & ~1LL) && (int64_t)ce <= (g.nnz & ~3LL);
        if (fits) {
            if (lane == 0) {
                unsigned char *st = wbase + (size_t)q * stage_bytes;
                const int r0 = row_base + (j * kCtaWarps + warp) * kSlab;
                const uint32_t vb = (uint32_t)(ve - vs) * 8u, cb = (uint32_t)(ce - cs) * 4u;
                mbar_expect_tx(&s_bar[warp][q], vb + cb + 144u + (NDOT >= 1 ? 256u : 0u));
                tma_bulk_g2s(st, a.val + vs, vb, &s_bar[warp][q]);
                tma_bulk_g2s(st + VB, a.ja + cs, cb, &s_bar[warp][q]);
                tma_bulk_g2s(st + VB + CB, a.ia + r0, 144u, &s_bar[warp][q]);
                if (NDOT >= 1) tma_bulk_g2s(st + VB + CB + 144, a.u + r0, 256u, &s_bar[warp][q]);
            }
            staged_mask |= 1u << j;
            parity_mask |= ((phase_bits >> q) & 1u) << j;
            phase_bits ^= 1u << q;
        }
    };
    // first stage of a slab: make its operands visible, fetch the lane's row bounds and issue the
    // x gathers of the first 8 entries (they complete while the previous slab is being computed)
    auto prefetch = [&](int j, int &s, int &e, double (&xv)[8]) {
        const int row = row_base + (j * kCtaWarps + warp) * kSlab + lane;
        const bool staged = (staged_mask >> j) & 1u;
        const int *cbase;
        if (staged) {
            const int q = j % S;
            unsigned spins = 0;
            while (!mbar_try_wait(&s_bar[warp][q], (parity_mask >> j) & 1u)) { if (++spins > (1u << 24)) __trap(); }
            const unsigned char *st = wbase + (size_t)q * stage_bytes;
            const int *rp = reinterpret_cast<const int *>(st + VB + CB);
            s = rp[lane]; e = rp[lane + 1];
            cbase = reinterpret_cast<const int *>(st + VB) - (__shfl_sync(0xffffffffu, my_s, j) & ~3);
        } else {
            s = 0; e = 0;
            if (row < a.n) { s = __ldg(a.ia + row); e = __ldg(a.ia + row + 1); }
            cbase = a.ja;
            (void)__shfl_sync(0xffffffffu, my_s, j);
        }
        const int len = e - s;
        const int shortlen = (len <= kLongRow) ? len : 0;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) xv[q8] = (q8 < shortlen) ? __ldg(a.x + cbase[s + q8]) : 0.0;
    };

    for (int j = 0; j < S - 1 && j < nj; ++j) issue(j);
    int cs_ = 0, ce_ = 0; double cxv[8];
    if (nj > 0) prefetch(0, cs_, ce_, cxv);
#pragma unroll 1
    for (int j = 0; j < nj; ++j) {
        if (j + S - 1 < nj) { __syncwarp(); issue(j + S - 1); }
        int ns_ = 0, ne_ = 0; double nxv[8];
        if (j + 1 < nj) prefetch(j + 1, ns_, ne_, nxv);
        const int slab = j * kCtaWarps + warp;
        const int row = row_base + slab * kSlab + lane;
        const bool active = row < a.n;
        const bool staged = (staged_mask >> j) & 1u;
        const int span_s = __shfl_sync(0xffffffffu, my_s, j);
        double sum, uval = 0.0;
        if (staged) {
            const unsigned char *st = wbase + (size_t)(j % S) * stage_bytes;
            const double *vb = reinterpret_cast<const double *>(st) - (span_s & ~1);
            const int *cb = reinterpret_cast<const int *>(st + VB) - (span_s & ~3);
            sum = staged_rowsum(vb, cb, a.x, cs_, ce_, cxv, lane);
            if (NDOT >= 1) uval = reinterpret_cast<const double *>(st + VB + CB + 144)[lane];
        } else {
            sum = staged_rowsum(a.val, a.ja, a.x, cs_, ce_, cxv, lane);
            if (NDOT >= 1) uval = active ? __ldg(a.u + row) : 0.0;
        }
        if (HAS_D) { if (active) sum = __dadd_rn(sum, __dmul_rn(__ldg(a.d + row), __ldg(a.x + row))); }
        if (active) a.y[row] = sum;
        if (NDOT >= 1) slab_deposit(a.rc, 0, blockIdx.x * kTileSlabs + (slab), active ? __dmul_rn(sum, uval) : 0.0, lane);
        if (NDOT >= 2) slab_deposit(a.rc, 1, blockIdx.x * kTileSlabs + (slab), active ? __dmul_rn(sum, sum) : 0.0, lane);
        cs_ = ns_; ce_ = ne_;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) cxv[q8] = nxv[q8];
    }
    if (NDOT >= 1) {
    }
}

int plan_staged(cudamat_solver *s) {
    // capacity = max slab span rounded up to 16; as many stages as fit ~110 KB per CTA (2 CTAs / SM)
    s->staged = StagedPlan();
    if (s->max_slab_nnz <= 0) return CUDAMAT_OK;
    int cap = ((s->max_slab_nnz + 15) / 16) * 16;
    if (cap > 1024) cap = 1024;                        // heavier slabs take the direct path
    const size_t stage_bytes = (size_t)staged_stage_bytes(cap);
    int stages = (int)((110 * 1024) / (stage_bytes * kCtaWarps));
    if (s->opt_staged_stages > 0) stages = s->opt_staged_stages;
    if (stages > 8) stages = 8;
    if (stages < 2) return CUDAMAT_OK;
    if (stage_bytes * kCtaWarps * stages > 220 * 1024) return CUDAMAT_OK;
    s->staged.cap_nnz = cap;
    s->staged.stages = stages;
    s->staged.smem_bytes = stage_bytes * kCtaWarps * stages;
    return CUDAMAT_OK;
}

// launch with programmatic stream serialization (PDL): the kernel may be scheduled while its predecessor
// in the stream drains; its pdl_prologue() waits for the predecessor's completion before reading.
bool pdl_enabled() {                      // CUDAMAT_NO_PDL=1 turns the PDL launch attribute off (tuning / debugging)
    static const int on = [] { const char *e = getenv("CUDAMAT_NO_PDL"); return (e && *e && *e != '0') ? 0 : 1; }();
    return on != 0;
}
template <typename Kern, typename Arg>
static cudaError_t launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t st, const Arg &arg) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, arg);
}

template <typename Kern, typename A1, typename A2, typename A3>
static cudaError_t launch_pdl2(Kern kern, int grid, int block, cudaStream_t st, const A1 &a1, const A2 &a2, const A3 &a3) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a1, a2, a3);
}

template <bool HAS_D, int NDOT>
static int launch_spmv_t(cudamat_solver *s, const SpmvArgs &a, int variant) {
    const int grid = (a.n + kTile - 1) / kTile;
    if (grid == 0) return CUDAMAT_OK;
    if (variant == CUDAMAT_SPMV_TILED && ((uintptr_t)a.x % 16) == 0 && (s->cls[1].h_tdict || s->cls[0].h_tdict)) {
        const int m = s->cls[1].h_tdict ? 1 : 0;                   // 1: values from the dictionary, 0: values from CSR
        const RowClasses &C = s->cls[m];
        const TiledArgs c{C.d_cls, C.d_tile_ok, C.d_tmask, C.d_sdict, C.ncls, s->n + s->nhalo};
        const void *kern = m ? (const void *)k_spmv_tiled<HAS_D, NDOT, true> : (const void *)k_spmv_tiled<HAS_D, NDOT, false>;
        // function attributes are per device and per template instance: one flag per (device, instance)
        static bool attr_set[64][2] = {};
        const int dv = s->device & 63;
        if (!attr_set[dv][m]) { CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024)); attr_set[dv][m] = true; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kCtaThreads); cfg.dynamicSmemBytes = C.tiled_smem; cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        void *args[] = {(void *)&a, (void *)&c, (void *)C.h_tdict};
        CM_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
        s->launches++;
        CM_CUDA(cudaGetLastError());
        return CUDAMAT_OK;
    }
    if (variant == CUDAMAT_SPMV_TILED) variant = CUDAMAT_SPMV_CLASS;
    if (variant == CUDAMAT_SPMV_CLASS && s->cls[1].ncls > 0) {
        const int T = std::max(1, s->opt_class_tiles_per_cta);
        const ClassArgs c{s->cls[1].d_cls, s->cls[1].ncls, T};
        CM_CUDA(launch_pdl2(k_spmv_class<HAS_D, NDOT, true>, (grid + T - 1) / T, kCtaThreads, s->stream, a, c, *s->cls[1].h_dict));
        s->launches++;
        CM_CUDA(cudaGetLastError());
        return CUDAMAT_OK;
    }
    if ((variant == CUDAMAT_SPMV_PATTERN || variant == CUDAMAT_SPMV_CLASS) && s->cls[0].ncls > 0) {
        const int T = std::max(1, s->opt_class_tiles_per_cta);
        const ClassArgs c{s->cls[0].d_cls, s->cls[0].ncls, T};
        CM_CUDA(launch_pdl2(k_spmv_class<HAS_D, NDOT, false>, (grid + T - 1) / T, kCtaThreads, s->stream, a, c, *s->cls[0].h_dict));
        s->launches++;
        CM_CUDA(cudaGetLastError());
        return CUDAMAT_OK;
    }
    if (variant == CUDAMAT_SPMV_STAGED && s->staged.cap_nnz > 0 && (NDOT == 0 || ((uintptr_t)a.u % 16) == 0)) {
        StagedArgs g{a, s->staged.cap_nnz, s->staged.stages, s->nnz};
        auto kern = k_spmv_staged<HAS_D, NDOT>;
        CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->staged.smem_bytes));
        CM_CUDA(launch_pdl(kern, grid, kCtaThreads, s->staged.smem_bytes, s->stream, g));
    } else {
        CM_CUDA(launch_pdl(k_spmv_rowlane<HAS_D, NDOT>, grid, kCtaThreads, 0, s->stream, a));
    }
    s->launches++;
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

static int launch_spmv_any(cudamat_solver *s, const SpmvArgs &a, int variant) {
    if (variant == CUDAMAT_SPMV_AUTO) variant = s->spmv_variant;
    if (variant == CUDAMAT_SPMV_MARCH) {
        if (march_spmv_usable(s, a)) return launch_march_spmv(s, a);
        variant = CUDAMAT_SPMV_TILED;
    }
    if (variant == CUDAMAT_SPMV_STREAM) return launch_stream_spmv(s, a);
    const bool hd = a.d != nullptr;
    switch (a.ndot) {
    case 0: return hd ? launch_spmv_t<true, 0>(s, a, variant) : launch_spmv_t<false, 0>(s, a, variant);
    case 1: return hd ? launch_spmv_t<true, 1>(s, a, variant) : launch_spmv_t<false, 1>(s, a, variant);
    default: return hd ? launch_spmv_t<true, 2>(s, a, variant) : launch_spmv_t<false, 2>(s, a, variant);
    }
}
int launch_spmv(cudamat_solver *s, const SpmvArgs &a, int variant) {       // exactly one kernel: event-bracketed when sampled
    int rc = ev_mark(s, true);
    if (rc) return rc;
    if ((rc = launch_spmv_any(s, a, variant))) return rc;
    return ev_mark(s, false);
}

// ------------------------------------------------------------------------------------------
// Fused vector updates. Thread t of CTA b owns rows b*2048 + j*256 + t, j = 0..7, i.e. warp w
// owns slabs j*8 + w — the same tile/slab geometry as the SpMV, so reductions share the tail.
// ------------------------------------------------------------------------------------------
struct VecArgs {
    int n;
    const double *in0, *in1, *in2, *in3, *in4;
    double *out0, *out1, *out2;
    RedCtx rc; DevScalars *sc; double *hist; int phase;
    HaloPush hp;           // multi-GPU peer-memory path: out0's halo rows go straight to the neighbours (npeer == 0: off)
};

#define VEC_PROLOGUE                                                                  \
    pdl_prologue();                                                                   \
    if (a.sc->status != ST_RUNNING) return;                                           \
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                    \
    const int row_base = blockIdx.x * kTile;                                          \
    (void)warp; (void)lane;

// r = b - y ; c1 = r ; c2 = r (optional) ; red0 = r.r          (pbicgstab.cu:67-74, 645-655)
__global__ void __launch_bounds__(kCtaThreads) k_init_resid(const VecArgs a) {
    VEC_PROLOGUE
    double rv[kSlabsPerWarp];
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        rv[j] = (row < a.n) ? __dsub_rn(__ldg(a.in0 + row), __ldg(a.in1 + row)) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        const bool active = row < a.n;
        if (active) {
            a.out0[row] = rv[j];
            a.out1[row] = rv[j];
            if (a.out2) a.out2[row] = rv[j];
        }
        if (row_base + j * kCtaThreads + warp * kSlab < a.n)
            slab_deposit(a.rc, 0, blockIdx.x * kTileSlabs + (j * kCtaWarps + warp), active ? __dmul_rn(rv[j], rv[j]) : 0.0, lane);
    }
}

// p' = r + beta*(p - omega*v)
//   unprec form (pbicgstab.cu:668-672): q=fl(-omega*v); q=fl(p+q); q=fl(beta*q); p'=fl(r+q)
//   ilu0 form   (pbicgstab.cu:86-88):   q=fma(-omega,v,p); q=fl(beta*q); p'=fma(1,r,q)
//   (ilu0: skipped on the first pass, i == 0, where p = r from the init)
template <bool FMA_FORM>
__global__ void __launch_bounds__(kCtaThreads) k_update_p(const VecArgs a) {
    VEC_PROLOGUE
    if (FMA_FORM && a.sc->iter == 0) return;
    const double beta = a.sc->beta, momega = -a.sc->omega;
    double rr[kSlabsPerWarp], vv[kSlabsPerWarp], pp[kSlabsPerWarp];
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        const bool act = row < a.n;
        rr[j] = act ? __ldg(a.in0 + row) : 0.0;
        vv[j] = act ? __ldg(a.in1 + row) : 0.0;
        pp[j] = act ? a.out0[row] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        if (row < a.n) {
            double q;
            if (FMA_FORM) {
                q = __fma_rn(momega, vv[j], pp[j]);
                q = __dmul_rn(beta, q);
                q = __fma_rn(1.0, rr[j], q);
            } else {
                q = __dmul_rn(momega, vv[j]);
                q = __dadd_rn(pp[j], q);
                q = __dmul_rn(beta, q);
                q = __dadd_rn(rr[j], q);
            }
            a.out0[row] = q;
            halo_store(a.hp, row, q);
        }
    }
    halo_signal(a.hp, row_base, min(kTile, a.n - row_base));
}

// s = r + fl(-alpha*v)                                            (pbicgstab.cu:698-700)
__global__ void __launch_bounds__(kCtaThreads) k_update_s(const VecArgs a) {
    VEC_PROLOGUE
    const double malpha = -a.sc->alpha;
    double rr[kSlabsPerWarp], vv[kSlabsPerWarp];
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        const bool act = row < a.n;
        rr[j] = act ? __ldg(a.in0 + row) : 0.0;
        vv[j] = act ? __ldg(a.in1 + row) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        if (row < a.n) {
            const double q = __dadd_rn(rr[j], __dmul_rn(malpha, vv[j]));
            a.out0[row] = q;
            halo_store(a.hp, row, q);
        }
    }
    halo_signal(a.hp, row_base, min(kTile, a.n - row_base));
}

// the same update on the first and the last plane of a slab shard only (S tiles each): the rows the neighbours need, pushed and
// flagged BEFORE the folded SpMV 2 (MARCH MAKE_S) forms s for all rows on the fly — the full-size pass over r, v, s goes away
__global__ void __launch_bounds__(kCtaThreads) k_update_s_boundary(const VecArgs a, int plane_tiles, int ntile) {
    pdl_prologue();
    if (a.sc->status != ST_RUNNING) return;
    const int tid = threadIdx.x;
    const int tile = (int)blockIdx.x < plane_tiles ? (int)blockIdx.x : ntile - 2 * plane_tiles + (int)blockIdx.x;
    const int row_base = tile * kTile;
    const double malpha = -a.sc->alpha;
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        if (row < a.n) {
            const double q = __dadd_rn(__ldg(a.in0 + row), __dmul_rn(malpha, __ldg(a.in1 + row)));
            a.out0[row] = q;
            halo_store(a.hp, row, q);
        }
    }
    halo_signal(a.hp, row_base, min(kTile, a.n - row_base));
}

// ilu0: r = fma(-alpha,v,r) ; x = fma(alpha,pw,x) ; red0 = r.r   (pbicgstab.cu:109-111)
__global__ void __launch_bounds__(kCtaThreads) k_update_rx_ilu(const VecArgs a) {
    VEC_PROLOGUE
    const double alpha = a.sc->alpha, malpha = -alpha;
    double vv[kSlabsPerWarp], pw[kSlabsPerWarp], rr[kSlabsPerWarp], xx[kSlabsPerWarp];
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        const bool act = row < a.n;
        vv[j] = act ? __ldg(a.in0 + row) : 0.0;
        pw[j] = act ? __ldg(a.in1 + row) : 0.0;
        rr[j] = act ? a.out0[row] : 0.0;
        xx[j] = act ? a.out1[row] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        const bool act = row < a.n;
        const double rn = __fma_rn(malpha, vv[j], rr[j]);
        const double xn = __fma_rn(alpha, pw[j], xx[j]);
        if (act) { a.out0[row] = rn; a.out1[row] = xn; }
        if (row_base + j * kCtaThreads + warp * kSlab < a.n)
            slab_deposit(a.rc, 0, blockIdx.x * kTileSlabs + (j * kCtaWarps + warp), act ? __dmul_rn(rn, rn) : 0.0, lane);
    }
}

// end-of-iteration update with the two fused dots.
//   unprec (pbicgstab.cu:694-696,714-723,665): h=fl(x+fl(alpha*p)); x=fl(h+fl(omega*s));
//           r=fl(s+fl(-omega*t)); red0 = rhat.r ; red1 = r.r
//   ilu0   (pbicgstab.cu:139-142,81):         x=fma(omega,s,x); r=fma(-omega,t,r); same dots
// in0 = p, in1 = s, in2 = t, in3 = rhat ; out0 = x, out1 = r
template <bool FMA_FORM>
__global__ void __launch_bounds__(kCtaThreads) k_update_xr(const VecArgs a) {
    VEC_PROLOGUE
    const double alpha = a.sc->alpha, omega = a.sc->omega, momega = -omega;
#pragma unroll 2
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        const bool act = row < a.n;
        double xn = 0.0, rn = 0.0, rh = 0.0;
        if (act) {
            const double sv = __ldg(a.in1 + row), tv = __ldg(a.in2 + row);
            rh = __ldg(a.in3 + row);
            const double xo = a.out0[row];
            if (FMA_FORM) {
                const double ro = a.out1[row];
                xn = __fma_rn(omega, sv, xo);
                rn = __fma_rn(momega, tv, ro);
            } else {
                const double pv = __ldg(a.in0 + row);
                const double h = __dadd_rn(xo, __dmul_rn(alpha, pv));
                xn = __dadd_rn(h, __dmul_rn(omega, sv));
                rn = __dadd_rn(sv, __dmul_rn(momega, tv));
            }
            a.out0[row] = xn;
            a.out1[row] = rn;
        }
        if (row_base + j * kCtaThreads + warp * kSlab < a.n) {
            slab_deposit(a.rc, 0, blockIdx.x * kTileSlabs + (j * kCtaWarps + warp), act ? __dmul_rn(rh, rn) : 0.0, lane);
            slab_deposit(a.rc, 1, blockIdx.x * kTileSlabs + (j * kCtaWarps + warp), act ? __dmul_rn(rn, rn) : 0.0, lane);
        }
    }
}

// spec dot product of two arbitrary vectors -> sc->red[0]
__global__ void __launch_bounds__(kCtaThreads) k_dot(const VecArgs a) {
    pdl_prologue();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_base = blockIdx.x * kTile;
#pragma unroll
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int row = row_base + j * kCtaThreads + tid;
        const bool act = row < a.n;
        const double p = act ? __dmul_rn(__ldg(a.in0 + row), __ldg(a.in1 + row)) : 0.0;
        if (row_base + j * kCtaThreads + warp * kSlab < a.n) slab_deposit(a.rc, 0, blockIdx.x * kTileSlabs + (j * kCtaWarps + warp), p, lane);
    }
}

// ------------------------------------------------------------------------------------------
// k_reduce_finish: launched right behind every reducing kernel (PDL).  Spec tree: slab sums
// (written by the reducing kernel) -> tile partials R(64 slabs) -> group partials R(1024 tiles) -> result
// R(groups) -> scalar recurrence.  Every warp of the grid forms tile partials; the last CTA to finish (ticket)
// does the rest:
//   stage 0: everything; on a sharded handle with the peer-memory path the group (or tile) partials of this rank
//            are pushed to every rank and the CTA waits for all arrivals before the final sum;
//   stage 1: up to the local group partials (NCCL path, before the allreduce);
//   stage 2: final sum + recurrence from `glob` (NCCL path, after the allreduce; one CTA).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads) k_reduce_finish(const RedCtx rc, int nq, DevScalars *sc, double *hist, int phase,
                                                               int stage, const double *glob, const unsigned long long *flags) {
    pdl_sync();
    if (phase != PH_STORE && sc->status != ST_RUNNING) return;
    __shared__ double s_grp[kMaxQ][64];
    __shared__ double s_red[kMaxQ];
    __shared__ int s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (stage != 2) {
        // tile partials: job = (q, local tile), one warp each
        const int njobs = nq * rc.ntile;
        for (int w = blockIdx.x * kCtaWarps + warp; w < njobs; w += gridDim.x * kCtaWarps) {
            const int q = w / rc.ntile, tile = w % rc.ntile;
            const int nslab = min(kTileSlabs, (rc.n_local - tile * kTile + kSlab - 1) / kSlab);
            const double tp = warp_reduce_values_cg(rc.slab_part + (size_t)q * rc.slab_stride + (size_t)tile * kTileSlabs, nslab, lane);
            if (lane == 0) {
                if (rc.exch_level == 1) __stcg(rc.exch + (size_t)q * rc.exch_stride + rc.tile0 + tile, tp);
                __stcg(rc.tile_part + (size_t)q * rc.tile_stride + tile, tp);
            }
        }
        // last CTA of this (small) grid continues
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned old = atomicAdd(rc.done_cnt, 1u);
            s_last = (old == gridDim.x - 1);
            if (s_last) *rc.done_cnt = 0u;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (rc.exch_level != 1) {                              // local groups into the slots array (global group index)
            for (int w = warp; w < nq * rc.ngroup_loc; w += kCtaWarps) {
                const int q = w / rc.ngroup_loc, g = w % rc.ngroup_loc;
                const int in_group = min(kGroupTiles, rc.ntile - g * kGroupTiles);
                const double gp = warp_reduce_values_cg(rc.tile_part + (size_t)q * rc.tile_stride + (size_t)g * kGroupTiles, in_group, lane);
                if (lane == 0) __stcg(rc.slots + (size_t)q * rc.slot_stride + rc.group0 + g, gp);
            }
            __threadfence_block();
            __syncthreads();
        }
    }
    if (stage == 1) return;
    const double *src = (stage == 2) ? glob : (rc.exch_level == 1 ? rc.exch : rc.slots);
    const int stride = (rc.exch_level == 1) ? rc.exch_stride : rc.slot_stride;
    if (stage == 0 && rc.p2p.world > 0) {
        p2p_push(rc, src, nq);
        if ((int)threadIdx.x < rc.p2p.world) {                 // every rank's partial sums must have arrived
            unsigned spins = 0;
            while (ld_acquire_sys_u64(flags + threadIdx.x) < rc.p2p.epoch)
                if (++spins > kPeerSpinLimit) { atomicExch(&sc->status, ST_COMM_TIMEOUT); break; }   // a lost rank must not hang the GPU
        }
        __syncthreads();
        src = rc.p2p.peers[rc.p2p.me].gather[rc.p2p.epoch & 1ull];
    }
    int nfinal = rc.nslots;
    if (rc.exch_level == 1) {                                  // unaligned shards: groups over the GLOBAL tile index
        const int ngroups = (rc.ntile_global + kGroupTiles - 1) / kGroupTiles;     // <= 64 (checked on the host)
        for (int w = warp; w < nq * ngroups; w += kCtaWarps) {
            const int q = w / ngroups, g = w % ngroups;
            const double v = warp_reduce_values_cg(src + (size_t)q * stride + (size_t)g * kGroupTiles,
                                                   min(kGroupTiles, rc.ntile_global - g * kGroupTiles), lane);
            if (lane == 0) s_grp[q][g] = v;
        }
        __syncthreads();
        nfinal = ngroups;
    }
    if (warp < nq) {
        const double f = (rc.exch_level == 1) ? warp_reduce_values_smem(s_grp[warp], nfinal, lane)
                                              : warp_reduce_values_cg(src + (size_t)warp * stride, nfinal, lane);
        if (lane == 0) s_red[warp] = f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double red[kMaxQ] = {0.0, 0.0};
        for (int q = 0; q < nq; ++q) red[q] = s_red[q];
        apply_phase(sc, hist, phase, red);
    }
}
int launch_reduce_finish(cudamat_solver *s, const RedCtx &rc, int nq, int phase, int stage, const double *glob,
                         const unsigned long long *flags) {
    cudaLaunchConfig_t cfg{};
    // one warp per (quantity, tile); a few hundred CTAs at most, so the ticket / fence cost stays negligible
    int grid = 1;
    static const int cap = [] { const char *e = getenv("CUDAMAT_RF_GRID"); return e && *e ? atoi(e) : 592; }();
    if (stage != 2) grid = std::max(1, std::min((nq * rc.ntile + kCtaWarps - 1) / kCtaWarps, cap));    // 4 CTAs per SM resident: one (quantity, tile) per warp up to 256^3
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kCtaThreads); cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CM_CUDA(cudaLaunchKernelEx(&cfg, k_reduce_finish, rc, nq, s->d_sc, s->d_hist, phase, stage, glob, flags));
    s->launches++;
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

__global__ void k_fill(double *p, double v, int64_t cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < cnt; i += stride) p[i] = v;
}

static VecArgs vec_args(cudamat_solver *s, int phase) {
    VecArgs a{};
    a.n = s->n; a.rc = s->rc; a.sc = s->d_sc; a.hist = s->d_hist; a.phase = phase;
    if (phase != PH_NONE) comm_begin_reduction(s, a.rc);       // reducing kernel: stamp the reduction epoch (p2p)
    return a;
}
static inline int tiles_of(int n) { return (n + kTile - 1) / kTile; }

#define LAUNCH_VEC(kern, a)                                                    \
    do {                                                                       \
        const int grid_ = tiles_of((a).n);                                     \
        if (grid_ > 0) {                                                       \
            { const int e_ = ev_mark(s, true); if (e_) return e_; }            \
            CM_CUDA(launch_pdl(kern, grid_, kCtaThreads, 0, s->stream, (a)));  \
            s->launches++;                                                     \
            CM_CUDA(cudaGetLastError());                                       \
            { const int e_ = ev_mark(s, false); if (e_) return e_; }           \
        }                                                                      \
    } while (0)

int launch_init_resid(cudamat_solver *s, const double *b, const double *y, double *r, double *c1, double *c2, int phase) {
    VecArgs a = vec_args(s, phase);
    a.in0 = b; a.in1 = y; a.out0 = r; a.out1 = c1; a.out2 = c2;
    LAUNCH_VEC(k_init_resid, a);
    return finish_reduction(s, a.rc, phase, 1);
}
int launch_update_p(cudamat_solver *s, bool fma_form, const double *r, const double *v, double *p, const HaloPush *hp) {
    VecArgs a = vec_args(s, PH_NONE);
    a.in0 = r; a.in1 = v; a.out0 = p;
    if (hp) a.hp = *hp;
    if (fma_form) LAUNCH_VEC(k_update_p<true>, a); else LAUNCH_VEC(k_update_p<false>, a);
    return CUDAMAT_OK;
}
int launch_update_s(cudamat_solver *s, const double *r, const double *v, double *sv, const HaloPush *hp) {
    VecArgs a = vec_args(s, PH_NONE);
    a.in0 = r; a.in1 = v; a.out0 = sv;
    if (hp) a.hp = *hp;
    LAUNCH_VEC(k_update_s, a);
    return CUDAMAT_OK;
}
int launch_update_s_boundary(cudamat_solver *s, const double *r, const double *v, double *sv, const HaloPush *hp, int plane_tiles) {
    VecArgs a = vec_args(s, PH_NONE);
    a.in0 = r; a.in1 = v; a.out0 = sv;
    if (hp) a.hp = *hp;
    const int ntile = tiles_of(a.n);
    if (ntile < 2 * plane_tiles) return launch_update_s(s, r, v, sv, hp);
    { const int e_ = ev_mark(s, true); if (e_) return e_; }
    CM_CUDA(launch_pdl2(k_update_s_boundary, 2 * plane_tiles, kCtaThreads, s->stream, a, plane_tiles, ntile));
    s->launches++;
    CM_CUDA(cudaGetLastError());
    { const int e_ = ev_mark(s, false); if (e_) return e_; }
    return CUDAMAT_OK;
}
int launch_update_rx_ilu(cudamat_solver *s, const double *v, const double *pw, double *r, double *x) {
    VecArgs a = vec_args(s, PH_I_A2);
    a.in0 = v; a.in1 = pw; a.out0 = r; a.out1 = x;
    LAUNCH_VEC(k_update_rx_ilu, a);
    return finish_reduction(s, a.rc, PH_I_A2, 1);
}
int launch_update_xr(cudamat_solver *s, bool fma_form, const double *p, const double *sv, const double *t,
                     const double *rhat, double *x, double *r) {
    VecArgs a = vec_args(s, fma_form ? PH_I_C : PH_U_C);
    a.in0 = p; a.in1 = sv; a.in2 = t; a.in3 = rhat; a.out0 = x; a.out1 = r;
    if (fma_form) LAUNCH_VEC(k_update_xr<true>, a); else LAUNCH_VEC(k_update_xr<false>, a);
    return finish_reduction(s, a.rc, a.phase, 2);
}
int launch_dot(cudamat_solver *s, const double *x, const double *y) {
    VecArgs a = vec_args(s, PH_STORE);
    a.in0 = x; a.in1 = y;
    LAUNCH_VEC(k_dot, a);
    return finish_reduction(s, a.rc, PH_STORE, 1);
}
int launch_fill(cudamat_solver *s, double *p, double v, int64_t cnt) {
    if (cnt <= 0) return CUDAMAT_OK;
    int grid = (int)((cnt + 1023) / 1024);
    if (grid > 148 * 16) grid = 148 * 16;
    k_fill<<<grid, 256, 0, s->stream>>>(p, v, cnt);
    s->launches++;
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

// ------------------------------------------------------------------------------------------
// analysis helpers
// ------------------------------------------------------------------------------------------
__global__ void k_row_stats(int n, const int *ia, int *out /*max_len, n_long, max_slab_nnz*/) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int mx = 0, nl = 0, ms = 0;
    if (i < n) {
        const int len = ia[i + 1] - ia[i];
        mx = len; nl = len > kLongRow ? 1 : 0;
        if ((i & 31) == 0) ms = ia[min(i + 32, n)] - ia[i];
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    ms = __reduce_max_sync(0xffffffffu, ms);
    nl = __reduce_add_sync(0xffffffffu, nl);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out + 0, mx);
        if (nl) atomicAdd(out + 1, nl);
        atomicMax(out + 2, ms);
    }
}

int launch_row_stats(cudamat_solver *s, int *h_out, double *mean) {
    int *d_out = nullptr;
    CM_CUDA(dev_alloc((void **)&d_out, 3 * sizeof(int)));
    CM_CUDA(cudaMemsetAsync(d_out, 0, 3 * sizeof(int), s->stream));
    if (s->n > 0) {
        k_row_stats<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->n, s->d_ia, d_out);
        s->launches++;
    }
    CM_CUDA(cudaMemcpyAsync(h_out, d_out, 3 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    dev_free(d_out);
    *mean = s->n > 0 ? (double)s->nnz / s->n : 0.0;
    return CUDAMAT_OK;
}

__global__ void k_sub_base(int *p, int64_t cnt, int base) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < cnt; i += stride) p[i] -= base;
}
int launch_normalize_base(cudaStream_t st, int *ia, int64_t n1, int *ja, int64_t nnz, int base) {
    if (base == 0) return CUDAMAT_OK;
    k_sub_base<<<1184, 256, 0, st>>>(ia, n1, base);
    k_sub_base<<<1184, 256, 0, st>>>(ja, nnz, base);
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

// first offending row (row pointers not monotone) / entry (column outside [0, ncols)) / entry whose column is not
// strictly larger than its predecessor's in the same row (unsorted or duplicate: the reference loader guarantees
// strictly ascending rows through verify_pattern, mmio_wrapper.h:123-126, and the ILU0 path relies on it), or INT_MAX
__global__ void k_validate_csr(const int *ia, int n, const int *ja, int64_t nnz, int64_t ncols, int *bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = i; k < nnz; k += stride) {
        const int c = ja[k];
        if (c < 0 || c >= ncols) atomicMin(bad + 1, (int)k);
    }
    for (int64_t r = i; r < n; r += stride) {
        const int s = ia[r], e = ia[r + 1];
        if (e < s) { atomicMin(bad + 0, (int)r); continue; }
        if (s < 0 || e > nnz) continue;                       // reported through the row-pointer end check of the caller
        for (int k = s + 1; k < e; ++k)
            if (ja[k] <= ja[k - 1]) { atomicMin(bad + 2, k); break; }
    }
}
int launch_validate_csr(cudaStream_t st, const int *ia, int n, const int *ja, int64_t nnz, int64_t ncols, int *h_bad) {
    int *d_bad = nullptr;
    CM_CUDA(dev_alloc((void **)&d_bad, 3 * sizeof(int)));
    CM_CUDA(cudaMemsetAsync(d_bad, 0x7f, 3 * sizeof(int), st));
    k_validate_csr<<<1184, 256, 0, st>>>(ia, n, ja, nnz, ncols, d_bad);
    CM_CUDA(cudaGetLastError());
    int h[3];
    CM_CUDA(cudaMemcpyAsync(h, d_bad, sizeof h, cudaMemcpyDeviceToHost, st));
    CM_CUDA(cudaStreamSynchronize(st));
    dev_free(d_bad);
    for (int q = 0; q < 3; ++q) h_bad[q] = h[q] == 0x7f7f7f7f ? -1 : h[q];
    return CUDAMAT_OK;
}

// ------------------------------------------------------------------------------------------
// generators: bit-identical to oracle/oracle.c orc_poisson3d / orc_xtrue / orc_random_dd
// ------------------------------------------------------------------------------------------
__global__ void k_poisson_count(int N, int64_t row0, int64_t row1, int *cnt) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q > row1 - row0) return;
    if (q == row1 - row0) { cnt[q] = 0; return; }
    const int64_t r = row0 + q, nn = (int64_t)N * N;
    const int i = (int)(r % N), j = (int)((r / N) % N), k = (int)(r / nn);
    cnt[q] = 1 + (k > 0) + (j > 0) + (i > 0) + (i < N - 1) + (j < N - 1) + (k < N - 1);
}
__global__ void k_poisson_fill(int N, int64_t row0, int64_t row1, const int *ia, int *ja, double *a) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= row1 - row0) return;
    const int64_t r = row0 + q, nn = (int64_t)N * N;
    const int i = (int)(r % N), j = (int)((r / N) % N), k = (int)(r / nn);
    int p = ia[q];
    if (k > 0)     { ja[p] = (int)(r - nn); a[p++] = -1.0; }
    if (j > 0)     { ja[p] = (int)(r - N);  a[p++] = -1.0; }
    if (i > 0)     { ja[p] = (int)(r - 1);  a[p++] = -1.0; }
    ja[p] = (int)r; a[p++] = 6.0;
    if (i < N - 1) { ja[p] = (int)(r + 1);  a[p++] = -1.0; }
    if (j < N - 1) { ja[p] = (int)(r + N);  a[p++] = -1.0; }
    if (k < N - 1) { ja[p] = (int)(r + nn); a[p++] = -1.0; }
}

int exclusive_scan_inplace(int *d, int64_t cnt, cudaStream_t st) {
    void *tmp = nullptr; size_t bytes = 0;
    CM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, d, d, cnt, st));
    CM_CUDA(dev_alloc(&tmp, bytes ? bytes : 16));
    CM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, d, d, cnt, st));
    CM_CUDA(cudaStreamSynchronize(st));
    dev_free(tmp);
    return CUDAMAT_OK;
}

int gen_poisson3d(int N, int64_t row0, int64_t row1, int *d_ia, int *d_ja, double *d_a, cudaStream_t st) {
    const int64_t m = row1 - row0;
    const int grid = (int)((m + 1 + 255) / 256);
    k_poisson_count<<<grid, 256, 0, st>>>(N, row0, row1, d_ia);
    CM_CUDA(cudaGetLastError());
    int rc = exclusive_scan_inplace(d_ia, m + 1, st);
    if (rc) return rc;
    if (d_ja && d_a && m > 0) {
        k_poisson_fill<<<(int)((m + 255) / 256), 256, 0, st>>>(N, row0, row1, d_ia, d_ja, d_a);
        CM_CUDA(cudaGetLastError());
    }
    return CUDAMAT_OK;
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash2(uint64_t seed, uint64_t a) {
    return mix64(seed + 0x9E3779B97F4A7C15ULL * (a + 1));
}
__host__ __device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
    return mix64(hash2(seed, a) + 0x9E3779B97F4A7C15ULL * (b + 1));
}
__device__ __forceinline__ double u01(uint64_t z) { return __dmul_rn((double)(z >> 11), 0x1.0p-53); }

__global__ void k_xtrue(uint64_t seed, int64_t i0, int64_t cnt, double *out) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; q < cnt; q += stride)
        out[q] = __dsub_rn(__dmul_rn(2.0, u01(hash2(seed, (uint64_t)(i0 + q)))), 1.0);
}
int gen_xtrue(uint64_t seed, int64_t i0, int64_t cnt, double *d_out, cudaStream_t st) {
    if (cnt <= 0) return CUDAMAT_OK;
    int grid = (int)((cnt + 255) / 256);
    if (grid > 148 * 32) grid = 148 * 32;
    k_xtrue<<<grid, 256, 0, st>>>(seed, i0, cnt, d_out);
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

// one row of the random diagonally-dominant generator; mirrors rdd_row() in oracle/oracle.c
__device__ int rdd_row(int n, uint64_t seed, int i, int *cols, double *vals) {
    const double uc = u01(hash3(seed, (uint64_t)i, 0));
    int k;
    if (uc < 0.90) k = __popcll(hash3(seed, (uint64_t)i, 1) & 0xFFFULL);
    else if (uc < 0.99) k = __popcll(hash3(seed, (uint64_t)i, 1) & 0xFFFFFFFFFFFFULL);
    else k = __popcll(hash3(seed, (uint64_t)i, 1)) + __popcll(hash3(seed, (uint64_t)i, 2)) + __popcll(hash3(seed, (uint64_t)i, 3));
    if (k > n - 1) k = n - 1;
    int m = 0;
    for (int s = 0; s < k; ++s) {
        const uint64_t hz = hash3(seed, (uint64_t)i, 16 + (uint64_t)s);
        int c = (int)__umul64hi(hz, (uint64_t)(n - 1));
        if (c >= i) c += 1;
        int pos = m;
        while (pos > 0 && cols[pos - 1] > c) --pos;
        if (pos > 0 && cols[pos - 1] == c) continue;
        for (int q = m; q > pos; --q) cols[q] = cols[q - 1];
        cols[pos] = c;
        ++m;
    }
    int pos = m;
    while (pos > 0 && cols[pos - 1] > i) --pos;
    for (int q = m; q > pos; --q) cols[q] = cols[q - 1];
    cols[pos] = i;
    ++m;
    if (vals) {
        double sum = 0.0;
        for (int q = 0; q < m; ++q) {
            if (cols[q] == i) continue;
            const double u = u01(hash3(seed, (uint64_t)i, 0x100000000ULL + (uint64_t)cols[q]));
            const double v = __dsub_rn(__dmul_rn(u, 20.0), 10.0);
            vals[q] = v;
            sum = __dadd_rn(sum, fabs(v));
        }
        const double ud = u01(hash3(seed, (uint64_t)i, 4));
        vals[pos] = __dadd_rn(sum, __dadd_rn(__dmul_rn(ud, 9.0), 1.0));
    }
    return m;
}
__global__ void k_rdd_count(int n, uint64_t seed, int *cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { cnt[n] = 0; return; }
    int cols[200];
    cnt[i] = rdd_row(n, seed, i, cols, nullptr);
}
__global__ void k_rdd_fill(int n, uint64_t seed, const int *ia, int *ja, double *a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cols[200]; double vals[200];
    const int m = rdd_row(n, seed, i, cols, vals);
    const int p = ia[i];
    for (int q = 0; q < m; ++q) { ja[p + q] = cols[q]; a[p + q] = vals[q]; }
}
int gen_random_dd(int n, uint64_t seed, int *d_ia, int *d_ja, double *d_a, int64_t *nnz_out, cudaStream_t st) {
    if (!d_ja) {
        k_rdd_count<<<(n + 1 + 127) / 128, 128, 0, st>>>(n, seed, d_ia);
        CM_CUDA(cudaGetLastError());
        int rc = exclusive_scan_inplace(d_ia, (int64_t)n + 1, st);
        if (rc) return rc;
        int last = 0;
        CM_CUDA(cudaMemcpyAsync(&last, d_ia + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        if (nnz_out) *nnz_out = last;
    } else {
        k_rdd_fill<<<(n + 127) / 128, 128, 0, st>>>(n, seed, d_ia, d_ja, d_a);
        CM_CUDA(cudaGetLastError());
        CM_CUDA(cudaStreamSynchronize(st));
    }
    return CUDAMAT_OK;
}

}  // namespace cudamat
