// stream.cu — SpMV variant STREAM for IRREGULAR matrices (row lengths that differ a lot inside a warp's 32 rows):
// replaces cusparseDcsrmv (pbicgstab.cu:67,104,132,646,676,704) + the cublasDdot that follows it, like every SpMV variant.
//
// The row-per-lane kernel (ROWLANE) walks every lane through the LONGEST row of its slab: with the 90 / 9 / 1 % length
// mixture of BASELINE config 4 (generator.cpp-style, mean 9.5 entries) ~95 % of the slabs hold a 24-entry row, so most
// lanes idle through most gather rounds, and a > 32-entry row serialises the whole warp.
// Here the expensive part — the (col, val) stream and the random x gathers — is done ENTRY-parallel:
//   1. a warp takes the contiguous entry range of its 32 rows in chunks of 256 entries; lane l loads entries
//      l, l + 32, ... (fully coalesced), gathers x[col] for its 8 entries (8 independent gathers in flight per lane, whatever
//      the row lengths) and parks (val, x) in the warp's shared-memory chunk;
//   2. every lane then runs ITS row's FMA chain over the part of the row that lies in the chunk, out of shared memory, in
//      storage order (rows <= 32 entries); rows > 32 entries are summed by the whole warp with the spec's 32 interleaved
//      lane chains + butterfly.  Chains continue across chunks in registers.
// Row sums are therefore the spec's bit for bit: same operations, same order as ROWLANE / the oracle.
#include "solver.h"
#include <algorithm>
#include <vector>

namespace cudamat {

constexpr int kStreamChunk = 256;                                  // entries per warp chunk (8 per lane)
constexpr int kStreamPerLane = kStreamChunk / 32;

// how the random x gathers are issued (tuning builds: -DCUDAMAT_GATHER_MODE=n)
#ifndef CUDAMAT_GATHER_MODE
#define CUDAMAT_GATHER_MODE 0
#endif
__device__ __forceinline__ double gather_x(const double *p) {
#if CUDAMAT_GATHER_MODE == 1
    return __ldcg(p);
#elif CUDAMAT_GATHER_MODE == 2
    return __ldcs(p);
#elif CUDAMAT_GATHER_MODE == 3
    double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
#elif CUDAMAT_GATHER_MODE == 4
    return __ldcv(p);
#elif CUDAMAT_GATHER_MODE == 5
    double v; asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
#else
    return __ldg(p);
#endif
}

// PASS: 0 = the whole matrix in one launch; 1 / 2 / 3 = first / middle / last launch of the column-blocked form (below): the row
// chain continues from y (first pass: from +0.0, or from y for rows flagged in long_bits, which the long-row kernel has already
// summed), HAS_D and the fused dots belong to the last pass only.
template <bool HAS_D, int NDOT, int PASS>
__global__ void __launch_bounds__(kCtaThreads, 3) k_spmv_stream(const SpmvArgs a, const unsigned *__restrict__ long_bits) {
    extern __shared__ __align__(16) double s_buf[];                // [warp][val 256 | x 256]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_base = blockIdx.x * kTile;
    pdl_sync();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    halo_wait(a.hw, blockIdx.x, a.sc ? &a.sc->status : nullptr);
    double *sv = s_buf + warp * 2 * kStreamChunk, *sx = sv + kStreamChunk;
#pragma unroll 1
    for (int j = 0; j < kSlabsPerWarp; ++j) {
        const int slab = j * kCtaWarps + warp;
        const int row0 = row_base + slab * kSlab;
        if (row0 >= a.n) break;                                    // warp-uniform
        const int row = row0 + lane;
        const bool active = row < a.n;
        int rs = 0, re = 0;
        if (active) { rs = __ldg(a.ia + row); re = __ldg(a.ia + row + 1); }
        double uval = 0.0;
        if (NDOT >= 1 && (PASS == 0 || PASS == 3) && active) uval = __ldg(a.u + row);
        const int lo_all = __shfl_sync(0xffffffffu, rs, 0);
        // entries of the slab end where its last existing row ends
        int hi_all = __reduce_max_sync(0xffffffffu, re);
        const bool is_long = (re - rs) > kLongRow;
        const unsigned long_mask = __ballot_sync(0xffffffffu, is_long);
        double acc = 0.0;                                          // this lane's row (<= 32 entries): sequential chain
        if (PASS == 1) { if (active && ((__ldg(long_bits + (row0 >> 5)) >> lane) & 1u)) acc = a.y[row]; }
        if (PASS >= 2) { if (active) acc = __ldcs(a.y + row); }
        double lacc = 0.0;                                         // interleaved lane chain of the long row in progress
#pragma unroll 1
        for (int lo = lo_all; lo < hi_all; lo += kStreamChunk) {
            const int hi = min(lo + kStreamChunk, hi_all);
            // ---- 1. entry-parallel: coalesced (col, val), independent gathers, park (val, x) ----
            int cj[kStreamPerLane];
            double av[kStreamPerLane];
#pragma unroll
            for (int q = 0; q < kStreamPerLane; ++q) {
                const int e = lo + lane + 32 * q;
                const bool p = e < hi;
                // the matrix streams past once: evict-first, so that the x entries the gathers reuse stay in L2
                cj[q] = p ? __ldcs(a.ja + e) : -1;
                av[q] = p ? __ldcs(a.val + e) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < kStreamPerLane; ++q) {
                const double xv = cj[q] >= 0 ? gather_x(a.x + cj[q]) : 0.0;
                sv[lane + 32 * q] = av[q];
                sx[lane + 32 * q] = xv;
            }
            __syncwarp();
            // ---- 2a. rows of <= 32 entries: the lane's own chain over [max(rs, lo), min(re, hi)) ----
            if (!is_long) {
                const int b = max(rs, lo) - lo, e = min(re, hi) - lo;
                for (int k = b; k < e; ++k) acc = __fma_rn(sv[k], sx[k], acc);
            }
            // ---- 2b. rows of > 32 entries: 32 interleaved lane chains (lane l takes k == l mod 32 from the row start) ----
            unsigned lm = long_mask;
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                const int ss = __shfl_sync(0xffffffffu, rs, src), ee = __shfl_sync(0xffffffffu, re, src);
                if (ee <= lo || ss >= hi) continue;                // the row has no entry in this chunk (warp-uniform)
                if (ss >= lo) lacc = 0.0;                          // the row starts in this chunk
                const int b = max(ss, lo), e = min(ee, hi);
                // first k >= b with (k - ss) % 32 == lane
                int k = b + ((lane - (b - ss)) & 31);
                for (; k < e; k += 32) lacc = __fma_rn(sv[k - lo], sx[k - lo], lacc);
                if (ee <= hi) {                                    // the row ends in this chunk: butterfly, owner takes it
                    const double tot = warp_butterfly(lacc);
                    if (lane == src) acc = tot;
                }
            }
            __syncwarp();                                          // the chunk buffer is free again
        }
        double sum = acc;
        if (PASS == 1 || PASS == 2) {                              // partial chain: parked in y until the next column block
            if (active) __stcs(a.y + row, sum);
            continue;
        }
        if (HAS_D) { if (active) sum = __dadd_rn(sum, __dmul_rn(__ldg(a.d + row), __ldg(a.x + row))); }
        if (active) a.y[row] = sum;
        if (NDOT >= 1) slab_deposit(a.rc, 0, blockIdx.x * kTileSlabs + slab, active ? __dmul_rn(sum, uval) : 0.0, lane);
        if (NDOT >= 2) slab_deposit(a.rc, 1, blockIdx.x * kTileSlabs + slab, active ? __dmul_rn(sum, sum) : 0.0, lane);
    }
}

// ---- column-blocked form -------------------------------------------------------------------------------------------
// When x is much larger than the L2 (BASELINE config 4: 50 M rows, x = 400 MB, columns uniform) every gather of a plain
// pass misses the L2 and costs ~96 B of DRAM traffic for 8 useful bytes (ncu, profiles/r2b: 50.8 GB per SpMV for 21 GB of
// sectors asked for).  The analysis therefore keeps a second copy of the matrix split into K column blocks whose x range
// (<= 64 MB) stays L2-resident while its block streams past: block b holds, for every row, the entries whose column lies
// in block b, in storage order.  One launch per block; a row's FMA chain is carried from block to block through y, so the
// operations and their order are exactly those of the one-pass kernel (bit-identical).  Rows of > 32 entries (spec: 32
// interleaved lane chains over the WHOLE row) are left out of the blocks and summed by a warp-per-row kernel up front.
struct StreamBlocks {
    int K = 0;
    int *ia = nullptr;                 // [K][n + 1] row pointers inside block b
    int *ja = nullptr; double *val = nullptr;
    std::vector<int64_t> off;          // start of block b in ja / val
    unsigned *long_bits = nullptr;     // bit r: row r has > 32 entries
    int *long_rows = nullptr; int n_long = 0;
};

__global__ void k_sblk_count(int n, const int *ia, const int *ja, int bw, int K, int *cnt, unsigned *long_bits, int *long_rows, int *n_long) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = r < n;
    int s = 0, e = 0;
    if (act) { s = ia[r]; e = ia[r + 1]; }
    const bool lng = act && (e - s) > kLongRow;
    const unsigned m = __ballot_sync(0xffffffffu, lng);
    if ((threadIdx.x & 31) == 0 && r < n) long_bits[r >> 5] = m;
    if (lng) long_rows[atomicAdd(n_long, 1)] = r;
    if (!act) return;
    int k = s;
    for (int b = 0; b < K; ++b) {
        int c = 0;
        if (!lng) { const int lim = (b + 1) * bw; while (k < e && ja[k] < lim) { ++k; ++c; } }
        cnt[(size_t)b * (n + 1) + r] = c;
    }
}
__global__ void k_sblk_scatter(int n, const int *ia, const int *ja, const double *val, int bw, int K, const int *bia,
                               const long long *boff, int *bja, double *bval) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int s = ia[r], e = ia[r + 1];
    if (e - s > kLongRow) return;
    int k = s;
    for (int b = 0; b < K; ++b) {
        const int lim = (b + 1) * bw;
        long long o = boff[b] + bia[(size_t)b * (n + 1) + r];
        while (k < e && ja[k] < lim) { bja[o] = ja[k]; bval[o] = val[k]; ++k; ++o; }
    }
}
// warp per long row: the spec's 32 interleaved lane chains + butterfly (== rowsum_long of the one-pass kernels)
__global__ void __launch_bounds__(kCtaThreads) k_spmv_longrows(const SpmvArgs a, const int *long_rows, int n_long) {
    pdl_sync();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    const int lane = threadIdx.x & 31;
    for (int w = blockIdx.x * kCtaWarps + (threadIdx.x >> 5); w < n_long; w += gridDim.x * kCtaWarps) {
        const int row = __ldg(long_rows + w);
        const int s = __ldg(a.ia + row), e = __ldg(a.ia + row + 1);
        double acc = 0.0;
        for (int k = s + lane; k < e; k += 32) acc = __fma_rn(__ldg(a.val + k), __ldg(a.x + __ldg(a.ja + k)), acc);
        acc = warp_butterfly(acc);
        if (lane == 0) a.y[row] = acc;
    }
}

// One launch of the column-blocked form: a block holds only a few entries per row, so a warp takes 128 CONSECUTIVE rows (4
// slabs; lane l owns rows l, l + 32, l + 64, l + 96 of them) and streams their contiguous entry range through the same
// entry-parallel chunk as above — 4 x the work per dependent memory round trip.  No row of a block has > 32 entries.
template <bool HAS_D, int NDOT, int PASS>
__global__ void __launch_bounds__(kCtaThreads, 2) k_spmv_stream_blk(const SpmvArgs a, const unsigned *__restrict__ long_bits) {
    extern __shared__ __align__(16) double s_buf[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kTile + warp * 128;
    pdl_sync();
    if (a.check_status && a.sc->status != ST_RUNNING) return;
    if (row0 >= a.n) return;
    double *sv = s_buf + warp * 2 * kStreamChunk, *sx = sv + kStreamChunk;
    int rs[4], re[4];
    double acc[4];
    bool act[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int row = row0 + q * 32 + lane;
        act[q] = row < a.n;
        rs[q] = act[q] ? __ldg(a.ia + row) : 0;
        re[q] = act[q] ? __ldg(a.ia + row + 1) : 0;
        acc[q] = 0.0;
        if (PASS == 1) { if (act[q] && ((__ldg(long_bits + ((row0 + q * 32) >> 5)) >> lane) & 1u)) acc[q] = a.y[row]; }
        if (PASS >= 2) { if (act[q]) acc[q] = __ldcs(a.y + row); }
    }
    const int lo_all = __shfl_sync(0xffffffffu, rs[0], 0);
    const int hi_all = __reduce_max_sync(0xffffffffu, max(max(re[0], re[1]), max(re[2], re[3])));
#pragma unroll 1
    for (int lo = lo_all; lo < hi_all; lo += kStreamChunk) {
        const int hi = min(lo + kStreamChunk, hi_all);
        int cj[kStreamPerLane];
        double av[kStreamPerLane];
#pragma unroll
        for (int q = 0; q < kStreamPerLane; ++q) {
            const int e = lo + lane + 32 * q;
            const bool p = e < hi;
            cj[q] = p ? __ldcs(a.ja + e) : -1;
            av[q] = p ? __ldcs(a.val + e) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < kStreamPerLane; ++q) {
            const double xv = cj[q] >= 0 ? gather_x(a.x + cj[q]) : 0.0;
            sv[lane + 32 * q] = av[q];
            sx[lane + 32 * q] = xv;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b = max(rs[q], lo) - lo, e = min(re[q], hi) - lo;
            for (int k = b; k < e; ++k) acc[q] = __fma_rn(sv[k], sx[k], acc[q]);
        }
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int row = row0 + q * 32 + lane;
        if (row0 + q * 32 >= a.n) break;                           // warp-uniform
        double sum = acc[q];
        if (PASS == 1 || PASS == 2) { if (act[q]) __stcs(a.y + row, sum); continue; }
        if (HAS_D) { if (act[q]) sum = __dadd_rn(sum, __dmul_rn(__ldg(a.d + row), __ldg(a.x + row))); }
        if (act[q]) a.y[row] = sum;
        const int slab = blockIdx.x * kTileSlabs + warp * 4 + q;
        if (NDOT >= 1) slab_deposit(a.rc, 0, slab, act[q] ? __dmul_rn(sum, __ldg(a.u + row)) : 0.0, lane);
        if (NDOT >= 2) slab_deposit(a.rc, 1, slab, act[q] ? __dmul_rn(sum, sum) : 0.0, lane);
    }
}

void stream_release(cudamat_solver *s) {
    StreamBlocks *B = s->sblk;
    if (!B) return;
    dev_free(B->ia); dev_free(B->ja); dev_free(B->val); dev_free(B->long_bits); dev_free(B->long_rows);
    delete B;
    s->sblk = nullptr;
}

int exclusive_scan_inplace(int *d, int64_t cnt, cudaStream_t st);      // kernels.cu

// builds the column-blocked copy when x does not fit the L2 (or option "stream_blocks" asks for K blocks)
int stream_plan(cudamat_solver *s) {
    stream_release(s);
    if (s->comm || s->n <= 0 || s->nnz <= 0) return CUDAMAT_OK;
    const int n = s->n;
    const size_t xbytes = sizeof(double) * (size_t)n;
    int K = s->opt_stream_blocks;
    if (K == 0) K = xbytes <= (size_t)96 << 20 ? 1 : (int)std::min<size_t>(32, (xbytes + ((size_t)64 << 20) - 1) / ((size_t)64 << 20));
    if (K <= 1) return CUDAMAT_OK;
    K = std::min(K, 32);
    const int bw = ((n + K - 1) / K + 31) / 32 * 32;
    StreamBlocks *B = new StreamBlocks();
    s->sblk = B;
    B->K = K;
    int *d_nl = nullptr; long long *d_off = nullptr;
    CM_CUDA(dev_alloc((void **)&B->ia, sizeof(int) * (size_t)K * (n + 1)));
    CM_CUDA(dev_alloc((void **)&B->long_bits, sizeof(unsigned) * (size_t)((n + 31) / 32 + 1)));
    CM_CUDA(dev_alloc((void **)&B->long_rows, sizeof(int) * (size_t)std::max(s->n_long_rows, 1)));
    CM_CUDA(dev_alloc((void **)&d_nl, sizeof(int)));
    CM_CUDA(dev_alloc((void **)&d_off, sizeof(long long) * 33));
    CM_CUDA(cudaMemsetAsync(d_nl, 0, sizeof(int), s->stream));
    CM_CUDA(cudaMemsetAsync(B->ia, 0, sizeof(int) * (size_t)K * (n + 1), s->stream));
    k_sblk_count<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->d_ia, s->d_ja, bw, K, B->ia, B->long_bits, B->long_rows, d_nl);
    CM_CUDA(cudaGetLastError());
    s->launches++;
    B->off.assign((size_t)K + 1, 0);
    for (int b = 0; b < K; ++b) {
        int rc = exclusive_scan_inplace(B->ia + (size_t)b * (n + 1), (int64_t)n + 1, s->stream);
        if (rc) return rc;
        int tot = 0;
        CM_CUDA(cudaMemcpyAsync(&tot, B->ia + (size_t)b * (n + 1) + n, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
        B->off[b + 1] = B->off[b] + tot;
    }
    CM_CUDA(cudaMemcpyAsync(&B->n_long, d_nl, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    long long hoff[33] = {0};
    for (int b = 0; b <= K; ++b) hoff[b] = B->off[b];
    CM_CUDA(cudaMemcpyAsync(d_off, hoff, sizeof hoff, cudaMemcpyHostToDevice, s->stream));
    const size_t tot = (size_t)std::max<int64_t>(B->off[K], 1);
    CM_CUDA(dev_alloc((void **)&B->ja, sizeof(int) * tot + 16));
    CM_CUDA(dev_alloc((void **)&B->val, sizeof(double) * tot + 16));
    k_sblk_scatter<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->d_ia, s->d_ja, s->d_a, bw, K, B->ia, d_off, B->ja, B->val);
    CM_CUDA(cudaGetLastError());
    CM_CUDA(cudaStreamSynchronize(s->stream));
    s->launches++;
    dev_free(d_nl); dev_free(d_off);
    return CUDAMAT_OK;
}

template <bool HAS_D, int NDOT, int PASS>
static int launch_stream_k(cudamat_solver *s, const SpmvArgs &a, const unsigned *long_bits) {
    const int grid = (a.n + kTile - 1) / kTile;
    if (grid == 0) return CUDAMAT_OK;
    cudaLaunchConfig_t cfg{};
    constexpr size_t smem = sizeof(double) * 2 * kStreamChunk * kCtaWarps;          // 64 KB: 3 CTAs per SM
    static bool attr_set[64] = {};
    const int dv = s->device & 63;
    const void *kern = PASS == 0 ? (const void *)k_spmv_stream<HAS_D, NDOT, PASS> : (const void *)k_spmv_stream_blk<HAS_D, NDOT, PASS>;
    if (!attr_set[dv]) { CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set[dv] = true; }
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kCtaThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    void *args[] = {(void *)&a, (void *)&long_bits};
    CM_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
    s->launches++;
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

template <bool HAS_D, int NDOT>
static int launch_stream_t(cudamat_solver *s, const SpmvArgs &a) {
    const StreamBlocks *B = s->sblk;
    if (!B || B->K <= 1 || a.ia != s->d_ia) return launch_stream_k<HAS_D, NDOT, 0>(s, a, nullptr);
    int rc;
    if (B->n_long > 0) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)std::min((B->n_long + kCtaWarps - 1) / kCtaWarps, 148 * 8)); cfg.blockDim = dim3(kCtaThreads); cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        CM_CUDA(cudaLaunchKernelEx(&cfg, k_spmv_longrows, a, (const int *)B->long_rows, B->n_long));
        s->launches++;
    }
    for (int b = 0; b < B->K; ++b) {
        SpmvArgs p = a;
        p.ia = B->ia + (size_t)b * (a.n + 1); p.ja = B->ja + B->off[b]; p.val = B->val + B->off[b];
        if (b > 0) { p.check_status = a.check_status; p.hw = HaloWait{}; }
        if (b == 0) rc = launch_stream_k<false, 0, 1>(s, p, B->long_bits);
        else if (b + 1 < B->K) rc = launch_stream_k<false, 0, 2>(s, p, B->long_bits);
        else rc = launch_stream_k<HAS_D, NDOT, 3>(s, p, B->long_bits);
        if (rc) return rc;
    }
    return CUDAMAT_OK;
}

int launch_stream_spmv(cudamat_solver *s, const SpmvArgs &a) {
    const bool hd = a.d != nullptr;
    switch (a.ndot) {
    case 0: return hd ? launch_stream_t<true, 0>(s, a) : launch_stream_t<false, 0>(s, a);
    case 1: return hd ? launch_stream_t<true, 1>(s, a) : launch_stream_t<false, 1>(s, a);
    default: return hd ? launch_stream_t<true, 2>(s, a) : launch_stream_t<false, 2>(s, a);
    }
}

}  // namespace cudamat
