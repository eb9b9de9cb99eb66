// ilu0.cu — ILU(0) preconditioner: level analysis, level-scheduled factorisation and the
// forward/backward triangular sweeps.  Replaces cusparseDcsrsv_analysis (pbicgstab.cu:338,345),
// cusparseDcsrilu0 (:359) and cusparseDcsrsv_solve (:94,98,123,127).
//
// Arithmetic (bit-identical to oracle/oracle.c orc_ilu0 / orc_sptrsv_*): row-wise IKJ elimination
// with ascending k, updates restricted to the row's own pattern, l = a_ik / a_kk (IEEE division),
// a_ij = fma(-l, a_kj, a_ij); L sweep acc = fma(-L_ik, y_k, acc) over ascending k from acc = rhs_i;
// U sweep the same over ascending upper columns, then one IEEE division by the diagonal.
// The parallel schedule (levels / sync-free flags) never changes the per-row operation order.
#include "solver.h"
#include "rowfuncs.cuh"
#include <algorithm>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

namespace cudamat {

// ------------------------------------------------------------------------------------------
// factorisation: one launch per level, one thread per row of the level
// ------------------------------------------------------------------------------------------
__global__ void k_ilu0_level(const int *order, int cnt, const int *ia, const int *ja, const int *diag,
                             double *M, int *zero_pivot) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const int i = order[t];
    if (i < 0) return;
    const int rs = ia[i], re = ia[i + 1];
    for (int p = rs; p < re; ++p) {
        const int k = ja[p];
        if (k >= i) break;
        const double piv = M[diag[k]];
        const double l = __ddiv_rn(M[p], piv);
        M[p] = l;
        int q = p + 1;
        const int ke = ia[k + 1];
        for (int pk = diag[k] + 1; pk < ke; ++pk) {
            const int j = ja[pk];
            while (q < re && ja[q] < j) ++q;
            if (q >= re) break;
            if (ja[q] == j) M[q] = __fma_rn(-l, M[pk], M[q]);
        }
    }
    if (M[diag[i]] == 0.0) atomicMin(zero_pivot, i);
}

// ------------------------------------------------------------------------------------------
// triangular sweeps, variant A: one launch per level (simple, used for small level counts and as
// the cross-check of variant B)
// ------------------------------------------------------------------------------------------
template <bool UPPER>
__global__ void k_sptrsv_level(const int *order, int cnt, const int *ia, const int *ja, const int *diag,
                               const double *M, const double *rhs, double *out, const int *status) {
    if (status && *status != ST_RUNNING) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const int i = order[t];
    if (i < 0) return;
    double acc = rhs[i];
    if (!UPPER) {
        const int pe = diag[i];
        for (int p = ia[i]; p < pe; ++p) acc = __fma_rn(-M[p], out[ja[p]], acc);
        out[i] = acc;
    } else {
        const int pd = diag[i], pe = ia[i + 1];
        for (int p = pd + 1; p < pe; ++p) acc = __fma_rn(-M[p], out[ja[p]], acc);
        out[i] = __ddiv_rn(acc, M[pd]);
    }
}

// ------------------------------------------------------------------------------------------
// triangular sweeps, variant B: ONE launch per sweep, sync-free.
//   * rows are laid out level by level (each level padded to a warp multiple so a warp never mixes
//     levels); CTAs take a ticket so that logical CTA order == start order, which makes "wait for an
//     earlier row" deadlock-free;
//   * the output vector itself carries the "ready" signal: it is pre-filled with a quiet-NaN sentinel
//     of a payload no arithmetic result can have, a dependent row spins (one L2 round trip per poll)
//     until the cell differs from the sentinel — a single 8-byte store publishes a row, no flag array,
//     no fence;
//   * the row's matrix entries are loaded before the first wait, so the dependent chain per level is
//     just poll -> fma -> store.
//   The output vector is armed (filled with the sentinel) by a coalesced fill right before the sweep (sptrsv_arm).
// ------------------------------------------------------------------------------------------
constexpr unsigned long long kSentinelBits = 0xFFF8B200C0DEFACEull;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_f64(double *p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double wait_value(const double *cell) {
    unsigned long long v = ld_relaxed_u64(cell);
    unsigned spins = 0;
    while (v == kSentinelBits) {
        if (++spins > (1u << 22)) __trap();            // never hang the GPU on a broken schedule
        __nanosleep(spins < 4 ? 64 : 256);             // back off: pollers must not saturate L2
        v = ld_relaxed_u64(cell);
    }
    return __longlong_as_double((long long)v);
}

// Sweep plan: the rows of one triangular factor re-laid out in LEVEL ORDER so that every load of the sweep is
// coalesced: position t holds row order[t], its number of off-diagonal entries, the first kPlanW of them
// (column, value) in struct-of-arrays form, the CSR position of the rest, and (U) the diagonal value.
// Built once after the factorisation (k_build_plan); values are bit copies of M, so arithmetic is unchanged.
constexpr int kPlanW = 4;

template <bool UPPER>
__global__ void k_build_plan(const int *order, int len, const int *ia, const int *ja, const int *diag, const double *M,
                             int *p_cnt, int *p_ptr, int *p_col, double *p_val, double *p_dg) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    const int i = order[t];
    int p = 0, pe = 0;
    if (i >= 0) { const int pd = diag[i]; if (!UPPER) { p = ia[i]; pe = pd; } else { p = pd + 1; pe = ia[i + 1]; p_dg[t] = M[pd]; } }
    else if (UPPER) p_dg[t] = 1.0;
    p_cnt[t] = pe - p;
    p_ptr[t] = p;
#pragma unroll
    for (int q = 0; q < kPlanW; ++q) {
        const bool has = (p + q) < pe;
        p_col[(size_t)q * len + t] = has ? ja[p + q] : 0;
        p_val[(size_t)q * len + t] = has ? M[p + q] : 0.0;
    }
}

// Persistent schedule: the grid is sized to what is co-resident (occupancy API) and CTA b processes the
// 256-row chunks b, b+G, b+2G, ... of the level-ordered plan in increasing order.  The smallest unfinished
// chunk only depends on finished chunks and its owner is (or becomes) resident, so the schedule is
// deadlock-free without tickets, and the look-ahead is bounded by G*256 rows.
// Waiting is two-phase so that pollers do not saturate L2: one representative lane per warp spins on its last
// dependency (a warp never mixes levels, so its rows become ready together; one request per L2 round trip and
// warp — __nanosleep back-off measured 3 % slower, profiles/r1b_sptrsv_experiments.md), then every lane verifies
// its own dependencies, all loads of a round issued together.
template <bool UPPER>
__global__ void __launch_bounds__(256) k_sptrsv_syncfree(const int *order, int len, const int *p_cnt, const int *p_ptr,
                                                        const int *p_col, const double *p_val, const double *p_dg,
                                                        const int *ja, const double *M, const double *rhs,
                                                        double *out, const int *status) {
    pdl_prologue();
    if (status && *status != ST_RUNNING) return;
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * 256 + threadIdx.x; t - lane < len; t += gridDim.x * 256) {       // warp-uniform trip count
        const bool inb = t < len;
        const int i = inb ? order[t] : -1;
        const bool act = i >= 0;
        const int cnt = act ? p_cnt[t] : 0;
        int c[kPlanW]; double m[kPlanW];
#pragma unroll
        for (int q = 0; q < kPlanW; ++q) {
            c[q] = (q < cnt) ? p_col[(size_t)q * len + t] : -1;
            m[q] = (q < cnt) ? p_val[(size_t)q * len + t] : 0.0;
        }
        const double dg = (UPPER && act) ? p_dg[t] : 1.0;
        double acc = act ? rhs[i] : 0.0;
        // phase 1: the first lane that has a dependency waits for its last one
        const unsigned have = __ballot_sync(0xffffffffu, cnt > 0);
        if (have) {
            const int rep = __ffs(have) - 1;
            if (lane == rep) {
                const int last = (cnt <= kPlanW) ? c[cnt - 1] : ja[p_ptr[t] + cnt - 1];
                unsigned spins = 0;
                while (ld_relaxed_u64(out + last) == kSentinelBits) {
                    if (++spins > (1u << 24)) __trap();          // never hang the GPU on a broken schedule
                }
            }
            __syncwarp();
        }
        // phase 2: every lane waits for all of its first kPlanW dependencies, one round = all loads in flight
        unsigned long long v[kPlanW];
        unsigned spins = 0;
        for (;;) {
            bool ready = true;
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) v[q] = (c[q] >= 0) ? ld_relaxed_u64(out + c[q]) : 0ull;
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) ready = ready && (v[q] != kSentinelBits);
            if (__all_sync(0xffffffffu, ready)) break;
            if (++spins > (1u << 24)) __trap();
        }
#pragma unroll
        for (int q = 0; q < kPlanW; ++q)
            if (q < cnt) acc = __fma_rn(-m[q], __longlong_as_double((long long)v[q]), acc);
        if (cnt > kPlanW) {                                       // rare: the rest of a long row straight from CSR
            const int p0 = p_ptr[t];
            for (int pp = p0 + kPlanW; pp < p0 + cnt; ++pp) acc = __fma_rn(-M[pp], wait_value(out + ja[pp]), acc);
        }
        if (UPPER) acc = __ddiv_rn(acc, dg);
        if (act) st_relaxed_f64(out + i, acc);
    }
}

// ------------------------------------------------------------------------------------------
// triangular sweeps, variant C: small systems (the output vector fits in shared memory, e.g. mat900 / mat10000:
// 88 / 199 levels of <= 15 / 100 rows).  ONE CTA walks the levels of the level-ordered plan with a CTA barrier
// between levels; solved values are read from shared memory, the operands of the next level are fetched before the
// barrier.  A level costs a barrier + a few shared-memory reads (~0.1 us) instead of an L2 round trip per
// dependent hop (~1.1 us) or a kernel launch.  Same per-row operation order as the other variants.
// ------------------------------------------------------------------------------------------
#ifndef CUDAMAT_SMEM_SWEEP_THREADS
#define CUDAMAT_SMEM_SWEEP_THREADS 1024
#endif
constexpr int kSmemSweepThreads = CUDAMAT_SMEM_SWEEP_THREADS;
template <bool UPPER>
__global__ void __launch_bounds__(kSmemSweepThreads) k_sptrsv_smem(const int *order, const int *level_ptr, int nlevels, int len, int n,
                                                                   const int *p_cnt, const int *p_ptr, const int *p_col, const double *p_val,
                                                                   const double *p_dg, const int *ja, const double *M,
                                                                   const double *rhs, double *out, const int *status) {
    extern __shared__ double y[];                                  // n entries
    pdl_prologue();
    if (status && *status != ST_RUNNING) return;
    const int tid = threadIdx.x;
    struct Row { int i, cnt, c[kPlanW]; double m[kPlanW], dg, rhs; };
    auto fetch = [&](int t, Row &r) {
        r.i = (t >= 0) ? order[t] : -1;
        r.cnt = 0; r.dg = 1.0; r.rhs = 0.0;
        if (r.i >= 0) {
            r.cnt = p_cnt[t];
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) { r.c[q] = p_col[(size_t)q * len + t]; r.m[q] = p_val[(size_t)q * len + t]; }
            if (UPPER) r.dg = p_dg[t];
            r.rhs = rhs[r.i];
        }
    };
    auto solve = [&](int t, const Row &r) {
        if (r.i < 0) return;
        double acc = r.rhs;
#pragma unroll
        for (int q = 0; q < kPlanW; ++q)
            if (q < r.cnt) acc = __fma_rn(-r.m[q], y[r.c[q]], acc);
        if (r.cnt > kPlanW) {
            const int p0 = p_ptr[t];
            for (int pp = p0 + kPlanW; pp < p0 + r.cnt; ++pp) acc = __fma_rn(-M[pp], y[ja[pp]], acc);
        }
        if (UPPER) acc = __ddiv_rn(acc, r.dg);
        y[r.i] = acc;
        out[r.i] = acc;
    };
    // The CTA is split into kGroups groups of 128 threads; group g owns the levels g, g + kGroups, ... and fetches
    // the operands of its next level a whole round (kGroups levels) before they are needed, so neither the plan
    // loads nor the dependent rhs[order[t]] load sit on the level-to-level chain.
    constexpr int kGroups = kSmemSweepThreads / 128;
    const int g = tid >> 7, lt = tid & 127;
    int tcur = -1, tnxt = -1;
    Row cur, nxt;
    auto fetch_level = [&](int l, Row &r, int &t) {
        t = -1;
        if (l < nlevels) { const int tt = level_ptr[l] + lt; if (tt < level_ptr[l + 1]) t = tt; }
        fetch(t, r);
    };
    fetch_level(g, cur, tcur);
    for (int l0 = 0; l0 < nlevels; l0 += kGroups) {
        fetch_level(l0 + kGroups + g, nxt, tnxt);
#pragma unroll 1
        for (int k = 0; k < kGroups; ++k) {
            const int l = l0 + k;
            if (l >= nlevels) break;                               // uniform
            if (k == g) {
                solve(tcur, cur);
                for (int t = level_ptr[l] + lt + 128; t < level_ptr[l + 1]; t += 128) { Row r; fetch(t, r); solve(t, r); }   // wide levels
            }
            __syncthreads();
        }
        cur = nxt; tcur = tnxt;
    }
}

// ------------------------------------------------------------------------------------------
// Small systems, variant RING: the level-to-level chain of k_sptrsv_smem is ~490 cycles per level — a 32-warp barrier and,
// every 8 levels, a stall on the operand prefetch (the loads' consumers sit in the same in-order warps as the chain).  Here
// the CTA is split by role: ONE group of 128 threads walks the levels with a 4-warp named barrier, three loader groups stream
// the plan (columns, factor values, diagonal and the reciprocal half of its division, rhs[order[t]]) into a ring of
// shared-memory slots of 128 rows, two mbarriers per slot (full / empty).  The walker's chain per level: full-wait -> slot row
// from shared memory -> y[c] from shared memory -> <= 4 dependent DFMA (-> division) -> y[i] -> barrier.  Levels wider than 128
// rows are consecutive slots with the barrier only after the last.  Same per-row operation order as every other sweep.
// ------------------------------------------------------------------------------------------
#ifndef CUDAMAT_RING_LOADERS
#define CUDAMAT_RING_LOADERS 3                                   // 2-4 measured equal (mat10000 268-275 us / iteration), 7: 296, 1: 379
#endif
constexpr int kRingSlots = 8, kRingBatch = 4, kRingRows = 128, kRingLoaders = CUDAMAT_RING_LOADERS;
static_assert(kRingSlots == 2 * kRingBatch, "two batches");
constexpr int kRingThreads = kRingRows * (1 + kRingLoaders);
struct RingSlot {
    double m[kPlanW][kRingRows];
    double dg[kRingRows], rcp[kRingRows], rhs[kRingRows];
    int c[kPlanW][kRingRows];
    int i[kRingRows], cnt[kRingRows];
    int rows, last, beg, pad;
};
constexpr size_t kRingBytes = sizeof(RingSlot) * kRingSlots;

template <bool BACKOFF>
__device__ __forceinline__ void ring_wait(unsigned long long *bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok) {
            if (++spins > (1u << 22)) __trap();                    // never hang on a broken ring
            if (BACKOFF) __nanosleep(256);                         // loaders: 28 polling warps would starve the walker's LDS / barrier traffic
        }
    }
}
__device__ __forceinline__ void ring_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// entries kPlanW .. cnt - 1 of a long row, straight from CSR (kept out of the walker's loop body)
static __device__ __noinline__ double ring_long_row_tail(double acc, int p0, int cnt, const double *M, const int *ja, const double *y) {
    for (int pp = p0 + kPlanW; pp < p0 + cnt; ++pp) acc = __fma_rn(-M[pp], y[ja[pp]], acc);
    return acc;
}

template <bool UPPER>
__global__ void __launch_bounds__(kRingThreads, 1) k_sptrsv_ring(const int *order, const int *chunk_beg, const int *chunk_info, int nchunks,
                                                                  int len, int n, const int *p_cnt, const int *p_ptr, const int *p_col,
                                                                  const double *p_val, const double *p_dg, const int *ja, const double *M,
                                                                  const double *rhs, double *out, const int *status) {
    extern __shared__ __align__(16) unsigned char ring_raw[];
    RingSlot *ring = reinterpret_cast<RingSlot *>(ring_raw);
    double *y = reinterpret_cast<double *>(ring_raw + kRingBytes);  // n solved values + one cell that stays +0.0
    // the ring is handed over in BATCHES of kRingBatch slots (two batches): one full / one empty barrier per batch, so the walker
    // touches an mbarrier once per kRingBatch levels instead of twice per level
    __shared__ unsigned long long full[2], empty[2];
    pdl_prologue();
    if (status && *status != ST_RUNNING) return;
    const int tid = threadIdx.x, g = tid >> 7, lt = tid & (kRingRows - 1);
    if (tid == 0) {
        for (int q = 0; q < 2; ++q) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(full + q)), "r"(kRingBatch * kRingRows));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(empty + q)), "r"(kRingRows));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        y[n] = 0.0;
    }
    __syncthreads();
    if (g > 0) {
        // ---- loaders: group g - 1 fills the slots of chunks g - 1, g - 1 + kRingLoaders, ...  A chunk's loads are issued one turn before
        //      they are stored (two register sets, no copies), its row indices order[t] two turns before: no load waits for
        //      the address of another one, and the only wait of a turn is the ring's empty barrier.  Entries a row does not
        //      have get the coefficient +0.0 and point at the zero cell y[n]: fma(-0.0, +0.0, acc) = acc for every acc, so the
        //      walker runs kPlanW unpredicated FMAs. ----
        struct Ld { int i, cnt, c[kPlanW], rows, info, beg; double m[kPlanW], dg, r; };
        auto chunk_of = [&](int k) { return g - 1 + k * kRingLoaders; };
        auto load_index = [&](int j) -> int {
            if (j >= nchunks) return -1;
            const int beg = chunk_beg[j], rows = chunk_info[j] & 0xffff;
            return lt < rows ? order[beg + lt] : -1;
        };
        auto load_rows = [&](int j, int i, Ld &d) {
            d.i = i; d.cnt = 0; d.dg = 1.0; d.r = 0.0; d.rows = 0; d.info = 0; d.beg = 0;
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) { d.c[q] = n; d.m[q] = 0.0; }
            if (j >= nchunks) return;
            d.beg = chunk_beg[j]; d.info = chunk_info[j]; d.rows = d.info & 0xffff;
            if (lt < d.rows) {
                const int t = d.beg + lt;
                d.cnt = p_cnt[t];
#pragma unroll
                for (int q = 0; q < kPlanW; ++q) { d.c[q] = p_col[(size_t)q * len + t]; d.m[q] = p_val[(size_t)q * len + t]; }
                if (UPPER) d.dg = p_dg[t];
                if (i >= 0) d.r = rhs[i];
            }
        };
        auto store_rows = [&](int j, const Ld &d) {
            const int slot = j & (kRingSlots - 1);
            const int bb = (j >> 2) & 1;                           // batch barrier of the chunk (kRingBatch = 4)
            if (j >= kRingSlots) ring_wait<true>(empty + bb, (unsigned)(((j >> 3) - 1) & 1));        // the walker is done with the batch of chunk j - 8
            RingSlot &S = ring[slot];
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) {
                const bool has = q < d.cnt && d.i >= 0;
                S.c[q][lt] = 8 * (has ? d.c[q] : n);               // byte offset into y
                S.m[q][lt] = has ? d.m[q] : 0.0;
            }
            S.i[lt] = d.i; S.cnt[lt] = d.cnt; S.rhs[lt] = d.r;
            if (UPPER) { S.dg[lt] = d.dg; S.rcp[lt] = div_prepare(d.dg); }
            if (lt == 0) { S.rows = d.rows; S.last = d.info >> 16; S.beg = d.beg; }
            ring_arrive(full + bb);
        };
        Ld da, db;
        int ia = load_index(chunk_of(0)), ib = load_index(chunk_of(1));
        load_rows(chunk_of(0), ia, da);
        for (int k = 0; chunk_of(k) < nchunks; k += 2) {
            load_rows(chunk_of(k + 1), ib, db);
            ia = load_index(chunk_of(k + 2));
            store_rows(chunk_of(k), da);
            if (chunk_of(k + 1) >= nchunks) break;
            load_rows(chunk_of(k + 2), ia, da);
            ib = load_index(chunk_of(k + 3));
            store_rows(chunk_of(k + 1), db);
        }
    } else {
        // ---- the walker: the row of the NEXT chunk is read from its slot before the barrier of the current level whenever the
        //      loaders are ahead (one non-blocking test of the slot's full barrier), so the level-to-level chain is only
        //      barrier -> y[c] -> kPlanW DFMAs (-> division) -> y[i] -> barrier ----
        struct Row { int i, cnt, last, beg; unsigned yo[kPlanW]; double m[kPlanW], dg, rcp, rhs; };
        // (laundered through asm: otherwise the compiler re-derives the window base, S2R + LEA, at every use)
        auto opaque_addr = [](const void *p) { unsigned a = (unsigned)__cvta_generic_to_shared(p), b; asm volatile("mov.u32 %0, %1;" : "=r"(b) : "r"(a)); return b; };
        const unsigned y_a = opaque_addr(y), ring_a = opaque_addr(ring), full_a = opaque_addr(full), empty_a = opaque_addr(empty);
        constexpr unsigned kOffM = offsetof(RingSlot, m), kOffDg = offsetof(RingSlot, dg), kOffRcp = offsetof(RingSlot, rcp),
                           kOffRhs = offsetof(RingSlot, rhs), kOffC = offsetof(RingSlot, c), kOffI = offsetof(RingSlot, i),
                           kOffCnt = offsetof(RingSlot, cnt), kOffLast = offsetof(RingSlot, last), kOffBeg = offsetof(RingSlot, beg);
        auto read_slot = [&](int j, Row &r) {
            const unsigned sb = ring_a + (unsigned)(j & (kRingSlots - 1)) * (unsigned)sizeof(RingSlot);
            const unsigned l4 = sb + 4u * lt, l8 = sb + 8u * lt;
            r.last = lds_s32(sb + kOffLast); r.beg = lds_s32(sb + kOffBeg);
            r.i = lds_s32(l4 + kOffI); r.cnt = lds_s32(l4 + kOffCnt);
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) r.yo[q] = (unsigned)lds_s32(l4 + kOffC + q * 4 * kRingRows);      // byte offset; the address is formed at use
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) r.m[q] = lds_f64(l8 + kOffM + q * 8 * kRingRows);
            r.rhs = lds_f64(l8 + kOffRhs);
            if (UPPER) { r.dg = lds_f64(l8 + kOffDg); r.rcp = lds_f64(l8 + kOffRcp); } else { r.dg = 1.0; r.rcp = 1.0; }
            if ((j & (kRingBatch - 1)) == kRingBatch - 1)          // last slot of its batch: hand the batch back
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_a + 8u * (unsigned)((j >> 2) & 1)) : "memory");
        };
        auto batch_full = [&](int j, bool block) -> bool {         // is the batch of chunk j loaded?
            unsigned ok = 0, spins = 0;
            const unsigned bar = full_a + 8u * (unsigned)((j >> 2) & 1), par = (unsigned)((j >> 3) & 1);
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(bar), "r"(par) : "memory");
                if (!ok && block && ++spins > (1u << 22)) __trap();             // never hang on a broken ring
            } while (!ok && block);
            return ok != 0;
        };
        auto turn = [&](int j, Row &cur, Row &nxt, bool &have) {
            if (!have) { batch_full(j, true); read_slot(j, cur); }
            // the chain: y loads first, the next row's slot reads in their shadow, then the FMAs
            double yv[kPlanW];
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) yv[q] = lds_f64(y_a + cur.yo[q]);
            have = false;
            if (j + 1 < nchunks) {
                if (((j + 1) & (kRingBatch - 1)) != 0 || batch_full(j + 1, false)) { read_slot(j + 1, nxt); have = true; }
            }
            double acc = cur.rhs;
#pragma unroll
            for (int q = 0; q < kPlanW; ++q) acc = __fma_rn(-cur.m[q], yv[q], acc);
            if (cur.i >= 0) {
                if (cur.cnt > kPlanW) acc = ring_long_row_tail(acc, p_ptr[cur.beg + lt], cur.cnt, M, ja, y);   // rare, out of line
                if (UPPER) acc = div_finish(acc, cur.dg, cur.rcp);
                sts_f64(y_a + 8u * (unsigned)cur.i, acc);          // (a global store here would sit on the chain)
            }
            if (cur.last) asm volatile("bar.sync 1, %0;" ::"n"(kRingRows) : "memory");
        };
        Row ra, rb;
        bool have = false;
        for (int j = 0; j < nchunks; j += 2) {                     // nchunks is a multiple of kRingBatch
            turn(j, ra, rb, have);
            turn(j + 1, rb, ra, have);
        }
    }
    // ---- the solved vector leaves shared memory once, coalesced, by all threads ----
    __syncthreads();
    for (int i = tid; i < n; i += kRingThreads) out[i] = y[i];
}

__global__ void k_fill_bits(unsigned long long *p, unsigned long long v, int64_t cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < cnt; i += stride) p[i] = v;
}
// which sweep kernel a handle uses: 0 one launch per level (also: few levels, e.g. the multicolour ordering — every level
// is a full-occupancy bandwidth-bound launch), 1 single-CTA shared-memory sweep (small systems), 2 sync-free persistent sweep,
// 3 role-split ring sweep (small systems, default)
static int sweep_mode(const cudamat_solver *s) {
    if (!s->opt_sptrsv_syncfree) return 0;
    if (sizeof(double) * ((size_t)s->n + 1) + kRingBytes <= 226 * 1024 && !s->opt_sptrsv_no_smem && s->opt_sptrsv_ring) return 3;
    if (sizeof(double) * (size_t)s->n <= 200 * 1024 && !s->opt_sptrsv_no_smem) return 1;
    if (std::max(s->lvl_l.nlevels, s->lvl_u.nlevels) <= 32) return 0;
    return 2;
}
// chunk table of the ring sweep: <= 128 plan positions per chunk, `last` marks the end of a level (built at analysis time:
// nothing may allocate or synchronise inside a captured iteration)
static int ring_prepare(cudamat_solver *s) {
    if (sweep_mode(s) != 3) return CUDAMAT_OK;
    for (LevelSchedule *P : {&s->lvl_l, &s->lvl_u}) {
        LevelSchedule &LS = *P;
        if (LS.d_chunk_beg || LS.order_len <= 0) continue;
        std::vector<int> cb, ci;
        for (int l = 0; l < LS.nlevels; ++l)
            for (int off = LS.level_ptr[l]; off < LS.level_ptr[l + 1]; off += kRingRows) {
                const int rows = std::min(kRingRows, LS.level_ptr[l + 1] - off);
                cb.push_back(off); ci.push_back(rows | ((off + kRingRows >= LS.level_ptr[l + 1]) ? 1 << 16 : 0));
            }
        while (cb.size() % kRingBatch) { cb.push_back(0); ci.push_back(0); }       // empty chunks: the ring is handed over in whole batches
        LS.nchunks = (int)cb.size();
        CM_CUDA(dev_alloc((void **)&LS.d_chunk_beg, sizeof(int) * std::max<size_t>(cb.size(), 1)));
        CM_CUDA(dev_alloc((void **)&LS.d_chunk_info, sizeof(int) * std::max<size_t>(ci.size(), 1)));
        CM_CUDA(cudaMemcpyAsync(LS.d_chunk_beg, cb.data(), sizeof(int) * cb.size(), cudaMemcpyHostToDevice, s->stream));
        CM_CUDA(cudaMemcpyAsync(LS.d_chunk_info, ci.data(), sizeof(int) * ci.size(), cudaMemcpyHostToDevice, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));             // the host vectors die here
    }
    CM_CUDA(cudaFuncSetAttribute(k_sptrsv_ring<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    CM_CUDA(cudaFuncSetAttribute(k_sptrsv_ring<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    return CUDAMAT_OK;
}

int sptrsv_arm(cudamat_solver *s, double *vec) {
    if (s->n <= 0) return CUDAMAT_OK;
    if (sweepblk_active(s)) return CUDAMAT_OK;                 // block-wavefront sweeps signal through per-block flags
    if (sweep_mode(s) != 2) return CUDAMAT_OK;                 // only the sync-free sweep needs the sentinel
    int grid = (s->n + 1023) / 1024;
    if (grid > 148 * 16) grid = 148 * 16;
    k_fill_bits<<<grid, 256, 0, s->stream>>>(reinterpret_cast<unsigned long long *>(vec), kSentinelBits, s->n);
    s->launches++;
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

// ------------------------------------------------------------------------------------------
// host-side analysis (v1): level sets from the CSR pattern on the host
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Level analysis on the device (replaces cusparseDcsrsv_analysis, pbicgstab.cu:338,345; the first version of this
// file downloaded the pattern and walked it on the host: 0.9 s at 256^3).
//   k_find_diag        : position of the diagonal entry of every row (missing -> smallest such row)
//   k_levels_syncfree  : level(i) = 1 + max level of its dependencies, one launch per factor.  A warp owns 32
//                        consecutive positions of the sweep order (ticketed, so earlier positions always started);
//                        dependencies outside the warp are polled (level array pre-filled with -1), dependencies
//                        inside the warp are resolved by 32 shuffle steps over a per-lane dependency mask.
//   radix sort by level (stable: ascending rows inside a level) + k_level_starts + k_scatter_order build the padded
//   level-ordered row list.
// ------------------------------------------------------------------------------------------
__global__ void k_find_diag(int n, const int *ia, const int *ja, int *diag, int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int dp = -1;
    for (int p = ia[i]; p < ia[i + 1]; ++p) if (ja[p] == i) { dp = p; break; }
    diag[i] = dp;
    if (dp < 0) atomicMin(bad, i);
}
__device__ __forceinline__ int ld_relaxed_i32(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <bool UPPER>
__global__ void __launch_bounds__(256) k_levels_syncfree(int n, const int *ia, const int *ja, const int *diag, int *level,
                                                        unsigned *ticket, int *maxlevel) {
    __shared__ unsigned s_chunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunk = (n + 255) / 256;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned chunk = s_chunk;
        if (chunk >= (unsigned)nchunk) break;
        const int pbase = (int)chunk * 256 + warp * 32;            // first position of this warp
        const int pos = pbase + lane;
        const bool act = pos < n;
        const int i = act ? (UPPER ? n - 1 - pos : pos) : 0;
        int lv = 0;
        unsigned mask = 0;
        if (act) {
            const int p0 = UPPER ? diag[i] + 1 : ia[i], p1 = UPPER ? ia[i + 1] : diag[i];
            for (int p = p0; p < p1; ++p) {
                const int c = ja[p];
                const int cpos = UPPER ? n - 1 - c : c;
                if (cpos >= pbase) { mask |= 1u << (cpos - pbase); continue; }     // inside this warp (cpos < pos)
                int v = ld_relaxed_i32(level + c);
                unsigned spins = 0;
                while (v < 0) { if (++spins > (1u << 24)) __trap(); v = ld_relaxed_i32(level + c); }
                lv = max(lv, v + 1);
            }
        }
        const unsigned any = __ballot_sync(0xffffffffu, mask != 0);
        if (any) {
#pragma unroll 1
            for (int t = 0; t < 31; ++t) {                         // lane t is final once lanes < t are
                const int lt = __shfl_sync(0xffffffffu, lv, t);
                if ((mask >> t) & 1u) lv = max(lv, lt + 1);
            }
        }
        if (act) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(level + i), "r"(lv) : "memory");
        const int wmax = __reduce_max_sync(0xffffffffu, act ? lv : 0);
        if (lane == 0) atomicMax(maxlevel, wmax);
    }
}
// natural-order greedy colouring, same sync-free scheme as the level kernel: colour(i) = smallest colour not used by
// the neighbours that precede i (lower part of the row).  On a 5/7-point grid this is the red-black colouring: two
// colours with a perfectly regular layout, so the permuted sweeps stay coalesced.
__global__ void __launch_bounds__(256) k_greedy_color_syncfree(int n, const int *ia, const int *ja, int *color, unsigned *ticket, int *overflow) {
    __shared__ unsigned s_chunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunk = (n + 255) / 256;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned chunk = s_chunk;
        if (chunk >= (unsigned)nchunk) break;
        const int pbase = (int)chunk * 256 + warp * 32;
        const int i = pbase + lane;
        const bool act = i < n;
        unsigned long long used = 0ull;
        unsigned mask = 0;
        if (act) {
            for (int p = ia[i]; p < ia[i + 1]; ++p) {
                const int c = ja[p];
                if (c >= i) continue;                                  // later rows avoid this row's colour themselves
                if (c >= pbase) { mask |= 1u << (c - pbase); continue; }
                int v = ld_relaxed_i32(color + c);
                unsigned spins = 0;
                while (v < 0) { if (++spins > (1u << 24)) __trap(); v = ld_relaxed_i32(color + c); }
                if (v < 64) used |= 1ull << v;
            }
        }
        int myc = 0;
#pragma unroll 1
        for (int t = 0; t < 32; ++t) {                                 // lane t is final once lanes < t are
            if (lane == t) myc = __ffsll((long long)~used) - 1;
            const int ct = __shfl_sync(0xffffffffu, myc, t);
            if ((mask >> t) & 1u) { if (ct >= 0 && ct < 64) used |= 1ull << ct; }
        }
        if (act) {
            if (myc < 0) { *overflow = 1; myc = 63; }
            asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(color + i), "r"(myc) : "memory");
        }
    }
}
__global__ void k_iota(int n, int *v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
__global__ void k_level_starts(int n, const int *sorted_level, int *start /*[nlevels + 1]*/, int nlevels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = sorted_level[i];
    if (i == 0) { for (int q = 0; q <= l; ++q) start[q] = 0; }
    else { const int lp = sorted_level[i - 1]; for (int q = lp + 1; q <= l; ++q) start[q] = i; }
    if (i == n - 1) { for (int q = l + 1; q <= nlevels; ++q) start[q] = n; }
}
__global__ void k_scatter_order(int n, const int *sorted_level, const int *sorted_row, const int *start, const int *padded_ptr, int *order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = sorted_level[i];
    order[padded_ptr[l] + (i - start[l])] = sorted_row[i];
}

// level set of one factor entirely on the device; fills out.{nlevels, level_ptr, order_len, d_order, d_level_ptr}
static int device_schedule(cudamat_solver *s, bool upper, LevelSchedule &out) {
    const int n = s->n;
    int *d_level = nullptr, *d_rows = nullptr, *d_level2 = nullptr, *d_rows2 = nullptr, *d_misc = nullptr;
    CM_CUDA(dev_alloc((void **)&d_level, sizeof(int) * (size_t)std::max(n, 1)));
    CM_CUDA(dev_alloc((void **)&d_misc, sizeof(int) * 4));
    CM_CUDA(cudaMemsetAsync(d_level, 0xff, sizeof(int) * (size_t)std::max(n, 1), s->stream));
    CM_CUDA(cudaMemsetAsync(d_misc, 0, sizeof(int) * 4, s->stream));
    if (n > 0) {
        int occ = 0, sms = 0;
        if (upper) CM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_levels_syncfree<true>, 256, 0));
        else CM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_levels_syncfree<false>, 256, 0));
        CM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
        const int grid = std::max(1, std::min(occ * sms, (n + 255) / 256));       // co-resident: a polling CTA never starves a producer
        if (upper) k_levels_syncfree<true><<<grid, 256, 0, s->stream>>>(n, s->pre_ia, s->pre_ja, s->d_diag, d_level, (unsigned *)d_misc, d_misc + 1);
        else k_levels_syncfree<false><<<grid, 256, 0, s->stream>>>(n, s->pre_ia, s->pre_ja, s->d_diag, d_level, (unsigned *)d_misc, d_misc + 1);
        CM_CUDA(cudaGetLastError());
        s->launches++;
    }
    int maxl = 0;
    CM_CUDA(cudaMemcpyAsync(&maxl, d_misc + 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    const int nlevels = n > 0 ? maxl + 1 : 0;
    out.nlevels = nlevels;
    out.level_ptr.assign((size_t)nlevels + 1, 0);
    std::vector<int> start((size_t)nlevels + 1, 0);
    int *d_start = nullptr, *d_pad = nullptr;
    if (n > 0) {
        CM_CUDA(dev_alloc((void **)&d_rows, sizeof(int) * (size_t)n));
        CM_CUDA(dev_alloc((void **)&d_level2, sizeof(int) * (size_t)n));
        CM_CUDA(dev_alloc((void **)&d_rows2, sizeof(int) * (size_t)n));
        k_iota<<<(n + 255) / 256, 256, 0, s->stream>>>(n, d_rows);
        int bits = 1;
        while ((1 << bits) < nlevels + 1 && bits < 31) ++bits;
        size_t tmp_bytes = 0;
        CM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_level, d_level2, d_rows, d_rows2, n, 0, bits, s->stream));
        void *d_tmp = nullptr;
        CM_CUDA(dev_alloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
        CM_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_level, d_level2, d_rows, d_rows2, n, 0, bits, s->stream));
        CM_CUDA(dev_alloc((void **)&d_start, sizeof(int) * (size_t)(nlevels + 1)));
        CM_CUDA(dev_alloc((void **)&d_pad, sizeof(int) * (size_t)(nlevels + 1)));
        k_level_starts<<<(n + 255) / 256, 256, 0, s->stream>>>(n, d_level2, d_start, nlevels);
        CM_CUDA(cudaMemcpyAsync(start.data(), d_start, sizeof(int) * (size_t)(nlevels + 1), cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
        dev_free(d_tmp);
        s->launches += 3;
    }
    for (int l = 0; l < nlevels; ++l) out.level_ptr[l + 1] = out.level_ptr[l] + ((start[l + 1] - start[l] + 31) / 32) * 32;
    out.order_len = nlevels > 0 ? out.level_ptr[nlevels] : 0;
    CM_CUDA(dev_alloc((void **)&out.d_order, sizeof(int) * (size_t)std::max(out.order_len, 1)));
    CM_CUDA(cudaMemsetAsync(out.d_order, 0xff, sizeof(int) * (size_t)std::max(out.order_len, 1), s->stream));
    CM_CUDA(dev_alloc((void **)&out.d_level_ptr, sizeof(int) * (size_t)(nlevels + 2)));
    {
        std::vector<int> lp(out.level_ptr);
        lp.push_back(out.order_len);                       // level_ptr[nlevels + 1]: lets the kernel look one level ahead
        CM_CUDA(cudaMemcpyAsync(out.d_level_ptr, lp.data(), sizeof(int) * (size_t)(nlevels + 2), cudaMemcpyHostToDevice, s->stream));
        if (n > 0) {
            CM_CUDA(cudaMemcpyAsync(d_pad, out.level_ptr.data(), sizeof(int) * (size_t)(nlevels + 1), cudaMemcpyHostToDevice, s->stream));
            k_scatter_order<<<(n + 255) / 256, 256, 0, s->stream>>>(n, d_level2, d_rows2, d_start, d_pad, out.d_order);
            s->launches++;
        }
        CM_CUDA(cudaGetLastError());
        CM_CUDA(cudaStreamSynchronize(s->stream));
    }
    dev_free(d_level); dev_free(d_misc); dev_free(d_rows); dev_free(d_level2); dev_free(d_rows2); dev_free(d_start); dev_free(d_pad);
    return CUDAMAT_OK;
}

static int build_schedule(cudamat_solver *s, const std::vector<int> &level, int nlevels, LevelSchedule &out) {
    const int n = s->n;
    std::vector<int> cnt(nlevels + 1, 0);
    for (int i = 0; i < n; ++i) cnt[level[i] + 1]++;
    out.level_ptr.assign(nlevels + 1, 0);
    for (int l = 0; l < nlevels; ++l) out.level_ptr[l + 1] = out.level_ptr[l] + ((cnt[l + 1] + 31) / 32) * 32;
    out.order_len = out.level_ptr[nlevels];
    std::vector<int> order(out.order_len, -1), fill(nlevels, 0);
    for (int i = 0; i < n; ++i) { const int l = level[i]; order[out.level_ptr[l] + fill[l]++] = i; }
    out.nlevels = nlevels;
    CM_CUDA(dev_alloc((void **)&out.d_order, sizeof(int) * (size_t)std::max(out.order_len, 1)));
    CM_CUDA(cudaMemcpyAsync(out.d_order, order.data(), sizeof(int) * (size_t)out.order_len, cudaMemcpyHostToDevice, s->stream));
    CM_CUDA(dev_alloc((void **)&out.d_level_ptr, sizeof(int) * (size_t)(nlevels + 2)));
    {
        std::vector<int> lp(out.level_ptr);
        lp.push_back(out.order_len);                       // level_ptr[nlevels + 1]: lets the kernel look one level ahead
        CM_CUDA(cudaMemcpyAsync(out.d_level_ptr, lp.data(), sizeof(int) * (size_t)(nlevels + 2), cudaMemcpyHostToDevice, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
    }
    return CUDAMAT_OK;
}

void ilu0_release(cudamat_solver *s) {
    // the buffers go back to the stream-ordered pool (cudaFree of the 2 GB a 256^3 factor + plans hold costs more than a
    // sweep pair): nothing enqueued on the handle's stream may still use them
    cudaStreamSynchronize(s->stream);
    sweepblk_release(s);
    if (s->d_M) dev_free(s->d_M);
    if (s->d_diag) dev_free(s->d_diag);
    for (LevelSchedule *P : {&s->lvl_l, &s->lvl_u}) {
        if (P->d_order) dev_free(P->d_order);
        if (P->d_level_ptr) dev_free(P->d_level_ptr);
        if (P->d_chunk_beg) dev_free(P->d_chunk_beg);
        if (P->d_chunk_info) dev_free(P->d_chunk_info);
        P->d_chunk_beg = P->d_chunk_info = nullptr; P->nchunks = 0;
        if (P->d_cnt) dev_free(P->d_cnt);
        if (P->d_ptr) dev_free(P->d_ptr);
        if (P->d_col) dev_free(P->d_col);
        if (P->d_val) dev_free(P->d_val);
        if (P->d_dg) dev_free(P->d_dg);
    }
    if (s->d_perm) dev_free(s->d_perm);
    if (s->prm_ia) dev_free(s->prm_ia);
    if (s->prm_ja) dev_free(s->prm_ja);
    if (s->prm_a) dev_free(s->prm_a);
    s->d_perm = nullptr; s->prm_ia = nullptr; s->prm_ja = nullptr; s->prm_a = nullptr;
    if (s->blk_ia) dev_free(s->blk_ia);
    if (s->blk_ja) dev_free(s->blk_ja);
    if (s->blk_a) dev_free(s->blk_a);
    s->blk_ia = nullptr; s->blk_ja = nullptr; s->blk_a = nullptr; s->blk_nnz = 0;
    if (s->d_flag) dev_free(s->d_flag);
    if (s->d_ticket) dev_free(s->d_ticket);
    s->d_M = nullptr; s->d_diag = nullptr; s->lvl_l = LevelSchedule(); s->lvl_u = LevelSchedule();
    s->d_flag = nullptr; s->d_ticket = nullptr;
    s->ticket_base_l = s->ticket_base_u = 0; s->epoch = 0;
}

static double now_s() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// Sharded handles: the preconditioner is block-Jacobi ILU(0) — each rank factors the diagonal block of its row
// shard (entries whose column lives on another rank are left out), so the sweeps need no communication.  It is a
// different (weaker) preconditioner than the global ILU(0) of the single-GPU run: iteration counts differ, the
// solution does not (SURVEY.md §8e "ILU0 across GPUs", option block-Jacobi).
__global__ void k_block_count(int n, const int *ia, const int *ja, int *cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { cnt[n] = 0; return; }
    int c = 0;
    for (int p = ia[i]; p < ia[i + 1]; ++p) c += (ja[p] < n) ? 1 : 0;
    cnt[i] = c;
}
__global__ void k_block_fill(int n, const int *ia, const int *ja, const double *a, const int *bia, int *bja, double *ba) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int q = bia[i];
    for (int p = ia[i]; p < ia[i + 1]; ++p)
        if (ja[p] < n) { bja[q] = ja[p]; ba[q] = a[p]; ++q; }
}
int exclusive_scan_inplace(int *d, int64_t cnt, cudaStream_t st);      // kernels.cu

static int build_local_block(cudamat_solver *s) {
    const int n = s->n;
    CM_CUDA(dev_alloc((void **)&s->blk_ia, sizeof(int) * (size_t)(n + 1)));
    k_block_count<<<(n + 1 + 255) / 256, 256, 0, s->stream>>>(n, s->d_ia, s->d_ja, s->blk_ia);
    CM_CUDA(cudaGetLastError());
    int rc = exclusive_scan_inplace(s->blk_ia, (int64_t)n + 1, s->stream);
    if (rc) return rc;
    int last = 0;
    CM_CUDA(cudaMemcpyAsync(&last, s->blk_ia + n, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    s->blk_nnz = last;
    CM_CUDA(dev_alloc((void **)&s->blk_ja, sizeof(int) * (size_t)std::max(last, 1)));
    CM_CUDA(dev_alloc((void **)&s->blk_a, sizeof(double) * (size_t)std::max(last, 1)));
    if (n > 0) k_block_fill<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->d_ia, s->d_ja, s->d_a, s->blk_ia, s->blk_ja, s->blk_a);
    CM_CUDA(cudaGetLastError());
    s->launches += 2;
    return CUDAMAT_OK;
}

// ------------------------------------------------------------------------------------------
// Opt-in multicolour reordering of the preconditioner matrix ("ilu0_reorder" = 1; SURVEY.md 8f-4, hard part H3).
// The natural ordering of a 7-point grid gives 3N-2 dependent levels per sweep; a colouring of the adjacency graph
// gives (number of colours) levels, i.e. bandwidth-bound sweeps.  ILU(0) of the permuted matrix P A P^T is a DIFFERENT
// (usually weaker) preconditioner, so iteration counts differ from the reference ordering — hence opt-in; the solution
// is the same.  Natural-order greedy colouring (sync-free kernel, red-black on 5/7-point grids), rows stably sorted by
// colour, the permuted CSR is rebuilt with ascending columns, and the usual analysis / factorisation / sweeps run on it;
// M^-1 v = P^T (LU)^-1 P v costs one gather and one scatter around each sweep pair.
// ------------------------------------------------------------------------------------------
__global__ void k_perm_count(int n, const int *perm, const int *ia, int *cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { cnt[n] = 0; return; }
    const int o = perm[i];
    cnt[i] = ia[o + 1] - ia[o];
}
__global__ void k_perm_fill(int n, const int *perm, const int *inv, const int *ia, const int *ja, const double *a,
                            const int *pia, int *pja, double *pa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = perm[i];
    const int q0 = pia[i];
    int m = 0;
    for (int p = ia[o]; p < ia[o + 1]; ++p) {                      // insertion sort by the new column index
        const int c = inv[ja[p]];
        const double v = a[p];
        int pos = m;
        while (pos > 0 && pja[q0 + pos - 1] > c) { pja[q0 + pos] = pja[q0 + pos - 1]; pa[q0 + pos] = pa[q0 + pos - 1]; --pos; }
        pja[q0 + pos] = c; pa[q0 + pos] = v;
        ++m;
    }
}
__global__ void k_invert_perm(int n, const int *perm, int *inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = i;
}
__global__ void k_gather_perm(int n, const int *perm, const double *in, double *out, const int *status) {
    if (status && *status != ST_RUNNING) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
__global__ void k_scatter_perm(int n, const int *perm, const double *in, double *out, const int *status) {
    if (status && *status != ST_RUNNING) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[perm[i]] = in[i];
}
int launch_permute(cudamat_solver *s, bool scatter, const double *in, double *out) {
    if (s->n <= 0) return CUDAMAT_OK;
    const int *status = s->d_sc ? &s->d_sc->status : nullptr;
    if (scatter) k_scatter_perm<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->n, s->d_perm, in, out, status);
    else k_gather_perm<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->n, s->d_perm, in, out, status);
    s->launches++;
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

// replaces pre_* by the multicolour-permuted matrix; on failure (more than 64 colours) keeps the original ordering
static int build_multicolor(cudamat_solver *s) {
    const int n = s->n;
    if (n <= 0) return CUDAMAT_OK;
    int *d_color = nullptr, *d_misc = nullptr, *d_color2 = nullptr, *d_iota = nullptr, *d_inv = nullptr;
    CM_CUDA(dev_alloc((void **)&d_color, sizeof(int) * (size_t)n));
    CM_CUDA(dev_alloc((void **)&d_misc, sizeof(int) * 2));
    CM_CUDA(cudaMemsetAsync(d_color, 0xff, sizeof(int) * (size_t)n, s->stream));
    CM_CUDA(cudaMemsetAsync(d_misc, 0, sizeof(int) * 2, s->stream));
    int h[2] = {0, 0};
    {
        int occ = 0, sms = 0;
        CM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_greedy_color_syncfree, 256, 0));
        CM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
        const int grid = std::max(1, std::min(occ * sms, (n + 255) / 256));
        k_greedy_color_syncfree<<<grid, 256, 0, s->stream>>>(n, s->pre_ia, s->pre_ja, d_color, (unsigned *)d_misc, d_misc + 1);
        CM_CUDA(cudaMemcpyAsync(h, d_misc, sizeof h, cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
        s->launches++;
        h[0] = 0;
    }
    if (h[1]) { dev_free(d_color); dev_free(d_misc); return CUDAMAT_OK; }                   // > 64 colours: keep the reference ordering
    // perm[new] = old: stable sort of the rows by colour
    CM_CUDA(dev_alloc((void **)&d_color2, sizeof(int) * (size_t)n));
    CM_CUDA(dev_alloc((void **)&d_iota, sizeof(int) * (size_t)n));
    CM_CUDA(dev_alloc((void **)&s->d_perm, sizeof(int) * (size_t)n));
    CM_CUDA(dev_alloc((void **)&d_inv, sizeof(int) * (size_t)n));
    k_iota<<<(n + 255) / 256, 256, 0, s->stream>>>(n, d_iota);
    size_t tmp_bytes = 0;
    CM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_color, d_color2, d_iota, s->d_perm, n, 0, 6, s->stream));
    void *d_tmp = nullptr;
    CM_CUDA(dev_alloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
    CM_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_color, d_color2, d_iota, s->d_perm, n, 0, 6, s->stream));
    k_invert_perm<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->d_perm, d_inv);
    // permuted CSR with ascending columns
    CM_CUDA(dev_alloc((void **)&s->prm_ia, sizeof(int) * (size_t)(n + 1)));
    k_perm_count<<<(n + 1 + 255) / 256, 256, 0, s->stream>>>(n, s->d_perm, s->pre_ia, s->prm_ia);
    int rc = exclusive_scan_inplace(s->prm_ia, (int64_t)n + 1, s->stream);
    if (rc) return rc;
    CM_CUDA(dev_alloc((void **)&s->prm_ja, sizeof(int) * (size_t)std::max<int64_t>(s->pre_nnz, 1)));
    CM_CUDA(dev_alloc((void **)&s->prm_a, sizeof(double) * (size_t)std::max<int64_t>(s->pre_nnz, 1)));
    k_perm_fill<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->d_perm, d_inv, s->pre_ia, s->pre_ja, s->pre_a, s->prm_ia, s->prm_ja, s->prm_a);
    CM_CUDA(cudaGetLastError());
    CM_CUDA(cudaStreamSynchronize(s->stream));
    s->launches += 5;
    dev_free(d_tmp); dev_free(d_color); dev_free(d_color2); dev_free(d_iota); dev_free(d_inv); dev_free(d_misc);
    s->pre_ia = s->prm_ia; s->pre_ja = s->prm_ja; s->pre_a = s->prm_a;
    return CUDAMAT_OK;
}

int ilu0_analyze_and_factor(cudamat_solver *s, cudamat_stats *st) {
    ilu0_release(s);
    const int n = s->n;
    // the matrix the preconditioner is built from: the whole matrix, or the local diagonal block of a shard
    s->pre_ia = s->d_ia; s->pre_ja = s->d_ja; s->pre_a = s->d_a; s->pre_nnz = s->nnz;
    if (s->nhalo > 0 || s->row0 != 0 || s->row1 != s->n_global) {
        int rcb = build_local_block(s);
        if (rcb) return rcb;
        s->pre_ia = s->blk_ia; s->pre_ja = s->blk_ja; s->pre_a = s->blk_a; s->pre_nnz = s->blk_nnz;
    }
    double t0 = now_s();
    if (s->opt_ilu0_reorder) { int rcm = build_multicolor(s); if (rcm) return rcm; }
    const int64_t nnz = s->pre_nnz;
    int nl = 0, nu = 0, rc;
    CM_CUDA(dev_alloc((void **)&s->d_diag, sizeof(int) * (size_t)std::max(n, 1)));
    if (!s->opt_host_analysis) {
        int *d_bad = nullptr;
        CM_CUDA(dev_alloc((void **)&d_bad, sizeof(int)));
        CM_CUDA(cudaMemsetAsync(d_bad, 0x7f, sizeof(int), s->stream));
        if (n > 0) k_find_diag<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->pre_ia, s->pre_ja, s->d_diag, d_bad);
        int bad = 0;
        CM_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
        dev_free(d_bad);
        s->launches++;
        if (bad != 0x7f7f7f7f) {
            set_error("ILU0: row %d has no structural diagonal entry (precondition pbicgstab.h:118)", bad);
            return CUDAMAT_E_NO_DIAGONAL;
        }
        if ((rc = device_schedule(s, false, s->lvl_l))) return rc;
        if ((rc = device_schedule(s, true, s->lvl_u))) return rc;
        nl = s->lvl_l.nlevels; nu = s->lvl_u.nlevels;
    } else {
        // host cross-check path ("host_analysis" option): download the pattern and walk it serially
        std::vector<int> ia(n + 1), ja((size_t)nnz);
        CM_CUDA(cudaMemcpyAsync(ia.data(), s->pre_ia, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaMemcpyAsync(ja.data(), s->pre_ja, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
        std::vector<int> diag(n), lvl(n);
        for (int i = 0; i < n; ++i) {
            int dp = -1;
            for (int p = ia[i]; p < ia[i + 1]; ++p) if (ja[p] == i) { dp = p; break; }
            if (dp < 0) {
                set_error("ILU0: row %d has no structural diagonal entry (precondition pbicgstab.h:118)", i);
                return CUDAMAT_E_NO_DIAGONAL;
            }
            diag[i] = dp;
        }
        for (int i = 0; i < n; ++i) {
            int lv = 0;
            for (int p = ia[i]; p < diag[i]; ++p) lv = std::max(lv, lvl[ja[p]] + 1);
            lvl[i] = lv; nl = std::max(nl, lv + 1);
        }
        if ((rc = build_schedule(s, lvl, nl, s->lvl_l))) return rc;
        for (int i = n - 1; i >= 0; --i) {
            int lv = 0;
            for (int p = diag[i] + 1; p < ia[i + 1]; ++p) lv = std::max(lv, lvl[ja[p]] + 1);
            lvl[i] = lv; nu = std::max(nu, lv + 1);
        }
        if ((rc = build_schedule(s, lvl, nu, s->lvl_u))) return rc;
        CM_CUDA(cudaMemcpyAsync(s->d_diag, diag.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    }
    s->epoch = 0;
    CM_CUDA(cudaStreamSynchronize(s->stream));
    if (st) { st->t_analysis += now_s() - t0; st->levels_l = nl; st->levels_u = nu; }

    // factorisation on a copy of A (pbicgstab.cu:316,359)
    t0 = now_s();
    CM_CUDA(dev_alloc((void **)&s->d_M, sizeof(double) * (size_t)std::max<int64_t>(nnz, 1)));
    CM_CUDA(cudaMemcpyAsync(s->d_M, s->pre_a, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice, s->stream));
    int *d_zp = nullptr;
    CM_CUDA(dev_alloc((void **)&d_zp, sizeof(int)));
    const int big = 0x7fffffff;
    CM_CUDA(cudaMemcpyAsync(d_zp, &big, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    for (int l = 0; l < nl; ++l) {
        const int off = s->lvl_l.level_ptr[l], cnt = s->lvl_l.level_ptr[l + 1] - off;
        if (cnt == 0) continue;
        k_ilu0_level<<<(cnt + 127) / 128, 128, 0, s->stream>>>(s->lvl_l.d_order + off, cnt, s->pre_ia, s->pre_ja, s->d_diag, s->d_M, d_zp);
        s->launches++;
    }
    CM_CUDA(cudaGetLastError());
    int zp = big;
    CM_CUDA(cudaMemcpyAsync(&zp, d_zp, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    dev_free(d_zp);
    s->zero_pivot = (zp == big) ? 0 : -(1 + zp);
    // level-ordered sweep plans (coalesced operands for the sync-free sweeps)
    for (int u = 0; u < 2; ++u) {
        LevelSchedule &P = u ? s->lvl_u : s->lvl_l;
        const size_t len = (size_t)std::max(P.order_len, 1);
        CM_CUDA(dev_alloc((void **)&P.d_cnt, sizeof(int) * len));
        CM_CUDA(dev_alloc((void **)&P.d_ptr, sizeof(int) * len));
        CM_CUDA(dev_alloc((void **)&P.d_col, sizeof(int) * len * kPlanW));
        CM_CUDA(dev_alloc((void **)&P.d_val, sizeof(double) * len * kPlanW));
        CM_CUDA(dev_alloc((void **)&P.d_dg, sizeof(double) * len));
        if (P.order_len > 0) {
            const int grid = (P.order_len + 255) / 256;
            if (u) k_build_plan<true><<<grid, 256, 0, s->stream>>>(P.d_order, P.order_len, s->pre_ia, s->pre_ja, s->d_diag, s->d_M, P.d_cnt, P.d_ptr, P.d_col, P.d_val, P.d_dg);
            else   k_build_plan<false><<<grid, 256, 0, s->stream>>>(P.d_order, P.order_len, s->pre_ia, s->pre_ja, s->d_diag, s->d_M, P.d_cnt, P.d_ptr, P.d_col, P.d_val, P.d_dg);
            s->launches++;
        }
    }
    CM_CUDA(cudaGetLastError());
    CM_CUDA(cudaStreamSynchronize(s->stream));
    // 7-point grid stencils (structure found by the MARCH analysis): block-wavefront sweeps
    if ((rc = sweepblk_plan(s))) return rc;
    if ((rc = ring_prepare(s))) return rc;
    if (st) { st->t_ilu0 += now_s() - t0; st->zero_pivot = s->zero_pivot; }
    return CUDAMAT_OK;
}

// With the sync-free schedule `out` must be armed (all sentinel) on entry.
int launch_sptrsv(cudamat_solver *s, bool upper, const double *rhs, double *out) {
    if (!s->d_M) { set_error("sptrsv: ILU0 factor not available (call cudamat_analyze with CUDAMAT_MODE_ILU0)"); return CUDAMAT_E_STATE; }
    if (sweepblk_active(s) && rhs != out) return launch_sptrsv_blocked(s, upper, rhs, out);
    const LevelSchedule &L = upper ? s->lvl_u : s->lvl_l;
    const int *status = s->d_sc ? &s->d_sc->status : nullptr;
    const size_t smem_y = sizeof(double) * (size_t)s->n;
    const int mode = sweep_mode(s);
    if (mode == 3 && L.order_len > 0) {
        // small system, role-split CTA: one walker group + loader groups around a shared-memory ring
        const LevelSchedule &LS = L;
        if (!LS.d_chunk_beg) { set_error("sptrsv: ring plan missing"); return CUDAMAT_E_STATE; }
        const void *kern = upper ? (const void *)k_sptrsv_ring<true> : (const void *)k_sptrsv_ring<false>;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(kRingThreads); cfg.stream = s->stream; cfg.dynamicSmemBytes = kRingBytes + smem_y + sizeof(double);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        const int *a_order = L.d_order, *a_cb = LS.d_chunk_beg, *a_ci = LS.d_chunk_info, *a_cnt = L.d_cnt, *a_ptr = L.d_ptr, *a_col = L.d_col, *a_ja = s->pre_ja;
        int a_nc = LS.nchunks, a_len = L.order_len, a_n = s->n;
        const double *a_val = L.d_val, *a_dg = L.d_dg, *a_M = s->d_M, *a_rhs = rhs;
        void *args[] = {&a_order, &a_cb, &a_ci, &a_nc, &a_len, &a_n, &a_cnt, &a_ptr, &a_col, &a_val, &a_dg, &a_ja, &a_M, &a_rhs, &out, &status};
        CM_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
        s->launches++;
    } else if (mode == 1 && L.order_len > 0) {
        // small system: one CTA, solved values in shared memory, a CTA barrier per level
        const void *kern = upper ? (const void *)k_sptrsv_smem<true> : (const void *)k_sptrsv_smem<false>;
        if (!s->sptrsv_smem_ready) {
            CM_CUDA(cudaFuncSetAttribute(k_sptrsv_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            CM_CUDA(cudaFuncSetAttribute(k_sptrsv_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            s->sptrsv_smem_ready = true;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(kSmemSweepThreads); cfg.stream = s->stream; cfg.dynamicSmemBytes = smem_y;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        const int *a_order = L.d_order, *a_lp = L.d_level_ptr, *a_cnt = L.d_cnt, *a_ptr = L.d_ptr, *a_col = L.d_col, *a_ja = s->pre_ja;
        int a_nl = L.nlevels, a_len = L.order_len, a_n = s->n;
        const double *a_val = L.d_val, *a_dg = L.d_dg, *a_M = s->d_M, *a_rhs = rhs;
        void *args[] = {&a_order, &a_lp, &a_nl, &a_len, &a_n, &a_cnt, &a_ptr, &a_col, &a_val, &a_dg, &a_ja, &a_M, &a_rhs, &out, &status};
        CM_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
        s->launches++;
    } else if (mode == 2 && L.order_len > 0) {
        if (s->sptrsv_grid == 0) {
            int occ_l = 0, occ_u = 0, sms = 0;
            CM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l, k_sptrsv_syncfree<false>, 256, 0));
            CM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_u, k_sptrsv_syncfree<true>, 256, 0));
            CM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
            int per_sm = std::min(occ_l, occ_u);
            // 2 CTAs/SM measured best on B200 (profiles/): enough look-ahead to hide the plan loads, few pollers
            per_sm = std::min(per_sm, s->opt_sptrsv_ctas_per_sm > 0 ? s->opt_sptrsv_ctas_per_sm : 2);
            s->sptrsv_grid = std::max(1, per_sm * sms);
        }
        const int grid = std::min(s->sptrsv_grid, (L.order_len + 255) / 256);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        if (upper)
            CM_CUDA(cudaLaunchKernelEx(&cfg, k_sptrsv_syncfree<true>, (const int *)L.d_order, L.order_len, (const int *)L.d_cnt, (const int *)L.d_ptr,
                                       (const int *)L.d_col, (const double *)L.d_val, (const double *)L.d_dg, s->pre_ja,
                                       (const double *)s->d_M, rhs, out, status));
        else
            CM_CUDA(cudaLaunchKernelEx(&cfg, k_sptrsv_syncfree<false>, (const int *)L.d_order, L.order_len, (const int *)L.d_cnt, (const int *)L.d_ptr,
                                       (const int *)L.d_col, (const double *)L.d_val, (const double *)L.d_dg, s->pre_ja,
                                       (const double *)s->d_M, rhs, out, status));
        s->launches++;
    } else {
        for (int l = 0; l < L.nlevels; ++l) {
            const int off = L.level_ptr[l], cnt = L.level_ptr[l + 1] - off;
            if (cnt == 0) continue;
            if (upper)
                k_sptrsv_level<true><<<(cnt + 127) / 128, 128, 0, s->stream>>>(L.d_order + off, cnt, s->pre_ia, s->pre_ja, s->d_diag, s->d_M, rhs, out, status);
            else
                k_sptrsv_level<false><<<(cnt + 127) / 128, 128, 0, s->stream>>>(L.d_order + off, cnt, s->pre_ia, s->pre_ja, s->d_diag, s->d_M, rhs, out, status);
            s->launches++;
        }
    }
    CM_CUDA(cudaGetLastError());
    return CUDAMAT_OK;
}

}  // namespace cudamat
