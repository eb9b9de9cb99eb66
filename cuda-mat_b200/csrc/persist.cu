// persist.cu — the unpreconditioned BiCGSTAB iteration (gpu_pbicgstab2, pbicgstab.cu:581-754) as ONE persistent,
// co-resident kernel per batch of iterations, for systems (or shards) small enough that the loop is bound by
// kernel-boundary latency instead of bandwidth.
//
// The per-kernel loop spends ~8 dependent launches + 3 reduction hops per iteration (~40 us at n = 10^4, ~160 us per
// iteration for a 2.1 M-row shard of 256^3 on 8 GPUs, where the bandwidth work is 60 us).  Here the grid is launched
// cooperatively (every CTA resident), the five phases of an iteration are separated by grid-wide barriers, and the scalar
// recurrences are computed REDUNDANTLY by every CTA from the tile partials — no kernel boundary, no host round trip, the
// scalars never leave the SMs:
//     A  p = r + beta (p - omega v)            [halo rows pushed to the neighbours]      barrier
//     B  v = (A + diag d) p ; tile partials of rhat.v                                     barrier  -> alpha
//     C  s = r - alpha v                       [halo rows pushed]                         barrier
//     D  t = (A + diag d) s ; tile partials of t.s, t.t                                   barrier  -> omega
//     E  x, r update ; tile partials of rhat.r, r.r                                       barrier  -> beta, ||r||, status
// Arithmetic: exactly the spec of DESIGN.md §3 — the element-wise forms of k_update_p/s/xr, the row sums of rowfuncs.cuh,
// the slab -> tile -> group -> final reduction tree — so the results are bit-identical to the per-kernel loop and the oracle.
// Multi-GPU (peer-memory path): a rank's group partials are pushed to every rank's gather array by CTA 0 right after the
// barrier, all CTAs then wait for the arrival flags (st.release.sys / ld.acquire.sys, as k_reduce_finish does); the halo rows
// of p and s travel exactly as in the per-kernel loop (halo_store / halo_signal / halo_wait of internal.cuh).
#include "solver.h"
#include "rowfuncs.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace cudamat {

struct PersistArgs {
    int n;                                   // local rows
    const int *ia, *ja; const double *val;   // CSR, [local | halo] column numbering
    const unsigned char *cls;                // class id per row (dictionary with values) or nullptr: CSR row sums
    double *r0, *r, *v, *p, *sv, *t, *x;     // work vectors
    const double *d;                         // diagonal shift or nullptr
    RedCtx rc; DevScalars *sc; double *hist;
    HaloPush hp_p, hp_s;                     // epochs of the FIRST iteration of this launch (+1 per iteration)
    HaloWait hw_p, hw_s;
    const unsigned long long *red_flags;     // arrival flags of the partial-sum gathers (peer-memory path)
    int iters;                               // iterations of this launch
};

// Operand gathers.  Vectors are written by other CTAs of THIS kernel, so the read-only path (__ldg) of the stand-alone SpMV
// kernels is not usable; ordinary loads are: the grid / cluster barrier in front of every SpMV phase carries a gpu-scope
// fence, which invalidates the SM's L1, and nobody reads the vector between that barrier and the phase.  Halo columns (>= n,
// stored by a PEER over NVLink, possibly after the barrier) bypass L1.
__device__ __forceinline__ double gather_op(const double *x, int col, int n) { return col < n ? x[col] : __ldcg(x + col); }

// R() of the spec over `m` doubles in global memory (read through L2)
__device__ __forceinline__ double R_global(const double *v, int m, int lane) { return warp_reduce_values_cg(v, m, lane); }

// one reduction point: tile partials are complete (grid barrier passed) -> red[0..nq)
// s_tmp: [kMaxQ][64] shared scratch.  All threads of the CTA call.
__device__ __forceinline__ void persist_reduce(const PersistArgs &A, int nq, unsigned long long epoch, double (*s_tmp)[64], double *s_red,
                                               int *status, int warp, int lane) {
    const RedCtx &rc = A.rc;
    const bool multi = rc.p2p.world > 0;
    const double *src = rc.slots;
    if (multi) {
        const bool tiles = rc.exch_level == 1;                     // unaligned shards exchange TILE partials (global tile index)
        // CTA 0: this rank's partials -> every rank's gather array + arrival flags
        if (blockIdx.x == 0) {
            if (!tiles) {
                for (int w = warp; w < nq * rc.ngroup_loc; w += kCtaWarps) {
                    const int q = w / rc.ngroup_loc, g = w % rc.ngroup_loc;
                    const int in_group = min(kGroupTiles, rc.ntile - g * kGroupTiles);
                    const double gp = R_global(rc.tile_part + (size_t)q * rc.tile_stride + (size_t)g * kGroupTiles, in_group, lane);
                    if (lane == 0) __stcg(rc.slots + (size_t)q * rc.slot_stride + rc.group0 + g, gp);
                }
                __threadfence();
                __syncthreads();
            }
            RedCtx rr = rc;
            rr.p2p.epoch = epoch;
            p2p_push(rr, tiles ? rc.exch : rc.slots, nq);
        }
        if ((int)threadIdx.x < rc.p2p.world) {                    // every rank's partial sums must have arrived
            unsigned spins = 0;
            while (ld_acquire_sys_u64(A.red_flags + threadIdx.x) < epoch)
                if (++spins > kPeerSpinLimit) { atomicExch(status, ST_COMM_TIMEOUT); break; }
        }
        __syncthreads();
        src = rc.p2p.peers[rc.p2p.me].gather[epoch & 1ull];
        if (tiles) {                                               // groups over the GLOBAL tile index, then the final sum
            const int ngroups = (rc.ntile_global + kGroupTiles - 1) / kGroupTiles;        // <= 64 (checked by comm_init)
            for (int w = warp; w < nq * ngroups; w += kCtaWarps) {
                const int q = w / ngroups, g = w % ngroups;
                const double gp = R_global(src + (size_t)q * rc.p2p.stride + (size_t)g * kGroupTiles,
                                           min(kGroupTiles, rc.ntile_global - g * kGroupTiles), lane);
                if (lane == 0) s_tmp[q][g] = gp;
            }
            __syncthreads();
            if (warp < nq) {
                const double f = warp_reduce_values_smem(s_tmp[warp], ngroups, lane);
                if (lane == 0) s_red[warp] = f;
            }
        } else if (warp < nq) {
            const double f = R_global(src + (size_t)warp * rc.p2p.stride, rc.nslots, lane);
            if (lane == 0) s_red[warp] = f;
        }
    } else {
        // single GPU: every CTA forms the group partials and the final sum redundantly
        const int ngroups = rc.ngroup_loc;                         // <= 64
        for (int w = warp; w < nq * ngroups; w += kCtaWarps) {
            const int q = w / ngroups, g = w % ngroups;
            const int in_group = min(kGroupTiles, rc.ntile - g * kGroupTiles);
            const double gp = R_global(rc.tile_part + (size_t)q * rc.tile_stride + (size_t)g * kGroupTiles, in_group, lane);
            if (lane == 0) s_tmp[q][g] = gp;
        }
        __syncthreads();
        if (warp < nq) {
            const double f = warp_reduce_values_smem(s_tmp[warp], ngroups, lane);
            if (lane == 0) s_red[warp] = f;
        }
    }
    __syncthreads();
}

// slab sums of one tile (one per warp-slab, in s_slab[q][64]) -> tile partial R(64) -> tile_part.  All threads call.
__device__ __forceinline__ void persist_tile_partial(const RedCtx &rc, int nq, int tile, int nslab, double (*s_slab)[64], int warp, int lane) {
    __syncthreads();
    if (warp < nq) {
        const double tp = warp_reduce_values_smem(s_slab[warp], nslab, lane);
        if (lane == 0) {
            __stcg(rc.tile_part + (size_t)warp * rc.tile_stride + tile, tp);
            if (rc.exch_level == 1) __stcg(rc.exch + (size_t)warp * rc.exch_stride + rc.tile0 + tile, tp);
        }
    }
    __syncthreads();
}

template <bool CLS>
__device__ __forceinline__ double persist_rowsum(const PersistArgs &A, const DictParam &D, const double *x, int row0, int row, bool active, int lane) {
    if (CLS) {
        // offsets and values of the row's class, entries in storage order; x is read through L2 (ld.cg): it was written by other
        // CTAs of THIS kernel, the non-coherent path (__ldg) of the stand-alone SpMV kernels could return stale L1 lines
        const int cid = active ? (int)__ldg(A.cls + row) : 0;
        const int len = active ? D.len[cid] : 0;
        const int *off = D.off + cid * kDictLen;
        const double *dv = D.val + cid * kDictLen;
        const int maxlen = __reduce_max_sync(0xffffffffu, len);
        double sum = 0.0;
#pragma unroll 1
        for (int k0 = 0; k0 < maxlen; k0 += 4) {
            double av[4], xv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool p = (k0 + q) < len;
                xv[q] = p ? gather_op(x, row + off[k0 + q], A.n) : 0.0;
                av[q] = p ? dv[k0 + q] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if ((k0 + q) < len) sum = __fma_rn(av[q], xv[q], sum);
        }
        return sum;
    }
    int s = 0, e = 0;
    if (active) { s = __ldg(A.ia + row); e = __ldg(A.ia + row + 1); }
    const int len = e - s;
    const int shortlen = (len <= kLongRow) ? len : 0;
    const int maxlen = __reduce_max_sync(0xffffffffu, shortlen);
    double sum = 0.0;
    for (int k0 = 0; k0 < maxlen; k0 += 4) {
        int cj[4]; double av[4], xv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool p = (k0 + q) < shortlen;
            cj[q] = p ? __ldg(A.ja + s + k0 + q) : -1;
            av[q] = p ? __ldg(A.val + s + k0 + q) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q] = (cj[q] >= 0) ? gather_op(x, cj[q], A.n) : 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if ((k0 + q) < shortlen) sum = __fma_rn(av[q], xv[q], sum);
    }
    unsigned lm = __ballot_sync(0xffffffffu, len > kLongRow);
    while (lm) {
        const int src = __ffs(lm) - 1;
        lm &= lm - 1;
        const int ss = __shfl_sync(0xffffffffu, s, src), ee = __shfl_sync(0xffffffffu, e, src);
        double acc = 0.0;
        for (int k = ss + lane; k < ee; k += 32) acc = __fma_rn(__ldg(A.val + k), gather_op(x, __ldg(A.ja + k), A.n), acc);
        acc = warp_butterfly(acc);
        if (lane == src) sum = acc;
    }
    return sum;
}

// y = (A + diag d) xin on the CTA's tiles; red0 = y.u (u = rhat) or, NDOT == 2, red0 = y.xin, red1 = y.y
template <bool CLS, int NDOT>
__device__ __forceinline__ void persist_spmv(const PersistArgs &A, const DictParam &D, const double *xin, double *y, const double *u,
                                             const HaloWait &hw, int *status, double (*s_slab)[64], int ntile, int tid, int warp, int lane) {
    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        halo_wait(hw, tile, status);
        const int row_base = tile * kTile;
#pragma unroll 1
        for (int j = 0; j < kSlabsPerWarp; ++j) {
            const int slab = j * kCtaWarps + warp;
            const int row0 = row_base + slab * kSlab;
            double p0 = 0.0, p1 = 0.0;
            if (row0 < A.n) {                                      // warp-uniform
                const int row = row0 + lane;
                const bool active = row < A.n;
                // vectors written inside this kernel are read through L2 (ld.cg): L1 is not coherent across the grid barrier
                double sum = persist_rowsum<CLS>(A, D, xin, row0, row, active, lane);
                if (A.d) { if (active) sum = __dadd_rn(sum, __dmul_rn(__ldg(A.d + row), __ldcg(xin + row))); }
                if (active) __stcg(y + row, sum);
                if (NDOT == 1) p0 = active ? __dmul_rn(sum, __ldcg(u + row)) : 0.0;
                if (NDOT == 2) { p0 = active ? __dmul_rn(sum, __ldcg(xin + row)) : 0.0; p1 = active ? __dmul_rn(sum, sum) : 0.0; }
            }
            p0 = warp_butterfly(p0);
            if (NDOT == 2) p1 = warp_butterfly(p1);
            if (lane == 0) { s_slab[0][slab] = p0; if (NDOT == 2) s_slab[1][slab] = p1; }
        }
        const int nslab = min(kTileSlabs, (A.n - row_base + kSlab - 1) / kSlab);
        persist_tile_partial(A.rc, NDOT, tile, nslab, s_slab, warp, lane);
    }
}

// grid-wide barrier with memory ordering: the hardware cluster barrier when the whole grid is ONE thread-block cluster (tiny
// systems: <= 8 tiles), else the cooperative-groups grid barrier
template <bool CLUSTER>
__device__ __forceinline__ void persist_barrier() {
    if (CLUSTER) {
        __threadfence();
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        cg::this_grid().sync();
    }
}

template <bool CLS, bool CLUSTER>
__global__ void __launch_bounds__(kCtaThreads, 2) k_bicgstab_persist(const PersistArgs A, const __grid_constant__ DictParam D) {
    __shared__ DevScalars sc;                                      // this CTA's copy of the scalars (kept identical in every CTA)
    __shared__ double s_slab[kMaxQ][64];
    __shared__ double s_red[kMaxQ];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntile = (A.n + kTile - 1) / kTile;
    if (tid == 0) sc = *A.sc;
    __syncthreads();
    double *hist = blockIdx.x == 0 ? A.hist : nullptr;             // one writer of the residual history
    for (int it = 0; it < A.iters; ++it) {
        if (sc.status != ST_RUNNING) break;                        // identical in every CTA (and on every rank)
        // ---- A: p = fl(r + fl(beta * fl(p + fl(-omega * v))))   (pbicgstab.cu:668-672) ----
        {
            HaloPush hp = A.hp_p; hp.epoch += (unsigned long long)it;
            const double beta = sc.beta, momega = -sc.omega;
            for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
                const int row_base = tile * kTile;
#pragma unroll
                for (int j = 0; j < kSlabsPerWarp; ++j) {
                    const int row = row_base + j * kCtaThreads + tid;
                    if (row < A.n) {
                        double q = __dmul_rn(momega, __ldcg(A.v + row));
                        q = __dadd_rn(__ldcg(A.p + row), q);
                        q = __dmul_rn(beta, q);
                        q = __dadd_rn(__ldcg(A.r + row), q);
                        __stcg(A.p + row, q);
                        halo_store(hp, row, q);
                    }
                }
                halo_signal(hp, row_base, min(kTile, A.n - row_base));
            }
        }
        persist_barrier<CLUSTER>();
        // ---- B: v = A p ; rhat.v -> alpha   (:675-689) ----
        {
            HaloWait hw = A.hw_p; hw.epoch += (unsigned long long)it;
            persist_spmv<CLS, 1>(A, D, A.p, A.v, A.r0, hw, &sc.status, s_slab, ntile, tid, warp, lane);
        }
        persist_barrier<CLUSTER>();
        persist_reduce(A, 1, A.rc.p2p.epoch + 3ull * it, s_slab, s_red, &sc.status, warp, lane);
        if (tid == 0) { double red[kMaxQ] = {s_red[0], 0.0}; apply_phase(&sc, hist, PH_U_A, red); }
        __syncthreads();
        // ---- C: s = fl(r + fl(-alpha * v))   (:698-700) ----
        {
            HaloPush hp = A.hp_s; hp.epoch += (unsigned long long)it;
            const double malpha = -sc.alpha;
            for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
                const int row_base = tile * kTile;
#pragma unroll
                for (int j = 0; j < kSlabsPerWarp; ++j) {
                    const int row = row_base + j * kCtaThreads + tid;
                    if (row < A.n) {
                        const double q = __dadd_rn(__ldcg(A.r + row), __dmul_rn(malpha, __ldcg(A.v + row)));
                        __stcg(A.sv + row, q);
                        halo_store(hp, row, q);
                    }
                }
                halo_signal(hp, row_base, min(kTile, A.n - row_base));
            }
        }
        persist_barrier<CLUSTER>();
        // ---- D: t = A s ; t.s, t.t -> omega   (:703-710) ----
        {
            HaloWait hw = A.hw_s; hw.epoch += (unsigned long long)it;
            persist_spmv<CLS, 2>(A, D, A.sv, A.t, nullptr, hw, &sc.status, s_slab, ntile, tid, warp, lane);
        }
        persist_barrier<CLUSTER>();
        persist_reduce(A, 2, A.rc.p2p.epoch + 3ull * it + 1ull, s_slab, s_red, &sc.status, warp, lane);
        if (tid == 0) { double red[kMaxQ] = {s_red[0], s_red[1]}; apply_phase(&sc, hist, PH_U_B, red); }
        __syncthreads();
        // ---- E: h = fl(x + fl(alpha p)); x = fl(h + fl(omega s)); r = fl(s + fl(-omega t)); rhat.r, r.r   (:694-696,714-723) ----
        {
            const double alpha = sc.alpha, omega = sc.omega, momega = -omega;
            for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
                const int row_base = tile * kTile;
#pragma unroll 1
                for (int j = 0; j < kSlabsPerWarp; ++j) {
                    const int row = row_base + j * kCtaThreads + tid;
                    const bool act = row < A.n;
                    double rh = 0.0, rn = 0.0;
                    if (act) {
                        const double s_ = __ldcg(A.sv + row), t_ = __ldcg(A.t + row), p_ = __ldcg(A.p + row);
                        rh = __ldg(A.r0 + row);
                        const double h = __dadd_rn(__ldcg(A.x + row), __dmul_rn(alpha, p_));
                        __stcg(A.x + row, __dadd_rn(h, __dmul_rn(omega, s_)));
                        rn = __dadd_rn(s_, __dmul_rn(momega, t_));
                        __stcg(A.r + row, rn);
                    }
                    // thread tid of pass j owns row j*512 + tid: warp w covers slab j*16 + w — the update kernels' geometry
                    const double e0 = warp_butterfly(act ? __dmul_rn(rh, rn) : 0.0), e1 = warp_butterfly(act ? __dmul_rn(rn, rn) : 0.0);
                    if (lane == 0) { s_slab[0][j * kCtaWarps + warp] = e0; s_slab[1][j * kCtaWarps + warp] = e1; }
                }
                const int nslab = min(kTileSlabs, (A.n - row_base + kSlab - 1) / kSlab);
                persist_tile_partial(A.rc, 2, tile, nslab, s_slab, warp, lane);
            }
        }
        persist_barrier<CLUSTER>();
        persist_reduce(A, 2, A.rc.p2p.epoch + 3ull * it + 2ull, s_slab, s_red, &sc.status, warp, lane);
        if (tid == 0) { double red[kMaxQ] = {s_red[0], s_red[1]}; apply_phase(&sc, hist, PH_U_C, red); }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0) *A.sc = sc;                   // the host polls this
}

// ---- host side ---------------------------------------------------------------------------------------------------
// Largest shard (rows) the persistent loop is used for when option "persist" is automatic: below this the per-kernel loop
// is bound by kernel-boundary latency (measured: 40 us / iteration at n = 10^4, 160 us at 2.1 M rows per GPU).
constexpr long long kPersistAutoRows = 600000;
static long long persist_auto_rows() {
    static const long long v = [] { const char *e = getenv("CUDAMAT_PERSIST_ROWS"); return e && *e ? atoll(e) : kPersistAutoRows; }();
    return v;
}

bool persist_eligible(const cudamat_solver *s) {
    if (s->opt_persist == 0 || s->n <= 0) return false;
    if (s->comm && !comm_p2p(s)) return false;                     // NCCL transport: calls between kernels are needed
    if (s->rc.ngroup_loc > 64) return false;
    if (s->opt_persist > 0) return true;
    // the same decision on every rank: from the GLOBAL size and the world size
    const long long per_rank = s->comm ? s->n_global / std::max(1, s->rc.p2p.world) : s->n_global;
    return per_rank <= persist_auto_rows();
}

int launch_persist(cudamat_solver *s, const PersistLaunch &L) {
    PersistArgs A{};
    A.n = s->n; A.ia = s->d_ia; A.ja = s->d_ja; A.val = s->d_a;
    const bool cls = s->cls[1].ncls > 0 && s->cls[1].h_dict != nullptr;
    A.cls = cls ? s->cls[1].d_cls : nullptr;
    A.r0 = L.r0; A.r = L.r; A.v = L.v; A.p = L.p; A.sv = L.sv; A.t = L.t; A.x = L.x; A.d = L.d;
    A.rc = L.rc; A.sc = s->d_sc; A.hist = s->d_hist;
    A.hp_p = L.hp_p; A.hp_s = L.hp_s; A.hw_p = L.hw_p; A.hw_s = L.hw_s; A.red_flags = L.red_flags; A.iters = L.iters;
    static DictParam empty_dict{};
    const DictParam *D = cls ? s->cls[1].h_dict : &empty_dict;
    const int ntile = (s->n + kTile - 1) / kTile;
    const bool cluster = ntile <= 8 && !s->comm;                   // the whole grid as one thread-block cluster
    const void *kern = cluster ? (cls ? (const void *)k_bicgstab_persist<true, true> : (const void *)k_bicgstab_persist<false, true>)
                               : (cls ? (const void *)k_bicgstab_persist<true, false> : (const void *)k_bicgstab_persist<false, false>);
    void *args[] = {(void *)&A, (void *)D};
    if (cluster) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)ntile); cfg.blockDim = dim3(kCtaThreads); cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)ntile; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CM_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
        s->launches++;
        return CUDAMAT_OK;
    }
    if (s->persist_grid == 0) {
        int per_sm = 0, sms = 148;
        CM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCtaThreads, 0));
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        s->persist_grid = std::max(1, std::min(ntile, per_sm * sms));
    }
    CM_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)s->persist_grid), dim3(kCtaThreads), args, 0, s->stream));
    s->launches++;
    return CUDAMAT_OK;
}

}  // namespace cudamat
