// internal.cuh — shared device helpers and host-side state of libcudamat_b200.
//
// Arithmetic spec (mirrored bit for bit by oracle/oracle.c):
//   * every spec'd floating-point operation is an explicit __dmul_rn/__dadd_rn/__fma_rn/__ddiv_rn
//     (the file is also compiled with --fmad=false) so no contraction can change a rounding;
//   * reductions use the slab(32) -> tile(64 slabs) -> group(1024 tiles) -> final tree, all
//     boundaries aligned in the GLOBAL row index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/cudamat_b200.h"

namespace cudamat {

constexpr int kSlab        = 32;      // rows per slab (one per lane)
constexpr int kTileSlabs   = 64;      // slabs per tile
constexpr int kTile        = 2048;    // rows per tile == rows per CTA in reducing kernels
constexpr int kGroupTiles  = 1024;    // tiles per group
constexpr int kLongRow     = 32;      // rows longer than this use the interleaved row sum
constexpr int kCtaThreads  = 512;
constexpr int kCtaWarps    = 16;
constexpr int kSlabsPerWarp = kTileSlabs / kCtaWarps;   // 4
constexpr int kMaxQ        = 2;       // reduced quantities per kernel

// solver status (device side)
enum : int { ST_RUNNING = 0, ST_CONVERGED = 1, ST_BRK_OMEGA = 2, ST_BRK_NAN = 3, ST_MAXIT = 4,
              ST_COMM_TIMEOUT = 5 /* a peer's halo rows / partial sums never arrived: the host returns CUDAMAT_E_COMM */ };

// phases of the scalar recurrences executed by the last CTA of a reducing kernel
enum : int {
    PH_NONE = 0,
    PH_U_INIT,   // unprec: red0 = r.r            -> nrm0, rho' = red0, beta
    PH_U_A,      // unprec: red0 = r0.v           -> alpha
    PH_U_B,      // unprec: red0 = t.s, red1 = t.t -> omega
    PH_U_C,      // unprec: red0 = r0.r', red1 = r'.r' -> nrm, checks, rho rotate, beta
    PH_I_INIT,   // ilu0:   red0 = r.r            -> nrm0, rho = red0
    PH_I_A,      // ilu0:   red0 = rw.v           -> alpha
    PH_I_A2,     // ilu0:   red0 = r.r            -> nrm, check 1
    PH_I_B,      // ilu0:   red0 = t.r, red1 = t.t -> omega
    PH_I_C,      // ilu0:   red0 = rw.r, red1 = r.r -> nrm, i++, check 2, rho rotate, beta
    PH_STORE     // red -> sc->red only (kernel-level dot entry point)
};

struct DevScalars {
    double rho;        // unprec: rho (previous); ilu0: rho
    double rho_new;    // unprec: rho' of the coming iteration
    double alpha, omega, beta;
    double nrm0, nrm, tol;
    double red[kMaxQ];
    int iter;          // reference loop counter i
    int status;        // ST_*
    int half;          // entries written to hist
    int maxit;
    int hist_cap;
    int pad;
};

// ---- multi-GPU peer-memory plumbing (comm.cu) ---------------------------------------------
// Every rank maps its neighbours' work-vector arena and every rank's small exchange arena (CUDA IPC) and the
// kernels talk to them directly over NVLink: no NCCL call inside the iteration.
constexpr int kMaxPeer = 4;        // halo neighbours of one shard (row slabs: 2)
constexpr int kMaxWorld = 64;      // ranks
struct PeerRed {                   // where rank r keeps the partial sums gathered from all ranks
    double *gather[2];             // [parity][kMaxQ * stride]
    unsigned long long *flag;      // r's arrival flag of THIS rank's contribution (value = epoch)
};
struct P2PRed {                    // part of RedCtx; world == 0: disabled
    int world, me, my_off, my_cnt, stride;
    const PeerRed *peers;          // device array [world]
    unsigned long long epoch;      // of the reduction this kernel feeds; parity = epoch & 1
};
// halo rows of a vector are stored straight into the neighbours' copy of that vector by the kernel that
// produces it; the last boundary CTA of the grid then raises the neighbour's flag to `epoch`
struct HaloPush {
    int npeer;
    double *dst[kMaxPeer];                 // neighbour's halo slot for this rank's rows
    int first[kMaxPeer], cnt[kMaxPeer];    // local rows [first, first + cnt) go to dst[0 .. cnt)
    int ntiles[kMaxPeer];                  // CTAs (tiles) that overlap the range
    unsigned long long *flag[kMaxPeer];
    unsigned *local_cnt;                   // [kMaxPeer] boundary CTAs finished
    unsigned long long epoch;
};
// the consuming SpMV: CTAs whose rows reference halo columns wait until every neighbour's flag reached `epoch`
struct HaloWait {
    int nsrc;
    const unsigned long long *flag[kMaxPeer];
    const unsigned char *tile_wait;        // [tiles] 1: the tile has halo columns
    unsigned long long epoch;
};

// reduction context of one shard
struct RedCtx {
    double   *slab_part;   // [kMaxQ][slab_stride]  slab sums, indexed by LOCAL slab
    int slab_stride;
    int n_local;           // rows of this shard
    unsigned *done_cnt;    // [1] CTAs of k_reduce_finish that have finished their tiles
    double   *tile_part;   // [kMaxQ][tile_stride]  tile partials, indexed by LOCAL tile
    double   *slots;       // [kMaxQ][slot_stride]  group partials, indexed by GLOBAL group
    int ntile;             // local tiles
    int tile_stride;
    int slot_stride;
    int ngroup_loc;        // local groups
    int group0;            // global index of the first local group
    int nslots;            // global number of groups
    int ntile_global;      // global number of tiles (exch_level 1)
    // multi-GPU: where this shard deposits its contribution for the zero-padded allreduce
    int exch_level;        // 0 none; 1 tile partials (global tile index); 2 group partials (== slots)
    int tile0;             // global index of the first local tile
    int exch_stride;
    double *exch;          // level 1: [kMaxQ][exch_stride] indexed by global tile
    P2PRed p2p;            // peer-memory all-gather of the partials instead of an NCCL allreduce
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL). Every loop kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and runs
//     [prefetch of operands that NO kernel of the solve writes (the CSR arrays)]  pdl_sync()  [everything else]
// pdl_sync() = griddepcontrol.launch_dependents (the successor may be scheduled as soon as every CTA of this
// grid has started, i.e. while this grid drains) + griddepcontrol.wait (predecessor complete, writes visible).
// Only solve-constant data may be touched before pdl_sync(): with this order a successor can start before the
// predecessor's predecessor has finished.
__device__ __forceinline__ void pdl_sync() {
#ifdef CUDAMAT_PDL_WAIT_FIRST
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#else
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_prologue() { pdl_sync(); }

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void halo_store(const HaloPush &hp, int row, double v) {
#pragma unroll
    for (int p = 0; p < kMaxPeer; ++p)
        if (p < hp.npeer && row >= hp.first[p] && row < hp.first[p] + hp.cnt[p]) hp.dst[p][row - hp.first[p]] = v;
}
// end of a producing CTA (all threads): publish this tile's halo stores; the last boundary CTA raises the flag
__device__ __forceinline__ void halo_signal(const HaloPush &hp, int row_base, int rows_here) {
    if (hp.npeer == 0) return;
    __syncthreads();
    const int p = threadIdx.x;
    if (p < hp.npeer && row_base < hp.first[p] + hp.cnt[p] && row_base + rows_here > hp.first[p]) {
        __threadfence_system();
        const unsigned old = atomicAdd(hp.local_cnt + p, 1u);
        if (old == (unsigned)(hp.ntiles[p] - 1)) {
            hp.local_cnt[p] = 0u;
            __threadfence_system();
            st_release_sys_u64(hp.flag[p], hp.epoch);
        }
    }
}
// A peer that never delivers must neither hang the GPU nor kill the CUDA context: after the spin limit the waiting
// thread records ST_COMM_TIMEOUT in the solver status (every later kernel of the solve returns at entry, the host poll
// turns it into CUDAMAT_E_COMM) and the kernel carries on with whatever the halo holds.
constexpr unsigned kPeerSpinLimit = 1u << 26;
__device__ __forceinline__ void halo_wait(const HaloWait &hw, int tile, int *status) {
    if (hw.nsrc == 0 || !hw.tile_wait[tile]) return;
    if ((int)threadIdx.x < hw.nsrc) {
        unsigned spins = 0;
        while (ld_acquire_sys_u64(hw.flag[threadIdx.x]) < hw.epoch)
            if (++spins > kPeerSpinLimit) { if (status) atomicExch(status, ST_COMM_TIMEOUT); break; }
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_butterfly(double v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, s));
    return v;
}

// R(): lane-strided chains over m values (read through L2), then butterfly
__device__ __forceinline__ double warp_reduce_values_cg(const double *v, int m, int lane) {
    double acc = 0.0;
    for (int j = lane; j < m; j += 32) acc = __dadd_rn(acc, __ldcg(v + j));
    return warp_butterfly(acc);
}
__device__ __forceinline__ double warp_reduce_values_smem(const double *v, int m, int lane) {
    double acc = 0.0;
    for (int j = lane; j < m; j += 32) acc = __dadd_rn(acc, v[j]);
    return warp_butterfly(acc);
}

__device__ __forceinline__ void hist_push(DevScalars *sc, double *hist, double v) {
    if (hist && sc->half < sc->hist_cap) hist[sc->half] = v;
    sc->half += 1;
}

// scalar recurrences; executed by one thread. Forms follow pbicgstab.cu (cited per phase).
__device__ __forceinline__ void apply_phase(DevScalars *sc, double *hist, int phase, const double *red) {
    if (sc->status == ST_COMM_TIMEOUT) return;          // sticky: the sums below were formed from missing peer data
    switch (phase) {
    case PH_STORE:
        sc->red[0] = red[0]; sc->red[1] = red[1];
        break;
    case PH_U_INIT: {                                   // pbicgstab.cu:655,665-666 (first pass)
        double n0 = sqrt(red[0]);
        sc->nrm0 = n0; sc->nrm = n0;
        hist_push(sc, hist, n0);
        sc->rho_new = red[0];                           // dot(r0,r) with r0 == r
        sc->beta = __dmul_rn(__ddiv_rn(sc->rho_new, sc->rho), __ddiv_rn(sc->alpha, sc->omega));
        if (sc->maxit <= 0) sc->status = ST_MAXIT;
    } break;
    case PH_U_A:                                        // :688-689
        sc->alpha = __ddiv_rn(sc->rho_new, red[0]);
        break;
    case PH_U_B:                                        // :708-710
        sc->omega = __ddiv_rn(red[0], red[1]);
        break;
    case PH_U_C: {                                      // :723-747
        double nr = sqrt(red[1]);
        sc->nrm = nr;
        hist_push(sc, hist, nr);
        sc->iter += 1;
        double om = sc->omega;
        if (nr < __dmul_rn(sc->tol, sc->nrm0)) sc->status = ST_CONVERGED;
        else if (isnan(om)) sc->status = ST_BRK_NAN;
        else if (fabs(om) < 1e-5) sc->status = ST_BRK_OMEGA;
        else if (sc->iter >= sc->maxit) sc->status = ST_MAXIT;
        sc->rho = sc->rho_new;                          // :748
        sc->rho_new = red[0];                           // :665 of the next pass
        sc->beta = __dmul_rn(__ddiv_rn(sc->rho_new, sc->rho), __ddiv_rn(sc->alpha, sc->omega));
    } break;
    case PH_I_INIT: {                                   // :74, :81 (first pass)
        double n0 = sqrt(red[0]);
        sc->nrm0 = n0; sc->nrm = n0;
        hist_push(sc, hist, n0);
        sc->rho = red[0];                               // dot(rw,r) with rw == r
        if (sc->maxit <= 0) sc->status = ST_MAXIT;
    } break;
    case PH_I_A:                                        // :106-107
        sc->alpha = __ddiv_rn(sc->rho, red[0]);
        break;
    case PH_I_A2: {                                     // :111-118  (break without i++)
        double nr = sqrt(red[0]);
        sc->nrm = nr;
        hist_push(sc, hist, nr);
        if (nr < __dmul_rn(sc->tol, sc->nrm0)) sc->status = ST_CONVERGED;
    } break;
    case PH_I_B:                                        // :135-137
        sc->omega = __ddiv_rn(red[0], red[1]);
        break;
    case PH_I_C: {                                      // :142-151, then :80-84 of the next pass
        double nr = sqrt(red[1]);
        sc->nrm = nr;
        hist_push(sc, hist, nr);
        sc->iter += 1;
        if (nr < __dmul_rn(sc->tol, sc->nrm0)) sc->status = ST_CONVERGED;
        else if (sc->iter >= sc->maxit) sc->status = ST_MAXIT;
        double rhop = sc->rho;
        sc->rho = red[0];
        sc->beta = __dmul_rn(__ddiv_rn(sc->rho, rhop), __ddiv_rn(sc->alpha, sc->omega));
    } break;
    default: break;
    }
}

// Finishing kernel, sharded handle: copy this rank's partial sums (its tiles or groups) into every rank's gather
// array over NVLink, then raise this rank's arrival flag there (all threads of the CTA call)
__device__ __forceinline__ void p2p_push(const RedCtx &rc, const double *local, int nq) {
    const P2PRed &pp = rc.p2p;
    const int par = (int)(pp.epoch & 1ull);
    for (int r = 0; r < pp.world; ++r) {
        double *dst = pp.peers[r].gather[par];
        for (int q = 0; q < nq; ++q)
            for (int i = threadIdx.x; i < pp.my_cnt; i += blockDim.x)
                dst[(size_t)q * pp.stride + pp.my_off + i] = __ldcg(local + (size_t)q * pp.stride + pp.my_off + i);
    }
    __syncthreads();
    if ((int)threadIdx.x < pp.world) {
        __threadfence_system();
        st_release_sys_u64(pp.peers[threadIdx.x].flag, pp.epoch);
    }
}

// Reducing kernels only publish SLAB sums: one product per lane (inactive lanes pass +0.0), butterfly, lane 0
// stores the sum at the slab's LOCAL index (tile * 64 + slab in tile).  No CTA barrier, no fence, no atomic in
// the bandwidth kernels: a gpu-scope fence per CTA invalidates the SM's L1 under the co-resident CTAs' x gathers
// and the barrier in front of a CTA-level tail parked 27 % of the warp samples (profiles/r1c_*).  Tiles, groups,
// the final sum and the scalar recurrence are formed by k_reduce_finish, launched right behind.
// M (= 4 or 8) independent butterflies in 1 + log2(M) .. shuffles instead of 5 M: at each of the first log2(M) steps
// a lane keeps one value of a pair and hands the other one to its partner, so the number of live values halves while
// the partner distance halves.  Every addition is `own + partner's` exactly as in warp_butterfly and IEEE addition is
// commutative, so the result is bit-identical to M separate warp_butterfly calls.  On return the sum of v[idx] is
// held by all lanes that share `idx` = (lane>>4 & 1) + 2 (lane>>3 & 1) [+ 4 (lane>>2 & 1)].
// one step for one pair: lanes with bit `s` clear keep a, the others keep b; result = own + partner's (distance s)
__device__ __forceinline__ double packed_pair(double a, double b, int s, int lane) {
    const bool up = (lane & s) != 0;
    const double send = up ? a : b, keep = up ? b : a;
    return __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, s));
}
template <int M>
__device__ __forceinline__ double packed_butterfly(double (&v)[M], int lane, int &idx) {
    static_assert(M == 4 || M == 8, "packed_butterfly: 4 or 8 values");
    int s = 16;
#pragma unroll
    for (int m = M; m > 1; m >>= 1, s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < m / 2; ++k) {
            const double a = v[2 * k], b = v[2 * k + 1];
            const double send = up ? a : b, keep = up ? b : a;
            v[k] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, s));
        }
    }
#pragma unroll
    for (; s >= 1; s >>= 1) v[0] = __dadd_rn(v[0], __shfl_xor_sync(0xffffffffu, v[0], s));
    idx = ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1) + (M == 8 ? 4 * ((lane >> 2) & 1) : 0);
    return v[0];
}

__device__ __forceinline__ void slab_deposit(const RedCtx &rc, int q, int slab_local, double prod, int lane) {
    const double s = warp_butterfly(prod);
    if (lane == 0) __stcg(rc.slab_part + (size_t)q * rc.slab_stride + slab_local, s);
}

}  // namespace cudamat
