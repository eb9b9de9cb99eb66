// sweepblk.cu — block-wavefront triangular sweeps for ILU(0) factors of 7-point grid stencils
// (replaces cusparseDcsrsv_solve, pbicgstab.cu:94,98,123,127, when the matrix has the grid structure the MARCH analysis found).
//
// The natural-order 7-point factor of an N^3 grid has 3N - 2 levels (wavefronts i + j + k); every level hop of the generic
// sync-free sweep (ilu0.cu) is a dependent L2 round trip of ~1.1-1.6 us: 766 hops = 1.25 ms per sweep at 256^3, 4 sweeps per
// iteration (profiles/r1b).  Here the grid is cut into 16^3 blocks (4096 rows).  Inside a block the 46 wavefronts are walked by
// ONE CTA (one per SM, 211 KB of shared memory) with everything a wavefront touches in shared memory — the block's factor
// records (TMA bulk copy, issued a whole block ahead), its right-hand side, the solved values — and a CTA barrier per wavefront
// instead of a global round trip (~170 cycles per wavefront for L, ~340 for U with its IEEE division, measured).  Only block-
// to-block dependencies (3 faces) go through global memory: blocks are handed out in block-wavefront order by a ticket, a
// block waits for the done-flags of its three predecessor blocks, pulls their faces of the output vector into its halo layer,
// and publishes its three outgoing faces + flag before the rest of its values.
// Measured (tools/sptrsv_bench.py, B200): 256^3 L 0.43 ms / U 0.63 ms against 1.26 / 1.31 ms of the generic sweeps; 64^3 0.10 /
// 0.14 ms against 0.21 / 0.22 ms; grids whose edges are not multiples of 16 (partial blocks): 100^3 0.17 / 0.26 against 0.33 /
// 0.36 ms, 250^3 0.53 / 0.75 against 1.21 / 1.26 ms.  CUDAMAT_SWEEP_DEBUG=1 prints the SM cycles per phase of a block.
//
// Arithmetic is the spec's (oracle orc_sptrsv_*): per row acc = rhs; acc = fma(-M_ik, y_k, acc) over the
// row's entries in ascending column order (L: -D, -a, -1; U: +1, +a, +D), U ends with one IEEE division — bit-identical to
// the other sweep kernels.  The coefficients come from a block-ordered copy of the factor (one 32-byte record per row, in the
// order the block walks its rows), built once after the factorisation.
#include "solver.h"
#include "rowfuncs.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace cudamat {

constexpr int kSB = 16;                                            // block edge
constexpr int kSBRows = kSB * kSB * kSB;                           // 4096 rows per block
constexpr int kSBLevels = 3 * kSB - 2;                             // 46 wavefronts inside a block
constexpr int kSBH = kSB + 1;                                      // edge of the shared-memory cube incl. one halo layer
// pitches of the cube: the cells of a wavefront (x + y + z = l) must spread over the shared-memory banks.  With pitches 17 / 289
// every cell of a wavefront falls into the SAME 8-byte bank (17 = 289 = 1 mod 16): 16-way conflicts on every access, 546 cycles
// per wavefront measured.  18 / 319 is the best pair of a brute-force search over the kernel's cell order (1.8 passes per half
// warp on average, 1.0 is ideal).
constexpr int kSBPY = 18, kSBPZ = 319;
constexpr int kSBThreads = 256;                                    // >= the widest wavefront (192 cells)
constexpr int kSBWalkers = 192;                                    // threads that walk the wavefronts (named barrier 1)

struct SweepTable {                                                // cells of a block in wavefront order (kernel parameter, 8.4 KB)
    unsigned short cell[kSBRows];                                  // lx | ly << 4 | lz << 8
    unsigned short lvl_ptr[kSBLevels + 2];
};

struct BlockSweep {
    int nbx, nby, nbz, nblk;
    int nx, ny, nz;                                                // line length a, lines per plane D / a, planes
    int *d_order[2] = {nullptr, nullptr};                          // block ids in block-wavefront order (L ascending, U descending)
    double *d_coef[2] = {nullptr, nullptr};                        // [blk][2][pos][2]: L: M(-D), M(-a) | M(-1), -; U: M(+1), M(+a) | M(+D), diag
    unsigned char *d_pres[2] = {nullptr, nullptr};                 // [blk][pos] presence bits of the three entries
    unsigned short *d_inv[2] = {nullptr, nullptr};                 // wavefront position of every natural cell index of a block
    int *d_flag = nullptr;                                         // [nblk] epoch of the last sweep that finished the block
    unsigned *d_ticket = nullptr;
    int epoch = 0;
    int grid = 0;
    SweepTable *h_table = nullptr;
};

// block-ordered factor records: thread per (block, position).  Blocks at the far faces of a grid whose edges are not multiples
// of 16 are partial: positions outside the grid get an all-zero record without the `valid` bit (8) and are never stored.
__global__ void k_sblk_records(int nblk, int nbx, int nby, int nx, int ny, int nz, const __grid_constant__ SweepTable T, const int *ia,
                               const unsigned char *tmask, const double *M, double *coef_l, unsigned char *pres_l, double *coef_u,
                               unsigned char *pres_u) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nblk * kSBRows) return;
    const int blk = (int)(t / kSBRows), pos = (int)(t % kSBRows);
    const int bx = blk % nbx, by = (blk / nbx) % nby, bz = blk / (nbx * nby);
    for (int u = 0; u < 2; ++u) {
        const unsigned c = T.cell[pos];
        int lx = c & 15, ly = (c >> 4) & 15, lz = c >> 8;
        if (u) { lx = kSB - 1 - lx; ly = kSB - 1 - ly; lz = kSB - 1 - lz; }           // U walks the block from the far corner
        const int gx = bx * kSB + lx, gy = by * kSB + ly, gz = bz * kSB + lz;
        double c4[4] = {0.0, 0.0, 0.0, u ? 1.0 : 0.0};
        unsigned pres = 0;
        if (gx < nx && gy < ny && gz < nz) {
            const long long g = ((long long)gz * ny + gy) * nx + gx;
            const unsigned m = tmask[g];                          // pattern (-D, -a, -1, 0, +1, +a, +D)
            const int s = ia[g];
            pres = 8u;
            for (int e = 0; e < 3; ++e) {
                const int q = u ? 4 + e : e;                       // ascending column order on either side of the diagonal
                if (m & (1u << q)) { c4[e] = M[s + __popc(m & ((1u << q) - 1u))]; pres |= 1u << e; }
            }
            if (u) c4[3] = M[s + __popc(m & 7u)];                 // the diagonal
        }
        double *dst = (u ? coef_u : coef_l) + (size_t)blk * kSBRows * 4 + (size_t)pos * 2;     // two planes of double2 per block
        dst[0] = c4[0]; dst[1] = c4[1]; dst[2 * kSBRows] = c4[2]; dst[2 * kSBRows + 1] = c4[3];
        (u ? pres_u : pres_l)[t] = (unsigned char)pres;
    }
}
// the kernel resolves a row's three predecessors by grid position: an entry that crosses a grid face (periodic stencils, a -1
// entry at the start of a line, ...) would be read from the wrong place — such matrices keep the generic sweeps
__global__ void k_sblk_check(int n, int nx, int ny, int nz, const unsigned char *tmask, int *bad) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int gx = row % nx, gy = (row / nx) % ny, gz = row / (nx * ny);
    const unsigned m = tmask[row];
    const bool ok = !((m & 1u) && gz == 0) && !((m & 2u) && gy == 0) && !((m & 4u) && gx == 0) && (m & 8u) &&
                    !((m & 16u) && gx == nx - 1) && !((m & 32u) && gy == ny - 1) && !((m & 64u) && gz == nz - 1);
    if (!ok) *bad = 1;
}

__device__ __forceinline__ uint32_t sb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int ld_relaxed_gpu_i32(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Shared memory of one CTA (one CTA per SM): the block's factor records, right-hand side and presence bits in WAVEFRONT order,
// so that the cell a thread owns in wavefront l sits at lvl_ptr[l] + tid in all three; the solved values Y as a 17^3 cube with
// one halo layer; the cube index of every wavefront position.
struct SweepSmem {
    double C[kSBRows * 4];                                         // 128 KB, TMA bulk copy from the block-ordered factor: two planes of double2
    double R[kSBRows];                                             // 32 KB
    double Y[kSBH * kSBPZ + 1];                                    // 43 KB
    unsigned char P[kSBRows];                                      // 4 KB, TMA bulk copy
    unsigned short li[kSBRows];                                    // 8 KB: byte offset of the position's cell in Y
    unsigned short lp[kSBLevels + 4];                              // wavefront pointers (an indexed constant-bank load costs ~150 cycles)
    unsigned long long bar;                                        // mbarrier of the bulk copies
    int next_blk;                                                  // block of the CTA's next ticket, -1 = none left
};

constexpr size_t kSBSmem = sizeof(SweepSmem);                     // 211 KB: one CTA per SM

template <bool UPPER>
__global__ void __launch_bounds__(kSBThreads, 1) k_sptrsv_blocked(int nblk, int nbx, int nby, int nbz, int nx, int ny, int nz, const int *order,
                                                                   const double *coef, const unsigned char *pres, const double *rhs,
                                                                   double *out, int *flag, unsigned *ticket, int epoch, const int *status,
                                                                   const __grid_constant__ SweepTable T, const unsigned short *inv,
                                                                   long long *dbg) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    SweepSmem &S = *reinterpret_cast<SweepSmem *>(s_raw);         // (no pointer arithmetic on integers: keeps LDS / STS)
    if (status && *status != ST_RUNNING) return;
    const int tid = threadIdx.x;
    // L: cell (lx, ly, lz) lives at (lx + 1, ly + 1, lz + 1), the halo layer at index 0, predecessors at -1 / -kSBH / -kSBH^2
    // U: cell lives at (lx, ly, lz), the halo layer at index kSB, predecessors at +1 / +pitch_y / +pitch_z
    constexpr int OFF = UPPER ? 0 : 1;
    constexpr int SX = UPPER ? 1 : -1, SY = UPPER ? kSBPY : -kSBPY, SZ = UPPER ? kSBPZ : -kSBPZ;
    constexpr int kPer = kSBRows / kSBThreads;                     // 16 cells per thread in the natural-order passes
    constexpr uint32_t kCBytes = sizeof(double) * 4 * kSBRows, kPBytes = kSBRows;
    const long long plane = (long long)nx * ny;
    for (int i = tid; i < kSBRows; i += kSBThreads) {
        const unsigned c = T.cell[i];
        int lx = c & 15, ly = (c >> 4) & 15, lz = c >> 8;
        if (UPPER) { lx = kSB - 1 - lx; ly = kSB - 1 - ly; lz = kSB - 1 - lz; }
        S.li[i] = (unsigned short)(8 * ((lz + OFF) * kSBPZ + (ly + OFF) * kSBPY + (lx + OFF)));     // byte offset into Y
    }
    if (tid < kSBLevels + 4) S.lp[tid] = T.lvl_ptr[min(tid, kSBLevels + 1)];
    unsigned short inv_r[kPer];                                    // wavefront position of natural cell tid + 256 k (same for every block)
#pragma unroll
    for (int k = 0; k < kPer; ++k) inv_r[k] = __ldg(inv + tid + k * kSBThreads);
    auto fetch_records = [&](int blk) {                            // one thread: the block's records and presence bits by TMA
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_u32(&S.bar)), "r"(kCBytes + kPBytes) : "memory");
        constexpr uint32_t kPiece = kCBytes / 4;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sb_smem_u32(reinterpret_cast<unsigned char *>(S.C) + q * kPiece)),
                           "l"(reinterpret_cast<const unsigned char *>(coef + (size_t)blk * kSBRows * 4) + q * kPiece), "r"(kPiece),
                           "r"(sb_smem_u32(&S.bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sb_smem_u32(S.P)), "l"(pres + (size_t)blk * kSBRows), "r"(kPBytes), "r"(sb_smem_u32(&S.bar)) : "memory");
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_smem_u32(&S.bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int tk0 = (int)atomicAdd(ticket, 1u);
        int b0 = -1;
        if (tk0 < nblk) { b0 = __ldg(order + tk0); fetch_records(b0); }
        S.next_blk = b0;
    }
    __syncthreads();
    int blk = S.next_blk;
    uint32_t phase = 0;
    while (blk >= 0) {
        long long ts[6]; if (dbg && tid == 0) ts[0] = clock64();
        const int bx = blk % nbx, by = (blk / nbx) % nby, bz = blk / (nbx * nby);
        const long long g0 = ((long long)(bz * kSB) * ny + by * kSB) * nx + bx * kSB;       // first cell of the block
        const int ex = min(kSB, nx - bx * kSB), ey = min(kSB, ny - by * kSB), ez = min(kSB, nz - bz * kSB);     // < 16: a partial block at a far face
        // ---- independent of the predecessors: the right-hand side, coalesced from the vector, scattered to wavefront order ----
        {
            double rv[kPer];
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = tid + k * kSBThreads;
                const bool in = (i & 15) < ex && ((i >> 4) & 15) < ey && (i >> 8) < ez;
                rv[k] = in ? __ldg(rhs + g0 + (i >> 8) * plane + (long long)((i >> 4) & 15) * nx + (i & 15)) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kPer; ++k) S.R[inv_r[k]] = rv[k];
        }
        if (dbg && tid == 0) ts[1] = clock64();
        // ---- wait for the three predecessor blocks, pull their faces into the halo layer ----
        if (tid < 3) {
            const int nb = tid == 0 ? (UPPER ? (bx + 1 < nbx ? blk + 1 : -1) : (bx > 0 ? blk - 1 : -1))
                         : tid == 1 ? (UPPER ? (by + 1 < nby ? blk + nbx : -1) : (by > 0 ? blk - nbx : -1))
                                    : (UPPER ? (bz + 1 < nbz ? blk + nbx * nby : -1) : (bz > 0 ? blk - nbx * nby : -1));
            if (nb >= 0) {
                unsigned spins = 0;
                while (ld_relaxed_gpu_i32(flag + nb) != epoch) { if (++spins > (1u << 24)) __trap(); }      // never hang on a broken schedule
                asm volatile("fence.acq_rel.gpu;" ::: "memory");   // acquire: the faces below are read after the flag
            }
        }
        __syncthreads();
        if (dbg && tid == 0) ts[2] = clock64();
        {
            // 16 x 16 cells per face; the face of the x-neighbour is strided, the other two are runs of 16
            const int u = tid & 15, w = tid >> 4;
            const int hx = UPPER ? kSB : -1;                      // coordinate of the halo layer relative to the block
            const bool has_x = UPPER ? bx + 1 < nbx : bx > 0, has_y = UPPER ? by + 1 < nby : by > 0, has_z = UPPER ? bz + 1 < nbz : bz > 0;
            double hv[3] = {0.0, 0.0, 0.0};
            if (has_x && u < ey && w < ez) hv[0] = __ldcg(out + g0 + w * plane + (long long)u * nx + hx);
            if (has_y && u < ex && w < ez) hv[1] = __ldcg(out + g0 + w * plane + (long long)hx * nx + u);
            if (has_z && u < ex && w < ey) hv[2] = __ldcg(out + g0 + hx * plane + (long long)w * nx + u);
            if (has_x) S.Y[(w + OFF) * kSBPZ + (u + OFF) * kSBPY + (hx + OFF)] = hv[0];
            if (has_y) S.Y[(w + OFF) * kSBPZ + (hx + OFF) * kSBPY + (u + OFF)] = hv[1];
            if (has_z) S.Y[(hx + OFF) * kSBPZ + (w + OFF) * kSBPY + (u + OFF)] = hv[2];
        }
        {                                                          // the block's records have arrived (issued a whole block earlier)
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(sb_smem_u32(&S.bar)), "r"(phase) : "memory");
            phase ^= 1u;
        }
        __syncthreads();
        if (dbg && tid == 0) ts[3] = clock64();
        // ---- the block's wavefronts: one CTA barrier per level; everything that does not depend on Y is read a level ahead ----
        // A thread's chain per wavefront is: barrier -> 3 LDS of solved neighbours -> 3 dependent DFMA (-> division) -> STS -> barrier.
        // Everything else (the next wavefront's record, right-hand side, cube offset, presence bits) is loaded a wavefront ahead into
        // the other of two register sets, in the shadow of that chain.  Absent entries have coefficient +0.0 and read y = +0.0:
        // fma(-0.0, +0.0, acc) = acc + (-0.0) = acc for every acc, so the FMA needs no predicate.  Threads beyond the wavefront run
        // on a clamped position and only skip the store.
        struct Ops { double2 a, b; double r, rcp; unsigned li, p; };
        const unsigned y_a = opaque_smem_addr(S.Y);                // explicit shared addresses on the chain (no S2R + LEA per access)
        auto load_ops = [&](int q, Ops &o) {
            o.a = reinterpret_cast<const double2 *>(S.C)[q]; o.b = reinterpret_cast<const double2 *>(S.C)[kSBRows + q];
            o.r = S.R[q]; o.li = S.li[q]; o.p = S.P[q];
        };
        // U: the diagonal of a wavefront is read TWO wavefronts ahead, so that the reciprocal half of its division (MUFU + 5
        // DFMA) has its operand in a register a whole wavefront before it is needed and issues in the shadow of the solved-value
        // loads instead of in front of the quotient (in-order issue: whatever waits in front of the chain's next instruction is
        // on the chain).  Tried and slower: the idle warps 6-7 computing the reciprocals into a shared-memory ring (U wavefront
        // 342 -> 400 cycles: the helpers' own chain plus the 8-warp barrier).
        auto load_diag = [&](int q) { return S.C[2 * (kSBRows + q) + 1]; };
        int p0 = 0, p1 = S.lp[1], p2 = S.lp[2];
        double dg_ahead = 1.0;
        auto step = [&](const Ops &cur, Ops &nxt, int l) {
            const int p3 = S.lp[l + 3];                            // read a wavefront before it is needed
            const unsigned ya = y_a + cur.li;
            double y0 = 0.0, y1 = 0.0, y2 = 0.0;
            if (cur.p & 1u) y0 = lds_f64((unsigned)((int)ya + 8 * (UPPER ? SX : SZ)));
            if (cur.p & 2u) y1 = lds_f64((unsigned)((int)ya + 8 * SY));
            if (cur.p & 4u) y2 = lds_f64((unsigned)((int)ya + 8 * (UPPER ? SZ : SX)));
            if (UPPER) {
                nxt.rcp = div_prepare(dg_ahead);                   // diagonal of wavefront l + 1, in a register since wavefront l - 1
                asm volatile("" : "+d"(nxt.rcp));                  // here, not sunk into the dependent chain
                dg_ahead = load_diag(min(p2 + tid, kSBRows - 1));  // diagonal of wavefront l + 2
            }
            load_ops(min(p1 + tid, kSBRows - 1), nxt);
            const bool mine = tid < p1 - p0 && (cur.p & 8u);      // bit 3: the cell exists (partial blocks)
            double acc = cur.r;                                    // ascending columns — U: +1, +a, +D, then the division; L: -D, -a, -1
            acc = __fma_rn(-cur.a.x, y0, acc);
            acc = __fma_rn(-cur.a.y, y1, acc);
            acc = __fma_rn(-cur.b.x, y2, acc);
            if (UPPER && mine) acc = div_finish(acc, cur.b.y, cur.rcp);
            if (mine) sts_f64(ya, acc);
            asm volatile("bar.sync 1, %0;" ::"n"(kSBWalkers) : "memory");
            p0 = p1; p1 = p2; p2 = p3;
        };
        if (tid < kSBWalkers) {                                    // the widest wavefront has 192 cells: warps 6-7 only copy
            Ops oa, ob;
            load_ops(min(tid, kSBRows - 1), oa);
            oa.rcp = 1.0; ob.rcp = 1.0;
            if (UPPER) { oa.rcp = div_prepare(oa.b.y); dg_ahead = load_diag(min(p1 + tid, kSBRows - 1)); }
#pragma unroll 1
            for (int l = 0; l < kSBLevels; l += 2) {               // 46 wavefronts: an even number
                step(oa, ob, l);
                step(ob, oa, l + 1);
            }
        }
        __syncthreads();
        if (dbg && tid == 0) ts[4] = clock64();
        // ---- publish.  Successors need only the three outgoing faces: those first, then the done-flag, then the rest ----
        {
            const int u = tid & 15, w = tid >> 4;
            const int fx = UPPER ? 0 : kSB - 1;                   // the face the successors read
            if (fx < ex && u < ey && w < ez) __stcg(out + g0 + w * plane + (long long)u * nx + fx, S.Y[(w + OFF) * kSBPZ + (u + OFF) * kSBPY + (fx + OFF)]);
            if (u < ex && fx < ey && w < ez) __stcg(out + g0 + w * plane + (long long)fx * nx + u, S.Y[(w + OFF) * kSBPZ + (fx + OFF) * kSBPY + (u + OFF)]);
            if (u < ex && w < ey && fx < ez) __stcg(out + g0 + fx * plane + (long long)w * nx + u, S.Y[(fx + OFF) * kSBPZ + (w + OFF) * kSBPY + (u + OFF)]);
        }
        __syncthreads();                                           // the release below is cumulative over the CTA's face stores
        if (tid == 0) {
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag + blk), "r"(epoch) : "memory");
            if (dbg) { ts[5] = clock64(); for (int q = 0; q < 6; ++q) dbg[(size_t)blk * 6 + q] = ts[q]; }
            // the records buffer is free (every thread is past the wavefronts): next ticket, next block's TMA copies
            const int tn = (int)atomicAdd(ticket, 1u);
            int bn = -1;
            if (tn < nblk) { bn = __ldg(order + tn); fetch_records(bn); }
            S.next_blk = bn;
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kSBThreads;
            const int lx = i & 15, ly = (i >> 4) & 15, lz = i >> 8;
            if (lx < ex && ly < ey && lz < ez) __stcg(out + g0 + lz * plane + (long long)ly * nx + lx, S.Y[(lz + OFF) * kSBPZ + (ly + OFF) * kSBPY + (lx + OFF)]);
        }
        __syncthreads();
        blk = S.next_blk;                                          // (rewritten only after the next block's wavefronts)
    }
}

void sweepblk_release(cudamat_solver *s) {
    BlockSweep *B = s->bsweep;
    if (!B) return;
    for (int u = 0; u < 2; ++u) { dev_free(B->d_order[u]); dev_free(B->d_coef[u]); dev_free(B->d_pres[u]); dev_free(B->d_inv[u]); }
    dev_free(B->d_flag); dev_free(B->d_ticket);
    delete B->h_table;
    delete B;
    s->bsweep = nullptr;
}

// builds the block plan after the factorisation when the class analysis (rowclass.cu) found the 7-point grid structure —
// superset pattern (-D, -a, -1, 0, +1, +a, +D) with n = nx ny nz, a = nx, D = nx ny — and no entry crosses a grid face;
// otherwise the generic sweeps stay.  Grid edges need not be multiples of the block edge (partial blocks at the far faces).
// Systems the single-CTA shared-memory sweeps hold (<= 25 600 rows) are left to them.
int sweepblk_plan(cudamat_solver *s) {
    sweepblk_release(s);
    const RowClasses &C = s->cls[1];
    if (!s->opt_sptrsv_blocked || s->comm || s->d_perm || !C.d_tmask || !C.h_tdict || C.h_tdict->sup_len != 7 || s->n <= 25600) return CUDAMAT_OK;
    const int *so = C.h_tdict->sup_off;
    const long long nx = so[5], D = so[6];
    if (nx < 2 || D < 2 * nx || so[0] != -D || so[1] != -nx || so[2] != -1 || so[3] != 0 || so[4] != 1 || D % nx != 0 || s->n % D != 0) return CUDAMAT_OK;
    const long long ny = D / nx, nz = s->n / D;
    if (ny < 2 || nz < 2 || nx * ny * nz != s->n) return CUDAMAT_OK;
    {
        int *d_bad = nullptr, bad = 1;
        CM_CUDA(dev_alloc((void **)&d_bad, sizeof(int)));
        CM_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), s->stream));
        k_sblk_check<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->n, (int)nx, (int)ny, (int)nz, C.d_tmask, d_bad);
        s->launches++;
        CM_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
        dev_free(d_bad);
        if (bad) return CUDAMAT_OK;
    }
    BlockSweep *B = new BlockSweep();
    s->bsweep = B;
    B->nx = (int)nx; B->ny = (int)ny; B->nz = (int)nz;
    B->nbx = (int)((nx + kSB - 1) / kSB); B->nby = (int)((ny + kSB - 1) / kSB); B->nbz = (int)((nz + kSB - 1) / kSB); B->nblk = B->nbx * B->nby * B->nbz;
    // cells of a block in wavefront order
    B->h_table = new SweepTable();
    std::vector<int> cnt(kSBLevels + 1, 0);
    for (int c = 0; c < kSBRows; ++c) cnt[(c & 15) + ((c >> 4) & 15) + (c >> 8)]++;
    int run = 0;
    for (int l = 0; l <= kSBLevels; ++l) { B->h_table->lvl_ptr[l] = (unsigned short)run; if (l < kSBLevels) run += cnt[l]; }
    B->h_table->lvl_ptr[kSBLevels + 1] = (unsigned short)run;
    std::vector<int> fill(kSBLevels, 0);
    for (int lz = 0; lz < kSB; ++lz) for (int ly = 0; ly < kSB; ++ly) for (int lx = 0; lx < kSB; ++lx) {
        const int l = lx + ly + lz;
        B->h_table->cell[B->h_table->lvl_ptr[l] + fill[l]++] = (unsigned short)(lx | (ly << 4) | (lz << 8));
    }
    // blocks in block-wavefront order
    std::vector<int> ord((size_t)B->nblk);
    for (int i = 0; i < B->nblk; ++i) ord[i] = i;
    auto wave = [&](int b) { return b % B->nbx + (b / B->nbx) % B->nby + b / (B->nbx * B->nby); };
    std::stable_sort(ord.begin(), ord.end(), [&](int p, int q) { return wave(p) < wave(q); });
    const size_t nrec = (size_t)B->nblk * kSBRows;
    for (int u = 0; u < 2; ++u) {
        CM_CUDA(dev_alloc((void **)&B->d_order[u], sizeof(int) * (size_t)B->nblk));
        CM_CUDA(dev_alloc((void **)&B->d_coef[u], sizeof(double) * 4 * nrec));
        CM_CUDA(dev_alloc((void **)&B->d_pres[u], nrec));
        if (u) std::reverse(ord.begin(), ord.end());
        std::vector<unsigned short> inv(kSBRows);
        for (int pos = 0; pos < kSBRows; ++pos) {
            const int c = B->h_table->cell[pos];
            inv[u ? (kSBRows - 1 - c) : c] = (unsigned short)pos;  // U: cell (15 - lx, 15 - ly, 15 - lz) = 4095 - c
        }
        CM_CUDA(dev_alloc((void **)&B->d_inv[u], sizeof(unsigned short) * kSBRows));
        CM_CUDA(cudaMemcpyAsync(B->d_inv[u], inv.data(), sizeof(unsigned short) * kSBRows, cudaMemcpyHostToDevice, s->stream));
        CM_CUDA(cudaMemcpyAsync(B->d_order[u], ord.data(), sizeof(int) * (size_t)B->nblk, cudaMemcpyHostToDevice, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
    }
    CM_CUDA(dev_alloc((void **)&B->d_flag, sizeof(int) * (size_t)B->nblk));
    CM_CUDA(dev_alloc((void **)&B->d_ticket, sizeof(unsigned)));
    CM_CUDA(cudaMemsetAsync(B->d_flag, 0, sizeof(int) * (size_t)B->nblk, s->stream));
    CM_CUDA(cudaMemsetAsync(B->d_ticket, 0, sizeof(unsigned), s->stream));
    const long long total = (long long)nrec;
    k_sblk_records<<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(B->nblk, B->nbx, B->nby, (int)nx, (int)ny, (int)nz, *B->h_table, s->pre_ia, s->cls[1].d_tmask,
                                                                            s->d_M, B->d_coef[0], B->d_pres[0], B->d_coef[1], B->d_pres[1]);
    CM_CUDA(cudaGetLastError());
    CM_CUDA(cudaStreamSynchronize(s->stream));
    s->launches++;
    int occ = 0, sms = 148;
    CM_CUDA(cudaFuncSetAttribute(k_sptrsv_blocked<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSBSmem));
    CM_CUDA(cudaFuncSetAttribute(k_sptrsv_blocked<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSBSmem));
    CM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sptrsv_blocked<false>, kSBThreads, kSBSmem));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    B->grid = std::max(1, std::min(B->nblk, std::max(1, occ) * sms));
    return CUDAMAT_OK;
}

bool sweepblk_active(const cudamat_solver *s) { return s->bsweep != nullptr; }
int sweepblk_blocks(const cudamat_solver *s) { return s->bsweep ? s->bsweep->nblk : 0; }

int launch_sptrsv_blocked(cudamat_solver *s, bool upper, const double *rhs, double *out) {
    BlockSweep *B = s->bsweep;
    const int *status = s->d_sc ? &s->d_sc->status : nullptr;
    const int epoch = ++B->epoch;
    CM_CUDA(cudaMemsetAsync(B->d_ticket, 0, sizeof(unsigned), s->stream));       // tickets of this sweep start at 0
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)B->grid); cfg.blockDim = dim3(kSBThreads); cfg.stream = s->stream; cfg.dynamicSmemBytes = kSBSmem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 0;                              // the persistent grid must not share SMs with a draining predecessor
    const int u = upper ? 1 : 0;
    static const bool want_dbg = getenv("CUDAMAT_SWEEP_DEBUG") != nullptr;
    long long *dbg = nullptr;
    if (want_dbg) CM_CUDA(cudaMalloc(&dbg, sizeof(long long) * 6 * (size_t)B->nblk));
    if (upper)
        CM_CUDA(cudaLaunchKernelEx(&cfg, k_sptrsv_blocked<true>, B->nblk, B->nbx, B->nby, B->nbz, B->nx, B->ny, B->nz, (const int *)B->d_order[u],
                                   (const double *)B->d_coef[u], (const unsigned char *)B->d_pres[u], rhs, out, B->d_flag, B->d_ticket, epoch,
                                   status, *B->h_table, (const unsigned short *)B->d_inv[u], dbg));
    else
        CM_CUDA(cudaLaunchKernelEx(&cfg, k_sptrsv_blocked<false>, B->nblk, B->nbx, B->nby, B->nbz, B->nx, B->ny, B->nz, (const int *)B->d_order[u],
                                   (const double *)B->d_coef[u], (const unsigned char *)B->d_pres[u], rhs, out, B->d_flag, B->d_ticket, epoch,
                                   status, *B->h_table, (const unsigned short *)B->d_inv[u], dbg));
    s->launches++;
    CM_CUDA(cudaGetLastError());
    if (dbg) {                                                     // CUDAMAT_SWEEP_DEBUG: average SM cycles per phase of a block
        std::vector<long long> h((size_t)6 * B->nblk);
        CM_CUDA(cudaStreamSynchronize(s->stream));
        CM_CUDA(cudaMemcpy(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        double ph[5] = {0, 0, 0, 0, 0};
        for (int b = 0; b < B->nblk; ++b) for (int q = 0; q < 5; ++q) ph[q] += (double)(h[(size_t)b * 6 + q + 1] - h[(size_t)b * 6 + q]);
        fprintf(stderr, "sweep %c: per block cycles: rhs+records %.0f, flag wait %.0f, halo %.0f, levels %.0f, publish %.0f\n", upper ? 'U' : 'L',
                ph[0] / B->nblk, ph[1] / B->nblk, ph[2] / B->nblk, ph[3] / B->nblk, ph[4] / B->nblk);
        cudaFree(dbg);
    }
    return CUDAMAT_OK;
}

}  // namespace cudamat
