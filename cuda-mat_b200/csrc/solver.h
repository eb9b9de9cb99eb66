// solver.h — host-side state of one cudamat_solver handle and the launch wrappers that
// solver.cu (orchestration + C ABI) calls into kernels.cu / ilu0.cu / comm.cu.
#pragma once
#include "internal.cuh"
#include <string>
#include <vector>

namespace cudamat {

void set_error(const char *fmt, ...);
cudaError_t dev_alloc(void **p, size_t bytes);     // solver.cu: pooled allocation for large buffers
void dev_free(void *p);
// hostcopy.cpp: copies between HOST arrays and the device, enqueued on / ordered against `stream`; pageable arrays of
// 8 MB or more go through a threaded pinned staging area.  copy_h2d: src may be reused at return, dst is ready in stream
// order.  copy_d2h: dst is complete only after the caller synchronises `stream` (the staged form completes it at return).
int copy_h2d(void *dst, const void *src, size_t bytes, cudaStream_t stream);
int copy_d2h(void *dst, const void *src, size_t bytes, cudaStream_t stream);
bool cuda_ok(cudaError_t e, const char *what, const char *file, int line);
#define CM_CUDA(call)                                                       \
    do {                                                                    \
        if (!::cudamat::cuda_ok((call), #call, __FILE__, __LINE__)) return CUDAMAT_E_CUDA; \
    } while (0)

struct SpmvArgs {
    int n;                 // local rows
    const int *ia;         // local row pointers (base 0)
    const int *ja;         // local/halo column ids (base 0)
    const double *val;
    const double *x;       // n + nhalo entries
    const double *d;       // optional diagonal shift (local rows) or nullptr
    double *y;
    const double *u;       // dot operand: red0 = y.u   (ndot >= 1)
    int ndot;              // 0, 1 (y.u) or 2 (y.u and y.y)
    int phase;
    RedCtx rc;
    DevScalars *sc;        // may be nullptr when ndot == 0 and status is not checked
    double *hist;
    int check_status;      // 1: return immediately unless sc->status == ST_RUNNING
    HaloWait hw;           // multi-GPU peer-memory path: wait for the neighbours' halo rows of x (nsrc == 0: none)
};

// staged (TMA bulk copy) SpMV plan
struct StagedPlan {
    int cap_nnz = 0;       // per-slab smem capacity in nnz (0 = variant unusable)
    int stages = 0;
    size_t smem_bytes = 0;
};

// row-class dictionaries of the compressed SpMV variants (rowclass.cu)
constexpr int kDictMax = 64;       // classes
constexpr int kDictLen = 16;       // entries per class
struct RowDict { int len[kDictMax]; int off[kDictMax * kDictLen]; double val[kDictMax * kDictLen]; };
// The dictionary travels to the SpMV kernel as a KERNEL PARAMETER (12.6 KB of the 32 KB parameter space): it is
// read through the constant cache (LDC), not through L1/shared memory, which is the busiest unit of that kernel.
struct DictParam {
    int len[kDictMax];
    int run[kDictMax];                 // position of offset -1 when the class holds (-1, 0, +1) consecutively, else -1
    int off[kDictMax * kDictLen];
    double val[kDictMax * kDictLen];
};
// TILED variant: x windows of a tile staged in shared memory by TMA bulk copies (rowclass.cu tiled_plan)
constexpr int kMaxSeg = 4;
struct TiledDict {                     // kernel parameter (16.9 KB)
    int len[kDictMax];
    int off[kDictMax * kDictLen];      // column offsets (fallback path: tiles that are not eligible)
    int disp[kDictMax * kDictLen];     // shared-memory index of the entry relative to the row's position in the tile
    double val[kDictMax * kDictLen];
    int nseg;
    int seg_lo[kMaxSeg];               // first column offset of the window (even)
    int seg_len[kMaxSeg];              // window length in elements = kTile + span (even)
    int seg_base[kMaxSeg];             // start of the window in shared memory (elements)
    int maxlen, minlen;                // longest / shortest class
    int disp0;                         // shared-memory index of x[row] itself relative to the row's position, -1 = not staged
    int sdict_base;                    // start of the shared-memory copy of the dictionary (elements, even)
    // Superset pattern (stencils): when every staged class is an order-preserving subset of ONE pattern of <= 8 entries
    // with identical values at identical offsets, a row is that pattern plus a presence mask (1 byte per row).  Byte
    // offsets and values are then kernel-wide constants with compile-time positions: constant-bank operands of the
    // address add and the FMA, no dictionary read at all in the row loop.  sup_len = 0: not available.
    int sup_len;
    int sup_boff[8];
    double sup_val[8];
    int sup_off[8];                    // column offsets of the superset pattern (ascending)
};
// MARCH variant (march.cu): the superset pattern's offsets split into planes -D / 0 / +D with D a multiple of the tile;
// a persistent CTA walks a column of tiles plane by plane and keeps three planes of the operand in a shared-memory ring.
struct MarchPlan {                     // kernel parameter
    int D;                             // plane stride in rows
    int H;                             // halo elements on each side of a tile buffer (256 or 512)
    int S;                             // tiles per plane = D / kTile
    int P;                             // planes = n / D
    int buf_elems;                     // kTile + 2 H
    int len;                           // entries of the superset pattern
    int shape;                         // 1: (-D, -a, -1, 0, +1, +a, +D) with a even (compile-time planes / alignment), 0: generic
    int Zc;                            // z-chunks: the grid is Zc x S work items (z-chunk-major), set at launch from the CTA budget
    int lo_base, hi_base;              // sharded handles: element offset of the lower / upper neighbour's plane in the operand's halo region (-1: none)
    int n_tot;                         // elements of an operand vector (local rows + halo)
    int dz[8];                         // -1 / 0 / +1: plane the entry reads
    int loff[8];                       // offset inside that plane's buffer relative to the row's own position
    double val[8];
};
// Shared-memory form of the TILED dictionary, one record per class, copied by the same TMA transaction as the windows:
// byte offset of every entry relative to the row's own position in the staged windows (-1 = no entry) and its value.
// Records are read per LANE (16-byte shared loads, broadcast when the lanes of a warp share a class): no indexed
// constant loads in the row loop and no special case for slabs that mix classes.  208 B = 52 words: neighbouring class
// ids fall into different banks.
struct TiledSmemClass { int boff[kDictLen]; double val[kDictLen]; int pad[4]; };
static_assert(sizeof(TiledSmemClass) == 208, "TiledSmemClass layout");
struct TiledArgs { const unsigned char *cls; const unsigned char *tile_ok; const unsigned char *tmask; const TiledSmemClass *sdict; int ncls; int nx; };
struct RowClasses {
    unsigned char *d_cls = nullptr;    // class id per row
    RowDict *d_dict = nullptr;
    DictParam *h_dict = nullptr;       // host copy handed to the kernel launches
    TiledDict *h_tdict = nullptr;      // TILED plan (only for the offsets+values dictionary), nullptr = unavailable
    unsigned char *d_tile_ok = nullptr;
    TiledSmemClass *d_sdict = nullptr; // shared-memory form of the dictionary (TILED)
    unsigned char *d_tmask = nullptr;  // presence mask per row w.r.t. the superset pattern (TILED, when sup_len > 0)
    size_t tiled_smem = 0;
    int ncls = 0;                      // 0 = not available
};
struct ClassArgs { const unsigned char *cls; int ncls; int tiles_per_cta; };

struct Comm;   // comm.cu
struct StreamBlocks;   // stream.cu
struct BlockSweep;     // sweepblk.cu
constexpr int kWorkVecsShared = 12;    // work vectors of a sharded handle's IPC-shared arena (largest solve mode + spare)

struct LevelSchedule {
    int nlevels = 0;
    int *d_order = nullptr;        // rows sorted by level, each level padded to a multiple of 32 with -1
    int *d_level_ptr = nullptr;    // device copy of level_ptr (+ one extra entry) for the single-CTA sweep
    // level-ordered sweep plan (ilu0.cu k_build_plan)
    int *d_cnt = nullptr, *d_ptr = nullptr, *d_col = nullptr; double *d_val = nullptr, *d_dg = nullptr;
    int order_len = 0;
    int *d_chunk_beg = nullptr, *d_chunk_info = nullptr; int nchunks = 0;    // ring sweep (small systems): <= 128 positions per chunk
    std::vector<int> level_ptr;    // host: offsets into d_order per level (padded)
};

}  // namespace cudamat

struct cudamat_solver {
    int64_t n_global = 0, row0 = 0, row1 = 0;
    int n = 0;                       // local rows
    int64_t nnz = 0;
    cudaStream_t stream = nullptr;
    int device = 0;
    // CSR (base 0). owned_* are freed at destroy; the active pointers may alias borrowed memory.
    const int *d_ia = nullptr; const int *d_ja = nullptr; const double *d_a = nullptr;
    int *own_ia = nullptr; int *own_ja = nullptr; double *own_a = nullptr;
    const int *d_ja_global = nullptr;      // global column ids (before halo remap)
    int nhalo = 0;
    // plan
    bool analyzed = false; int analyzed_mode = -1;
    bool csr_checked = false;              // borrowed device CSR went through k_validate_csr (ILU0 needs sorted rows)
    int max_row_len = 0; double mean_row_len = 0; int n_long_rows = 0; int max_slab_nnz = 0;
    int spmv_variant = CUDAMAT_SPMV_ROWLANE;
    int opt_spmv_variant = CUDAMAT_SPMV_AUTO;
    int opt_poll_every = 8;
    int opt_sptrsv_syncfree = 1;
    int opt_debug = 0;
    int opt_time_spmv = 0;
    int loop_it = 0;
    int opt_staged_stages = 0;
    int opt_class_tiles_per_cta = 1;
    int opt_sptrsv_ctas_per_sm = 0;
    int opt_host_analysis = 0;             // 1: ILU0 level analysis on the host (cross-check of the device analysis)
    int opt_graph = -1;                    // CUDA-graph replay of iteration batches: -1 auto, 0 off, 1 force
    bool graph_used = false;
    int opt_sptrsv_no_smem = 0;            // 1: never use the single-CTA shared-memory sweep
    int opt_march_shards = 1;              // sharded handles: MARCH with the neighbours' planes read from the halo region (0: TILED, 2: even when the grid underfills)
    int opt_shard_fuse = 0;                // sharded handles with MARCH: fold the s update into SpMV 2 (boundary planes pushed by a small kernel)
    int opt_sptrsv_ring = 1;               // small systems: role-split ring sweep (0: the barrier-per-level kernel of round 1)
    bool sptrsv_smem_ready = false;
    int sptrsv_grid = 0;
    std::vector<cudaEvent_t> ev_pool; int ev_used = 0;
    std::vector<int> ev_slot; int time_slot = -1; bool ev_open = false;   // event-timed kernels (solver.cu ev_mark)
    int last_fused = 0;
    cudamat::StagedPlan staged;
    cudamat::RowClasses cls[2];            // [0] offsets only (PATTERN), [1] offsets + values (CLASS)
    cudamat::BlockSweep *bsweep = nullptr;  // block-wavefront ILU0 sweeps (7-point grids), nullptr = generic sweeps
    int opt_sptrsv_blocked = 1;
    cudamat::StreamBlocks *sblk = nullptr;  // STREAM variant: column-blocked copy (x larger than the L2), nullptr = one pass
    int opt_persist = -1;                  // persistent cooperative iteration kernel: -1 auto (small systems / shards), 0 off, 1 force
    int persist_grid = 0;
    int opt_stream_blocks = 0;             // 0: automatic (x bytes / 64 MB), 1: never block, K: K column blocks
    cudamat::MarchPlan *march = nullptr;   // MARCH plan (host copy handed to the launches), nullptr = unavailable
    cudamat::RowClasses cls_g;             // sharded handles: classes over GLOBAL column offsets (MARCH's presence masks)
    const unsigned char *march_tmask = nullptr;   // presence mask per row MARCH reads (cls[1] or cls_g)
    int march_grid = 296;                  // persistent CTAs of the MARCH kernels (2 per SM)
    int opt_fuse = 2;                      // bit 0: fold the p update into MARCH SpMV 1, bit 1: the s update into SpMV 2
    int pp = 0;                            // ping-pong parity of the p / v buffers of the fused loop
    int opt_resume = 0;                    // 1: the next solve continues the previous one (no re-initialisation)
    int last_mode = -1;                    // mode of the last finished solve (resume)
    // reduction context + scalars
    cudamat::RedCtx rc{};
    double *slots_own = nullptr;           // the allocation behind rc.slots (rc.slots may be redirected by comm)
    cudamat::DevScalars *d_sc = nullptr;
    cudamat::DevScalars *h_sc = nullptr;   // pinned mirrors: [0] synchronous, [1..2] pipelined polls
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    double *d_hist = nullptr; int hist_cap = 0;
    // work vectors (n + nhalo each)
    double *work = nullptr; size_t work_elems = 0; int work_nvec = 0; bool work_pooled = false; size_t work_bytes = 0;
    // ILU0
    double *d_M = nullptr; int *d_diag = nullptr;
    // matrix behind the preconditioner: the CSR itself, or (sharded handles) the local diagonal block blk_*
    const int *pre_ia = nullptr; const int *pre_ja = nullptr; const double *pre_a = nullptr; int64_t pre_nnz = 0;
    int *blk_ia = nullptr; int *blk_ja = nullptr; double *blk_a = nullptr; int64_t blk_nnz = 0;
    // opt-in multicolour reordering of the preconditioner matrix: perm[new] = old, permuted CSR prm_*
    int opt_ilu0_reorder = 0;
    int *d_perm = nullptr; int *prm_ia = nullptr; int *prm_ja = nullptr; double *prm_a = nullptr;
    cudamat::LevelSchedule lvl_l, lvl_u;
    int *d_flag = nullptr;                 // sync-free epoch flags (n)
    unsigned *d_ticket = nullptr;          // sync-free CTA ticket
    int epoch = 0;
    unsigned ticket_base_l = 0, ticket_base_u = 0;
    int zero_pivot = 0;
    // comm
    cudamat::Comm *comm = nullptr;
    // stats
    int64_t launches = 0;
    std::vector<double> last_hist;
};

namespace cudamat {

// Every handle entry point runs on the handle's own device and restores the caller's current device on return
// (the reference is single-device, callers choose the device beforehand, example.cpp:237).
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != dev && cudaSetDevice(dev) == cudaSuccess) switched = true;
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

int ev_mark(cudamat_solver *s, bool begin);      // solver.cu: event brackets of sampled kernels

// kernels.cu
int launch_spmv(cudamat_solver *s, const SpmvArgs &a, int variant);
int launch_init_resid(cudamat_solver *s, const double *b, const double *y, double *r, double *c1, double *c2, int phase);
int launch_update_p(cudamat_solver *s, bool fma_form, const double *r, const double *v, double *p, const HaloPush *hp = nullptr);
int launch_update_s(cudamat_solver *s, const double *r, const double *v, double *sv, const HaloPush *hp = nullptr);
int launch_update_s_boundary(cudamat_solver *s, const double *r, const double *v, double *sv, const HaloPush *hp, int plane_tiles);
int launch_update_rx_ilu(cudamat_solver *s, const double *v, const double *pw, double *r, double *x);
int launch_update_xr(cudamat_solver *s, bool fma_form, const double *p, const double *sv, const double *t,
                     const double *rhat, double *x, double *r);
int launch_dot(cudamat_solver *s, const double *a, const double *b);
int launch_fill(cudamat_solver *s, double *p, double v, int64_t cnt);
int launch_row_stats(cudamat_solver *s, int *h_out /*[max_len, n_long, max_slab_nnz]*/, double *mean);
int launch_normalize_base(cudaStream_t st, int *ia, int64_t n1, int *ja, int64_t nnz, int base);
int launch_validate_csr(cudaStream_t st, const int *ia, int n, const int *ja, int64_t nnz, int64_t ncols, int *h_bad /*[row, entry out of range, entry out of order] or -1*/);
int plan_staged(cudamat_solver *s);
bool pdl_enabled();

// march.cu
bool march_plan_host(const TiledDict &T, long long n, MarchPlan &M);
bool march_available(const cudamat_solver *s);
bool march_spmv_usable(const cudamat_solver *s, const SpmvArgs &a);     // incl. the 16-byte alignment of the operands
int launch_march_spmv(cudamat_solver *s, const SpmvArgs &a);
int launch_march_make_p(cudamat_solver *s, const double *r, const double *p_old, const double *v_old, double *p_new, double *v_new,
                        const double *rhat, const double *d, const RedCtx &rc);
int launch_march_make_s(cudamat_solver *s, const double *r, const double *v, double *sv, double *t, const double *d, const RedCtx &rc,
                        const HaloWait *hw = nullptr);

// sweepblk.cu: block-wavefront triangular sweeps for ILU0 factors of 7-point grid stencils
int sweepblk_plan(cudamat_solver *s);
void sweepblk_release(cudamat_solver *s);
bool sweepblk_active(const cudamat_solver *s);
int launch_sptrsv_blocked(cudamat_solver *s, bool upper, const double *rhs, double *out);

// persist.cu: the unpreconditioned iteration as one persistent cooperative kernel per batch of iterations
struct PersistLaunch {
    double *r0, *r, *v, *p, *sv, *t, *x; const double *d;
    RedCtx rc; HaloPush hp_p, hp_s; HaloWait hw_p, hw_s; const unsigned long long *red_flags; int iters;
};
bool persist_eligible(const cudamat_solver *s);
int launch_persist(cudamat_solver *s, const PersistLaunch &L);

// stream.cu
int launch_stream_spmv(cudamat_solver *s, const SpmvArgs &a);
int stream_plan(cudamat_solver *s);         // column-blocked copy of an irregular matrix whose x does not fit the L2
void stream_release(cudamat_solver *s);

// rowclass.cu
int rowclass_analyze(cudamat_solver *s);
void rowclass_release(cudamat_solver *s);

// ilu0.cu
int ilu0_analyze_and_factor(cudamat_solver *s, cudamat_stats *st);
int launch_sptrsv(cudamat_solver *s, bool upper, const double *rhs, double *out);
int sptrsv_arm(cudamat_solver *s, double *vec);
int launch_permute(cudamat_solver *s, bool scatter, const double *in, double *out);
void ilu0_release(cudamat_solver *s);

// comm.cu
int comm_halo_exchange(cudamat_solver *s, double *vec);
int march_choose_zc(int S, int P, int G);
int sweepblk_blocks(const cudamat_solver *s);
bool comm_halo_planes(const cudamat_solver *s, int D, int *lo_base, int *hi_base);   // halo = whole planes of the two slab neighbours?
int finish_reduction(cudamat_solver *s, const RedCtx &rc, int phase, int nq);   // groups + cross-rank exchange + final + scalar recurrence
int launch_reduce_finish(cudamat_solver *s, const RedCtx &rc, int nq, int phase, int stage, const double *glob,
                         const unsigned long long *flags);               // kernels.cu
bool comm_p2p(const cudamat_solver *s);                           // peer-memory path active
void comm_begin_reduction(cudamat_solver *s, RedCtx &rc);          // stamps the next reduction epoch into rc (p2p)
bool comm_halo_push(cudamat_solver *s, double *vec, int slot, HaloPush *hp);   // fills hp for the kernel that writes vec
void comm_halo_wait(cudamat_solver *s, int slot, HaloWait *hw);    // fills hw for the SpMV that reads the pushed vector
int ensure_work(cudamat_solver *s, int nvec);
void comm_release(cudamat_solver *s);
// epochs of the peer-memory path {halo slot 0, halo slot 1, reduction}: the persistent kernel consumes one halo epoch per
// slot and three reduction epochs per iteration on the device; the host keeps its counters in step
void comm_epochs_get(const cudamat_solver *s, unsigned long long e[3]);
void comm_epochs_set(cudamat_solver *s, const unsigned long long e[3]);
const unsigned long long *comm_red_flags(const cudamat_solver *s);

// generators (kernels.cu)
int gen_poisson3d(int N, int64_t row0, int64_t row1, int *d_ia, int *d_ja, double *d_a, cudaStream_t st);
int gen_xtrue(uint64_t seed, int64_t i0, int64_t cnt, double *d_out, cudaStream_t st);
int gen_random_dd(int n, uint64_t seed, int *d_ia, int *d_ja, double *d_a, int64_t *nnz_out, cudaStream_t st);

}  // namespace cudamat
