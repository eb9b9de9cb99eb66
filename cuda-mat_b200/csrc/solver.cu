// solver.cu — host orchestration of the BiCGSTAB loops and the C ABI (include/cudamat_b200.h).
//
// The iteration is stream-ordered: scalars (rho, alpha, omega, beta, norms, status, iteration
// counter) live in a device struct updated by the last CTA of each reducing kernel; the host only
// enqueues kernels and polls the status word every `poll_every` iterations through a pinned
// mirror.  Kernels of iterations enqueued past the stopping point return immediately.
#include "solver.h"
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <mutex>
#include <vector>
#include <time.h>

namespace cudamat {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
bool cuda_ok(cudaError_t e, const char *what, const char *file, int line) {
    if (e == cudaSuccess) return true;
    set_error("CUDA error at %s:%d code=%d(%s) \"%s\"", file, line, (int)e, cudaGetErrorName(e), what);
    return false;
}
static double now_s() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// Large buffers of the one-shot host entry points come from a stream-ordered memory pool that keeps freed memory
// mapped: cudaFree of the 2.5 GB a 256^3 solve holds costs ~0.4 s of page unmapping, more than the solve itself.
// (IPC-shared buffers of the multi-GPU path cannot live in the default pool and stay with cudaMalloc.)
// One pool (+ its allocation stream) PER DEVICE, created on first use from that device; an allocation remembers nothing:
// cudaFreeAsync returns memory to the pool it came from whatever device is current.
constexpr int kMaxDev = 64;
struct DevPool { std::once_flag once; bool ok = false; cudaMemPool_t pool = nullptr; cudaStream_t stream = nullptr; };
static DevPool g_pools[kMaxDev];
static DevPool *pool_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { cudaGetLastError(); return nullptr; }
    DevPool &P = g_pools[dev];
    std::call_once(P.once, [&P, dev] {
        const char *off = getenv("CUDAMAT_NO_POOL");
        int supported = 0;
        if (!(off && *off && *off != '0') &&
            cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && supported) {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            if (cudaMemPoolCreate(&P.pool, &props) == cudaSuccess &&
                cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking) == cudaSuccess) {
                // freed memory stays mapped up to this much: half of the device by default.  An 8 GB limit (round 1) sat inside
                // the footprint of one 256^3 ILU0 solve (~8.5 GB of matrix, factor, sweep plans, vectors): the pool trimmed at
                // every synchronisation and re-grew, 1.1 s per analysis
                const char *mb = getenv("CUDAMAT_POOL_KEEP_MB");
                size_t free_b = 0, total_b = 0;
                if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 16ull << 30; }
                unsigned long long keep = mb ? (strtoull(mb, nullptr, 10) << 20) : (unsigned long long)(total_b / 2);
                cudaMemPoolSetAttribute(P.pool, cudaMemPoolAttrReleaseThreshold, &keep);
                P.ok = true;
            }
        }
        cudaGetLastError();
    });
    return P.ok ? &P : nullptr;
}
cudaError_t dev_alloc(void **p, size_t bytes) {
    DevPool *P = pool_for_current_device();
    if (!P) return cudaMalloc(p, bytes);
    cudaError_t e = cudaMallocFromPoolAsync(p, bytes, P->pool, P->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(P->stream);        // usable from any stream on return
}
void dev_free(void *p) {                            // no work may still use p (callers synchronise their stream first)
    if (!p) return;
    // the pointer's own device decides (the caller may have switched devices since the allocation)
    cudaPointerAttributes at{};
    int cur = -1;
    cudaGetDevice(&cur);
    const bool known = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    if (!known) cudaGetLastError();
    const int dev = known ? at.device : cur;
    if (dev >= 0 && dev < kMaxDev && g_pools[dev].ok) { cudaFreeAsync(p, g_pools[dev].stream); return; }
    cudaFree(p);
}

// pinned status mirrors are recycled process-wide: cudaMallocHost / cudaFreeHost cost up to 0.1-0.4 s when other
// large pinned regions exist (page-locking + device-wide synchronisation)
static std::vector<DevScalars *> g_pinned_free;
static std::mutex g_pinned_mu;
static DevScalars *pinned_scalars_get() {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        if (!g_pinned_free.empty()) { DevScalars *p = g_pinned_free.back(); g_pinned_free.pop_back(); return p; }
    }
    DevScalars *p = nullptr;
    if (cudaMallocHost(&p, 3 * sizeof(DevScalars)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
static void pinned_scalars_put(DevScalars *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned_free.push_back(p);
}

static int alloc_reduction(cudamat_solver *s) {
    RedCtx &rc = s->rc;
    rc.ntile = (s->n + kTile - 1) / kTile;
    rc.tile_stride = std::max(rc.ntile, 1);
    rc.ngroup_loc = (rc.ntile + kGroupTiles - 1) / kGroupTiles;
    const int64_t tiles_global = (s->n_global + kTile - 1) / kTile;
    rc.nslots = (int)((tiles_global + kGroupTiles - 1) / kGroupTiles);
    rc.slot_stride = std::max(rc.nslots, 1);
    rc.group0 = (int)(s->row0 / ((int64_t)kTile * kGroupTiles));
    rc.exch_level = 2;
    rc.ntile_global = (int)tiles_global;
    CM_CUDA(dev_alloc((void **)&rc.tile_part, sizeof(double) * kMaxQ * (size_t)rc.tile_stride));
    rc.n_local = s->n;
    rc.slab_stride = std::max((s->n + kSlab - 1) / kSlab, 1);
    CM_CUDA(dev_alloc((void **)&rc.slab_part, sizeof(double) * kMaxQ * (size_t)rc.slab_stride));
    CM_CUDA(dev_alloc((void **)&rc.done_cnt, sizeof(unsigned)));
    CM_CUDA(cudaMemsetAsync(rc.done_cnt, 0, sizeof(unsigned), s->stream));
    CM_CUDA(dev_alloc((void **)&rc.slots, sizeof(double) * kMaxQ * (size_t)rc.slot_stride));
    s->slots_own = rc.slots;
    CM_CUDA(cudaMemsetAsync(rc.slots, 0, sizeof(double) * kMaxQ * (size_t)rc.slot_stride, s->stream));
    CM_CUDA(dev_alloc((void **)&s->d_sc, sizeof(DevScalars)));
    CM_CUDA(cudaMemsetAsync(s->d_sc, 0, sizeof(DevScalars), s->stream));
    s->h_sc = pinned_scalars_get();                                    // [0] synchronous mirror, [1..2] pipelined polls
    if (!s->h_sc) { set_error("cudaMallocHost failed for the status mirror"); return CUDAMAT_E_CUDA; }
    memset(s->h_sc, 0, 3 * sizeof(DevScalars));
    return CUDAMAT_OK;
}

// IPC-shared work arenas of sharded handles are recycled process-wide: mapping / unmapping a 0.6 GB arena in the
// neighbour processes (cudaIpcOpenMemHandle / CloseMemHandle) and cudaFree of it cost ~0.4 s per handle; a recycled arena
// keeps its IPC handle, so the neighbours' cached mappings (comm.cu) stay valid.
static std::mutex g_arena_mu;
constexpr size_t kIpcBlock = 2u << 20;
static double *g_arena_ptr = nullptr; static size_t g_arena_bytes = 0;
static cudaError_t shared_arena_get(double **p, size_t bytes, size_t *got) {
    // whole 2 MB blocks: cudaMalloc packs smaller allocations into shared blocks, and an IPC handle names the BLOCK — a block that a
    // neighbour process still maps (its cached mapping of an arena recycled or freed here) cannot be opened again for another
    // buffer that happens to land in it (seen as: the second sharded handle of a process silently falling back to NCCL)
    bytes = (bytes + kIpcBlock - 1) / kIpcBlock * kIpcBlock;
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        if (g_arena_ptr && g_arena_bytes >= bytes) { *p = g_arena_ptr; *got = g_arena_bytes; g_arena_ptr = nullptr; g_arena_bytes = 0; return cudaSuccess; }
        if (g_arena_ptr) { cudaFree(g_arena_ptr); g_arena_ptr = nullptr; g_arena_bytes = 0; }
    }
    *got = bytes;
    return cudaMalloc(p, bytes);
}
static void shared_arena_put(double *p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_arena_mu);
    if (g_arena_ptr) cudaFree(g_arena_ptr);
    g_arena_ptr = p; g_arena_bytes = bytes;
}

int ensure_work(cudamat_solver *s, int nvec) {
    const size_t elems = (size_t)s->n + (size_t)s->nhalo;
    // keep each vector 256-byte aligned
    const size_t stride = ((elems + 31) / 32) * 32;
    if (s->work && s->work_nvec >= nvec && s->work_elems == stride) return CUDAMAT_OK;
    if (s->work && comm_p2p(s)) {
        // the neighbours hold IPC mappings of this arena and push halo rows into it: reallocating it would leave them
        // writing into freed memory.  p2p_setup sizes it for kWorkVecsShared vectors, which covers every solve mode.
        set_error("work arena of a peer-memory handle cannot grow (%d > %d vectors)", nvec, s->work_nvec);
        return CUDAMAT_E_STATE;
    }
    if (s->work) { cudaStreamSynchronize(s->stream); if (s->work_pooled) dev_free(s->work); else shared_arena_put(s->work, s->work_bytes); s->work = nullptr; }
    s->work_pooled = (s->comm == nullptr);          // sharded handles share the arena over CUDA IPC: plain cudaMalloc, recycled
    const size_t need = sizeof(double) * std::max<size_t>(stride * nvec, 32);
    if (s->work_pooled) { CM_CUDA(dev_alloc((void **)&s->work, need)); s->work_bytes = need; }
    else CM_CUDA(shared_arena_get(&s->work, need, &s->work_bytes));
    s->work_elems = stride; s->work_nvec = nvec;
    return CUDAMAT_OK;
}
static inline double *wv(cudamat_solver *s, int k) { return s->work + (size_t)k * s->work_elems; }

constexpr int64_t kHistCapMax = 1 << 20;      // residual norms kept per solve (8 MB)
static int ensure_hist(cudamat_solver *s, int cap) {
    if (s->d_hist && s->hist_cap >= cap) return CUDAMAT_OK;
    if (s->d_hist) { cudaStreamSynchronize(s->stream); dev_free(s->d_hist); }
    s->d_hist = nullptr;
    CM_CUDA(dev_alloc((void **)&s->d_hist, sizeof(double) * (size_t)std::max(cap, 1)));
    s->hist_cap = cap;
    return CUDAMAT_OK;
}

static SpmvArgs spmv_args(cudamat_solver *s, const double *x, const double *d, double *y, const double *u, int ndot, int phase, int check) {
    SpmvArgs a{};
    a.n = s->n; a.ia = s->d_ia; a.ja = s->d_ja; a.val = s->d_a; a.x = x; a.d = d; a.y = y; a.u = u;
    a.ndot = ndot; a.phase = phase; a.rc = s->rc; a.sc = s->d_sc; a.hist = s->d_hist; a.check_status = check;
    if (ndot > 0) comm_begin_reduction(s, a.rc);
    return a;
}

// Event timing of single kernels ("time_spmv" = k: the main kernels of every k-th iteration are bracketed by CUDA events on
// the launching stream; an event record between two kernels suspends the programmatic-dependent-launch overlap at that
// boundary, so sparse sampling keeps the loop undisturbed).  The loop sets s->time_slot around a launch wrapper; the
// wrappers call ev_mark() right before and after the ONE kernel they launch.
//   slot 0: SpMV 1 (MARCH fused loop: incl. the folded p update)   slot 1: SpMV 2 (incl. the folded s update)
//   slot 2: x / r update with its two dots                          slot 3: separate p / s updates (unfused loop)
int ev_mark(cudamat_solver *s, bool begin) {
    if (s->time_slot < 0) return CUDAMAT_OK;
    if (begin) {
        s->ev_open = false;
        if (s->ev_used + 2 > 16384) return CUDAMAT_OK;
        while ((int)s->ev_pool.size() < s->ev_used + 2) {
            cudaEvent_t e; CM_CUDA(cudaEventCreate(&e)); s->ev_pool.push_back(e);
        }
        CM_CUDA(cudaEventRecord(s->ev_pool[s->ev_used], s->stream));
        s->ev_open = true;
    } else if (s->ev_open) {
        CM_CUDA(cudaEventRecord(s->ev_pool[s->ev_used + 1], s->stream));
        s->ev_slot.push_back(s->time_slot);
        s->ev_used += 2;
        s->ev_open = false;
    }
    return CUDAMAT_OK;
}
static inline bool timed_iteration(const cudamat_solver *s) { return s->opt_time_spmv > 0 && (s->loop_it % s->opt_time_spmv) == 0; }

// SpMV step of the loop: halo exchange of the operand (multi-GPU), the kernel, then the cross-rank
// part of its fused reductions
// (pushed_slot >= 0: the kernel that produced x has already stored its halo rows into the neighbours' copies over
// NVLink — the SpMV's boundary CTAs wait for the neighbours' flags instead of an NCCL exchange)
static int spmv_step(cudamat_solver *s, double *x, const double *d, double *y, const double *u, int ndot, int phase, int check,
                     int pushed_slot, int tslot) {
    SpmvArgs a = spmv_args(s, x, d, y, u, ndot, phase, check);
    if (pushed_slot >= 0 && comm_p2p(s)) comm_halo_wait(s, pushed_slot, &a.hw);
    else { int rc = comm_halo_exchange(s, x); if (rc) return rc; }
    s->time_slot = timed_iteration(s) ? tslot : -1;
    int rc = launch_spmv(s, a, s->spmv_variant);
    s->time_slot = -1;
    if (rc) return rc;
    if (a.ndot > 0) return finish_reduction(s, a.rc, a.phase, a.ndot);
    return CUDAMAT_OK;
}

static int poll_status(cudamat_solver *s) {
    CM_CUDA(cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    return CUDAMAT_OK;
}
// Pipelined status polls: the host enqueues a copy of the scalars every `poll_every` iterations and only waits
// for the copy enqueued one period EARLIER, so the stream always holds at least one period of work and never
// drains while the host looks at the status.  Iterations enqueued past the stopping point return at kernel entry.
static int poll_enqueue(cudamat_solver *s, int k) {
    if (!s->poll_ev[k]) CM_CUDA(cudaEventCreateWithFlags(&s->poll_ev[k], cudaEventDisableTiming));
    CM_CUDA(cudaMemcpyAsync(s->h_sc + 1 + k, s->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaEventRecord(s->poll_ev[k], s->stream));
    return CUDAMAT_OK;
}
static int poll_finished(cudamat_solver *s, int k, bool *stop) {
    CM_CUDA(cudaEventSynchronize(s->poll_ev[k]));
    *stop = s->h_sc[1 + k].status != ST_RUNNING;
    return CUDAMAT_OK;
}
// called after each enqueued iteration; *stop = true when the loop may end
static int poll_step(cudamat_solver *s, int it, int maxit, int *npoll, bool *stop) {
    *stop = false;
    const int poll = std::max(1, s->opt_poll_every);
    if (it % poll != 0 && it != maxit) return CUDAMAT_OK;
    int rc = poll_enqueue(s, *npoll & 1);
    if (rc) return rc;
    if (*npoll > 0 && (rc = poll_finished(s, (*npoll - 1) & 1, stop))) return rc;
    ++*npoll;
    return CUDAMAT_OK;
}

static void fill_stats(cudamat_solver *s, cudamat_stats *st) {
    const DevScalars &h = *s->h_sc;
    st->iterations = h.iter;
    st->converged = h.status == ST_CONVERGED;
    st->breakdown = h.status == ST_CONVERGED ? CUDAMAT_BRK_NONE
                  : h.status == ST_BRK_OMEGA ? CUDAMAT_BRK_OMEGA
                  : h.status == ST_BRK_NAN   ? CUDAMAT_BRK_NAN : CUDAMAT_BRK_MAXIT;
    st->half_steps = std::min(h.half, h.hist_cap);
    st->nrm_r0 = h.nrm0;
    st->nrm_r = h.nrm;
    st->spmv_variant = s->spmv_variant;
    st->levels_l = s->lvl_l.nlevels;
    st->levels_u = s->lvl_u.nlevels;
    st->zero_pivot = s->zero_pivot;
    st->kernel_launches = s->launches;
    st->graph_replay = s->graph_used ? 1 : 0;
}

static int start_scalars(cudamat_solver *s, int maxit, double tol, int hist_cap) {
    DevScalars h{};
    h.rho = 1.0; h.rho_new = 1.0; h.alpha = 1.0; h.omega = 1.0; h.beta = 0.0;   // pbicgstab.cu:614-618
    h.tol = tol; h.maxit = maxit; h.hist_cap = hist_cap; h.status = ST_RUNNING;
    *s->h_sc = h;
    CM_CUDA(cudaMemcpyAsync(s->d_sc, s->h_sc, sizeof(DevScalars), cudaMemcpyHostToDevice, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    return CUDAMAT_OK;
}

// Runs the iteration loop. Small single-GPU systems are launch-bound (8-19 kernels of a few microseconds per
// iteration): there one batch of `poll_every` iterations is captured ONCE into a CUDA graph (the kernels' arguments do
// not change between iterations — every scalar lives on the device) and replayed; the programmatic-dependent-launch
// edges are kept by the capture.  Large systems and sharded handles (per-launch epochs) launch directly.
template <typename Iter>
static int run_iterations(cudamat_solver *s, int maxit, Iter &&one_iteration, bool even_batches = false, bool auto_graph = true) {
    int rc, npoll = 0, it = 0;
    bool stop = false;
    const int poll = std::max(1, s->opt_poll_every);
    static const bool no_graph = [] { const char *e = getenv("CUDAMAT_NO_GRAPH"); return e && *e && *e != '0'; }();
    const bool use_graph = !no_graph && s->opt_graph != 0 && !s->comm && s->opt_time_spmv == 0 && s->stream != nullptr &&
                           s->stream != cudaStreamLegacy && s->stream != cudaStreamPerThread &&
                           (s->opt_graph > 0 || (auto_graph && s->n <= (1 << 22))) && maxit >= 2 * poll &&
                           (!even_batches || poll % 2 == 0);       // ping-pong buffers: a replayed batch must restore the parity
    if (use_graph) {
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        const int64_t l0 = s->launches;
        bool ok = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            rc = CUDAMAT_OK;
            for (int k = 0; k < poll && rc == CUDAMAT_OK; ++k) { s->loop_it = k; rc = one_iteration(); }
            ok = cudaStreamEndCapture(s->stream, &graph) == cudaSuccess && rc == CUDAMAT_OK && graph != nullptr;
        }
        const int64_t per_batch = s->launches - l0;
        s->launches = l0;
        if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        if (!ok) cudaGetLastError();
        while (ok && !stop && it + poll <= maxit) {
            if (!cuda_ok(cudaGraphLaunch(exec, s->stream), "cudaGraphLaunch", __FILE__, __LINE__)) {
                cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
                return CUDAMAT_E_CUDA;
            }
            s->launches += per_batch;
            it += poll;
            if ((rc = poll_step(s, it, maxit, &npoll, &stop))) { cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); return rc; }
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        s->graph_used = ok;
    }
    while (!stop && it < maxit) {
        s->loop_it = it;
        if ((rc = one_iteration())) return rc;
        ++it;
        if ((rc = poll_step(s, it, maxit, &npoll, &stop))) return rc;
    }
    return CUDAMAT_OK;
}

// ---- unpreconditioned loop (gpu_pbicgstab2 shifted overload, pbicgstab.cu:581-754) --------------
// Two schedules, same arithmetic (bit-identical results):
//   * unfused (any SpMV variant, sharded handles): p update | SpMV 1 + dot | s update | SpMV 2 + 2 dots | x, r update + 2 dots
//   * fused (MARCH variant, single GPU): the p and s updates are formed inside the SpMV kernels while the operand is staged
//     (march.cu): 3 bandwidth kernels per iteration; p' and v' go to a second buffer each (ping-pong by iteration parity).
// Option "resume" = 1: continue the previous solve of this handle for `maxit` more iterations without re-initialising
// (bench.py uses it to keep the residual set-up outside a short timed window).
static int solve_unprec(cudamat_solver *s, int mode, const double *d_b, const double *d_x0, const double *d_d,
                        double *d_x, int maxit, double tol) {
    int rc;
    // option "fuse": bit 0 folds the p update into SpMV 1 (MAKE_P), bit 1 the s update into SpMV 2 (MAKE_S)
    const bool persist = persist_eligible(s);                      // one cooperative kernel per batch of iterations (persist.cu)
    // sharded handles: only the s update can be folded (bit 1), and only with the peer-memory halo path (the neighbours' planes of s
    // are pushed by k_update_s_boundary) on shards of at least two planes
    static const int env_sf = [] { const char *e = getenv("CUDAMAT_SHARD_FUSE"); return e && *e ? atoi(e) : -1; }();
    const bool shard_fold = env_sf >= 0 ? env_sf != 0 : s->opt_shard_fuse != 0;
    const int fuse_ok = !s->comm ? 3 : (shard_fold && comm_p2p(s) && s->march && s->march->P >= 2 && (s->march->lo_base >= 0 || s->march->hi_base >= 0)) ? 2 : 0;
    const int fuse = persist ? 4 : (s->spmv_variant == CUDAMAT_SPMV_MARCH && march_available(s) &&
                      (reinterpret_cast<uintptr_t>(d_d) & 15u) == 0) ? (s->opt_fuse & fuse_ok) : 0;   // 16-byte pairs (work vectors are 256-byte aligned)
    const bool fold_p = (fuse & 1) != 0, fold_s = (fuse & 2) != 0;
    const bool resume = s->opt_resume != 0 && s->last_mode == mode && s->last_fused == fuse && s->work != nullptr;
    if ((rc = ensure_work(s, fold_p ? 9 : 7))) return rc;
    // residual history: capped (a caller may pass a huge maxit as "no limit"); hist_push tolerates the overflow
    const int hcap = (int)std::min<int64_t>((int64_t)maxit + 2, kHistCapMax);
    if (!resume && (rc = ensure_hist(s, hcap))) return rc;
    double *r0 = wv(s, 0), *r = wv(s, 1), *v = wv(s, 2), *p = wv(s, 3), *sv = wv(s, 4), *t = wv(s, 5), *xk = wv(s, 6);
    double *pb[2] = {p, fold_p ? wv(s, 8) : p}, *vb[2] = {v, fold_p ? wv(s, 7) : v};
    const int var = s->spmv_variant;
    const size_t nb = sizeof(double) * (size_t)s->n;
    if (resume) {
        DevScalars h = *s->h_sc;                                   // final state of the previous solve (poll_status)
        h.maxit = h.iter + maxit; h.tol = tol; h.status = maxit > 0 ? ST_RUNNING : ST_MAXIT;
        *s->h_sc = h;
        CM_CUDA(cudaMemcpyAsync(s->d_sc, s->h_sc, sizeof(DevScalars), cudaMemcpyHostToDevice, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
    } else {
        s->last_mode = -1;
        if ((rc = start_scalars(s, maxit, tol, hcap))) return rc;
        if (d_x0) CM_CUDA(cudaMemcpyAsync(xk, d_x0, nb, cudaMemcpyDeviceToDevice, s->stream));
        else if ((rc = launch_fill(s, xk, 1.0, s->n))) return rc;
        s->pp = 0;
        CM_CUDA(cudaMemsetAsync(pb[0], 0, sizeof(double) * s->work_elems, s->stream));     // v = p = 0 (:611)
        CM_CUDA(cudaMemsetAsync(vb[0], 0, sizeof(double) * s->work_elems, s->stream));
        // r = b - (A0 + diag d) x0 ; r0 = r ; ||r0|| (:645-655)
        if ((rc = comm_halo_exchange(s, xk))) return rc;
        if ((rc = launch_spmv(s, spmv_args(s, xk, d_d, t, nullptr, 0, PH_NONE, 0), var))) return rc;
        if ((rc = launch_init_resid(s, d_b, t, r, r0, nullptr, PH_U_INIT))) return rc;
        if (maxit <= 0) CM_CUDA(cudaMemsetAsync(xk, 0, nb, s->stream));                 // x stays zero-filled (:1003)
    }
    if (persist) {
        unsigned long long e0[3];
        comm_epochs_get(s, e0);
        const int iter0 = resume ? s->h_sc->iter : 0;
        const int batch = std::max(1, s->opt_poll_every);
        int it = 0, npoll = 0;
        bool stop = false;
        s->graph_used = false;
        while (!stop && it < maxit) {
            const int m = std::min(batch, maxit - it);
            PersistLaunch L{};
            L.r0 = r0; L.r = r; L.v = v; L.p = p; L.sv = sv; L.t = t; L.x = xk; L.d = d_d; L.iters = m;
            L.rc = s->rc;
            if (comm_p2p(s)) {
                unsigned long long e[3];
                comm_halo_push(s, p, 0, &L.hp_p);  comm_halo_wait(s, 0, &L.hw_p);      // epochs of the batch's first iteration
                comm_halo_push(s, sv, 1, &L.hp_s); comm_halo_wait(s, 1, &L.hw_s);
                comm_begin_reduction(s, L.rc);
                comm_epochs_get(s, e);
                e[0] += (unsigned long long)(m - 1); e[1] += (unsigned long long)(m - 1); e[2] += 3ull * m - 1ull;
                comm_epochs_set(s, e);
                L.red_flags = comm_red_flags(s);
            }
            if ((rc = launch_persist(s, L))) return rc;
            it += m;
            if ((rc = poll_step(s, it, maxit, &npoll, &stop))) return rc;
        }
        if ((rc = poll_status(s))) return rc;
        // a batch that stopped early (converged / break-down) consumed fewer epochs than the host assumed
        const unsigned long long done = (unsigned long long)std::max(0, s->h_sc->iter - iter0);
        const unsigned long long e1[3] = {e0[0] + done, e0[1] + done, e0[2] + 3ull * done};
        comm_epochs_set(s, e1);
        rc = CUDAMAT_OK;
    } else
    rc = run_iterations(s, maxit, [&]() -> int {
        int r2;
        const bool tm = timed_iteration(s);
        HaloPush hp;
        double *p_old = pb[s->pp], *p_new = pb[s->pp ^ (fold_p ? 1 : 0)], *v_old = vb[s->pp], *v_new = vb[s->pp ^ (fold_p ? 1 : 0)];
        if (fold_p) {
            s->time_slot = tm ? 0 : -1;
            r2 = launch_march_make_p(s, r, p_old, v_old, p_new, v_new, r0, d_d, s->rc);            // :668-689
            s->time_slot = -1;
            if (r2 || (r2 = finish_reduction(s, s->rc, PH_U_A, 1))) return r2;
        } else {
            int slot = comm_halo_push(s, p_new, 0, &hp) ? 0 : -1;
            s->time_slot = tm ? 3 : -1;
            r2 = launch_update_p(s, false, r, v_old, p_new, &hp);                                   // :668-672 (in place)
            s->time_slot = -1;
            if (r2) return r2;
            if ((r2 = spmv_step(s, p_new, d_d, v_new, r0, 1, PH_U_A, 1, slot, 0))) return r2;       // :675-689
        }
        if (fold_s && s->comm) {
            // slab shard: only the two boundary planes of s are formed (and pushed to the neighbours) by a vector kernel; the folded
            // SpMV 2 forms s for the shard's own planes on the fly and reads the neighbours' planes from the halo region of s
            HaloWait hw{};
            int slot = comm_halo_push(s, sv, 1, &hp) ? 1 : -1;
            s->time_slot = tm ? 3 : -1;
            r2 = launch_update_s_boundary(s, r, v_new, sv, &hp, s->march->S);
            s->time_slot = -1;
            if (r2) return r2;
            if (slot >= 0) comm_halo_wait(s, slot, &hw);
            s->time_slot = tm ? 1 : -1;
            RedCtx rcs = s->rc;
            comm_begin_reduction(s, rcs);                          // stamp the reduction epoch (peer-memory all-gather of the partials)
            r2 = launch_march_make_s(s, r, v_new, sv, t, d_d, rcs, &hw);
            s->time_slot = -1;
            if (r2 || (r2 = finish_reduction(s, rcs, PH_U_B, 2))) return r2;
        } else if (fold_s) {
            s->time_slot = tm ? 1 : -1;
            r2 = launch_march_make_s(s, r, v_new, sv, t, d_d, s->rc);                               // :698-710
            s->time_slot = -1;
            if (r2 || (r2 = finish_reduction(s, s->rc, PH_U_B, 2))) return r2;
        } else {
            int slot = comm_halo_push(s, sv, 1, &hp) ? 1 : -1;
            s->time_slot = tm ? 3 : -1;
            r2 = launch_update_s(s, r, v_new, sv, &hp);                                             // :698-700
            s->time_slot = -1;
            if (r2) return r2;
            if ((r2 = spmv_step(s, sv, d_d, t, sv, 2, PH_U_B, 1, slot, 1))) return r2;              // :703-710
        }
        s->time_slot = tm ? 2 : -1;
        r2 = launch_update_xr(s, false, p_new, sv, t, r0, xk, r);                                   // :694-696,714-747
        s->time_slot = -1;
        if (fold_p) s->pp ^= 1;
        return r2;
    }, /*even_batches=*/fold_p);
    if (rc) return rc;
    if ((rc = poll_status(s))) return rc;
    if (s->h_sc->status == ST_COMM_TIMEOUT) { set_error("a peer rank's halo rows or partial sums did not arrive (spin limit reached)"); return CUDAMAT_E_COMM; }
    CM_CUDA(cudaMemcpyAsync(d_x, xk, nb, cudaMemcpyDeviceToDevice, s->stream));
    s->last_mode = mode; s->last_fused = fuse;
    return CUDAMAT_OK;
}

// ---- ILU0 right-preconditioned loop (gpu_pbicgstab, pbicgstab.cu:45-154) -------------------------
static int solve_ilu0(cudamat_solver *s, const double *d_b, double *d_x, int maxit, double tol) {
    int rc;
    if ((rc = ensure_work(s, s->d_perm ? 10 : 9))) return rc;
    const int hcap = (int)std::min<int64_t>(2 * (int64_t)maxit + 2, kHistCapMax);
    if ((rc = ensure_hist(s, hcap))) return rc;
    double *r = wv(s, 0), *rw = wv(s, 1), *p = wv(s, 2), *pw = wv(s, 3), *sv = wv(s, 4), *t = wv(s, 5), *v = wv(s, 6), *xk = wv(s, 7);
    double *tl = wv(s, 8);          // private output of the L sweeps (carries the sync-free ready sentinel)
    double *tp = s->d_perm ? wv(s, 9) : nullptr;      // multicolour ordering: permuted right-hand side / permuted solution
    const int sf = s->opt_sptrsv_syncfree;
    // out = M^-1 in (pbicgstab.cu:92-98, 121-127); with the multicolour ordering M^-1 = P^T (LU)^-1 P
    auto precond = [&](const double *in, double *out) -> int {
        int rc;
        if (!tp) {
            // sync-free sweeps: each output vector is armed (filled with the ready sentinel) by a coalesced fill
            // right before the sweep that produces it
            if (sf && ((rc = sptrsv_arm(s, tl)) || (rc = sptrsv_arm(s, out)))) return rc;
            if ((rc = launch_sptrsv(s, false, in, tl))) return rc;
            return launch_sptrsv(s, true, tl, out);
        }
        if (sf && (rc = sptrsv_arm(s, tl))) return rc;
        if ((rc = launch_permute(s, false, in, tp))) return rc;
        if ((rc = launch_sptrsv(s, false, tp, tl))) return rc;
        if (sf && (rc = sptrsv_arm(s, tp))) return rc;
        if ((rc = launch_sptrsv(s, true, tl, tp))) return rc;
        return launch_permute(s, true, tp, out);
    };
    const int var = s->spmv_variant;
    if ((rc = start_scalars(s, maxit, tol, hcap))) return rc;
    const size_t nb = sizeof(double) * (size_t)s->n;
    if ((rc = launch_fill(s, xk, 1.0, s->n))) return rc;                                            // :306-308
    CM_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * s->work_elems, s->stream));
    if ((rc = comm_halo_exchange(s, xk))) return rc;
    if ((rc = launch_spmv(s, spmv_args(s, xk, nullptr, t, nullptr, 0, PH_NONE, 0), var))) return rc; // :67
    if ((rc = launch_init_resid(s, d_b, t, r, rw, p, PH_I_INIT))) return rc;                        // :69-74
    rc = run_iterations(s, maxit, [&]() -> int {
        int rc;
        if ((rc = launch_update_p(s, true, r, v, p))) return rc;                                    // :83-89 (skips i == 0)
        if ((rc = precond(p, pw))) return rc;                                                      // :92-98
        if ((rc = spmv_step(s, pw, nullptr, v, rw, 1, PH_I_A, 1, -1, 0))) return rc;        // :104-107
        if ((rc = launch_update_rx_ilu(s, v, pw, r, xk))) return rc;                                // :109-118
        if ((rc = precond(r, sv))) return rc;                                                      // :121-127
        if ((rc = spmv_step(s, sv, nullptr, t, r, 2, PH_I_B, 1, -1, 1))) return rc;         // :132-137
        return launch_update_xr(s, true, nullptr, sv, t, rw, xk, r);                                // :139-151, :81
    }, /*even_batches=*/false, /*auto_graph=*/false);             // sweeps of 20-600 us hide the launches behind PDL: replay + instantiate measured slower
    if (rc) return rc;
    if ((rc = poll_status(s))) return rc;
    if (s->h_sc->status == ST_COMM_TIMEOUT) { set_error("a peer rank's halo rows or partial sums did not arrive (spin limit reached)"); return CUDAMAT_E_COMM; }
    CM_CUDA(cudaMemcpyAsync(d_x, xk, nb, cudaMemcpyDeviceToDevice, s->stream));
    return CUDAMAT_OK;
}

static void print_debug_trace(cudamat_solver *s, int mode) {
    // line formats of pbicgstab.cu:76,113,144 (ILU0) and :484,550 / :657,727 (unpreconditioned)
    const std::vector<double> &h = s->last_hist;
    if (h.empty()) return;
    if (mode == CUDAMAT_MODE_ILU0) {
        printf("gpu, init residual:norm %20.16f\n", h[0]);
        for (size_t k = 1; k < h.size(); ++k) {
            const size_t i = (k - 1) / 2;
            if ((k - 1) % 2 == 0) printf("i = %zu, residual norm (before precond) = %g\n", i, h[k]);
            else printf("i = %zu, residual norm = %g\n", i, h[k]);
        }
    } else {
        printf("initial norm = %g\n", h[0]);
        for (size_t k = 1; k < h.size(); ++k) printf("k = %zu, norm = %g\n", k - 1, h[k]);
        if (s->h_sc->status == ST_BRK_OMEGA || s->h_sc->status == ST_BRK_NAN)
            printf("omega is close to zero, cannot continue\nomega = %g\n", s->h_sc->omega);
    }
}

}  // namespace cudamat

using namespace cudamat;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int cudamat_abi_version(void) { return CUDAMAT_ABI_VERSION; }
const char *cudamat_last_error(void) { return cudamat::g_err; }

int cudamat_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

static int require_device() {
    if (cudamat_device_count() <= 0) {
        set_error("no CUDA device available: libcudamat_b200 has no CPU fallback");
        return CUDAMAT_E_NO_DEVICE;
    }
    return CUDAMAT_OK;
}

int cudamat_create(cudamat_solver **out, int64_t n_global, int64_t row0, int64_t row1, void *stream) {
    if (!out || n_global < 0 || row0 < 0 || row1 < row0 || row1 > n_global || row1 - row0 > 0x7fffffffLL) {
        set_error("cudamat_create: invalid shard [%lld,%lld) of %lld", (long long)row0, (long long)row1, (long long)n_global);
        return CUDAMAT_E_INVALID;
    }
    int rc = require_device();
    if (rc) return rc;
    if ((row0 % kTile) != 0 && row0 != 0) {
        set_error("cudamat_create: row0 must be a multiple of %d (reduction tile)", kTile);
        return CUDAMAT_E_INVALID;
    }
    cudamat_solver *s = new cudamat_solver();
    s->n_global = n_global; s->row0 = row0; s->row1 = row1; s->n = (int)(row1 - row0);
    s->stream = (cudaStream_t)stream;
    cudaGetDevice(&s->device);
    rc = alloc_reduction(s);
    if (rc) { cudamat_destroy(s); return rc; }
    *out = s;
    return CUDAMAT_OK;
}

int cudamat_destroy(cudamat_solver *s) {
    if (!s) return CUDAMAT_OK;
    DeviceGuard dg(s->device);
    const bool tm = getenv("CUDAMAT_TIMING") != nullptr;
    const double t0 = now_s();
    cudaStreamSynchronize(s->stream);
    comm_release(s);
    ilu0_release(s);
    rowclass_release(s);
    stream_release(s);
    const double t1 = now_s();
    dev_free(s->own_ia); dev_free(s->own_ja); dev_free(s->own_a);
    dev_free(s->rc.tile_part); dev_free(s->rc.slab_part); dev_free(s->rc.done_cnt);
    dev_free(s->slots_own);
    dev_free(s->d_sc);
    pinned_scalars_put(s->h_sc);
    dev_free(s->d_hist);
    if (s->work) { if (s->work_pooled) dev_free(s->work); else shared_arena_put(s->work, s->work_bytes); }
    for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : s->poll_ev) if (e) cudaEventDestroy(e);
    if (tm) fprintf(stderr, "cudamat_destroy: sync+release %.4f s, frees %.4f s\n", t1 - t0, now_s() - t1);
    delete s;
    return CUDAMAT_OK;
}

int cudamat_set_option(cudamat_solver *s, const char *key, int64_t value) {
    if (!s || !key) return CUDAMAT_E_INVALID;
    if (!strcmp(key, "spmv_variant")) { s->opt_spmv_variant = (int)value; s->analyzed = false; }
    else if (!strcmp(key, "poll_every")) s->opt_poll_every = (int)value;
    else if (!strcmp(key, "sptrsv_syncfree")) s->opt_sptrsv_syncfree = (int)value;
    else if (!strcmp(key, "debug")) s->opt_debug = (int)value;
    else if (!strcmp(key, "time_spmv")) s->opt_time_spmv = (int)value;
    else if (!strcmp(key, "sptrsv_ctas_per_sm")) { s->opt_sptrsv_ctas_per_sm = (int)value; s->sptrsv_grid = 0; }
    else if (!strcmp(key, "ilu0_reorder")) { s->opt_ilu0_reorder = (int)value; s->analyzed = false; }
    else if (!strcmp(key, "host_analysis")) { s->opt_host_analysis = (int)value; s->analyzed = false; }
    else if (!strcmp(key, "graph")) s->opt_graph = (int)value;          // -1 auto (small systems), 0 off, 1 force
    else if (!strcmp(key, "sptrsv_no_smem")) s->opt_sptrsv_no_smem = (int)value;
    else if (!strcmp(key, "sptrsv_ring")) s->opt_sptrsv_ring = (int)value;
    else if (!strcmp(key, "march_shards")) s->opt_march_shards = (int)value;
    else if (!strcmp(key, "shard_fuse")) s->opt_shard_fuse = (int)value;
    else if (!strcmp(key, "class_tiles_per_cta")) s->opt_class_tiles_per_cta = (int)value;
    else if (!strcmp(key, "staged_stages")) { s->opt_staged_stages = (int)value; s->analyzed = false; }
    else if (!strcmp(key, "l2_fetch")) {
        // device-wide hint: bytes the L2 fetches from DRAM per miss (32 / 64 / 128); random gathers want 32 (see analyze)
        DeviceGuard dg(s->device);
        CM_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
    }
    else if (!strcmp(key, "persist")) s->opt_persist = (int)value;
    else if (!strcmp(key, "sptrsv_blocked")) { s->opt_sptrsv_blocked = (int)value; s->analyzed = false; }
    else if (!strcmp(key, "stream_blocks")) { s->opt_stream_blocks = (int)value; s->analyzed = false; }
    else if (!strcmp(key, "fuse")) s->opt_fuse = (int)value;
    else if (!strcmp(key, "resume")) s->opt_resume = (int)value;
    else if (!strcmp(key, "march_grid")) s->march_grid = std::max(1, (int)value);
    else { set_error("unknown option '%s'", key); return CUDAMAT_E_INVALID; }
    return CUDAMAT_OK;
}

int cudamat_set_csr_host(cudamat_solver *s, int nnz, const double *A, const int *iA, const int *jA) {
    if (!s || !iA || (nnz > 0 && (!A || !jA)) || nnz < 0) { set_error("set_csr_host: null argument"); return CUDAMAT_E_INVALID; }
    DeviceGuard dg(s->device);
    const int n = s->n;
    const int base = iA[0];
    if (base != 0 && base != 1) { set_error("set_csr_host: index base %d is neither 0 nor 1 (pbicgstab.cu:201)", base); return CUDAMAT_E_INVALID; }
    if (iA[n] - base != nnz) { set_error("set_csr_host: iA[n]-iA[0]=%d != nnz=%d", iA[n] - base, nnz); return CUDAMAT_E_INVALID; }
    if (s->own_ia) { cudaStreamSynchronize(s->stream); dev_free(s->own_ia); dev_free(s->own_ja); dev_free(s->own_a); s->own_ia = nullptr; s->own_ja = nullptr; s->own_a = nullptr; }
    // +16 bytes of slack so aligned bulk copies of the last slab stay inside the allocation
    CM_CUDA(dev_alloc((void **)&s->own_ia, sizeof(int) * (size_t)(n + 1)));
    CM_CUDA(dev_alloc((void **)&s->own_ja, sizeof(int) * (size_t)std::max(nnz, 1) + 16));
    CM_CUDA(dev_alloc((void **)&s->own_a, sizeof(double) * (size_t)std::max(nnz, 1) + 16));
    int rc = copy_h2d(s->own_ia, iA, sizeof(int) * (size_t)(n + 1), s->stream);
    if (!rc && nnz > 0) rc = copy_h2d(s->own_ja, jA, sizeof(int) * (size_t)nnz, s->stream);
    if (!rc && nnz > 0) rc = copy_h2d(s->own_a, A, sizeof(double) * (size_t)nnz, s->stream);
    if (rc) return rc;
    rc = launch_normalize_base(s->stream, s->own_ia, n + 1, s->own_ja, nnz, base);
    if (rc) return rc;
    // structure checks run on the device over the uploaded arrays (a host pass over nnz entries would cost as much
    // as the upload itself): monotone row pointers, column indices inside [0, n_global)
    int bad[3] = {-1, -1, -1};
    if ((rc = launch_validate_csr(s->stream, s->own_ia, n, s->own_ja, nnz, s->n_global, bad))) return rc;
    if (bad[0] >= 0) { set_error("set_csr_host: row pointers not monotone at row %d", bad[0]); return CUDAMAT_E_INVALID; }
    if (bad[1] >= 0) { set_error("set_csr_host: column index out of range at entry %d", bad[1]); return CUDAMAT_E_INVALID; }
    if (bad[2] >= 0) { set_error("set_csr_host: column indices must be strictly ascending within a row (entry %d; mmio_wrapper.h:123-126)", bad[2]); return CUDAMAT_E_INVALID; }
    s->d_ia = s->own_ia; s->d_ja = s->own_ja; s->d_a = s->own_a; s->d_ja_global = s->own_ja;
    s->nnz = nnz; s->analyzed = false;
    return CUDAMAT_OK;
}

int cudamat_set_csr_device(cudamat_solver *s, int64_t nnz, const double *dA, const int *dIA, const int *dJA) {
    if (!s || !dIA || (nnz > 0 && (!dA || !dJA)) || nnz < 0 || nnz > 0x7fffffffLL) { set_error("set_csr_device: invalid argument"); return CUDAMAT_E_INVALID; }
    DeviceGuard dg(s->device);
    int ends[2] = {0, 0};
    CM_CUDA(cudaMemcpyAsync(&ends[0], dIA, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaMemcpyAsync(&ends[1], dIA + s->n, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    if (ends[0] != 0 || ends[1] != nnz) { set_error("set_csr_device: expects base-0 row pointers with ia[n]==nnz (got ia[0]=%d ia[n]=%d nnz=%lld)", ends[0], ends[1], (long long)nnz); return CUDAMAT_E_INVALID; }
    s->d_ia = dIA; s->d_ja = dJA; s->d_a = dA; s->d_ja_global = dJA; s->nnz = nnz; s->analyzed = false; s->csr_checked = false;
    return CUDAMAT_OK;
}

int cudamat_analyze(cudamat_solver *s, int mode, cudamat_stats *st) {
    if (!s || !s->d_ia) { set_error("analyze: no matrix set"); return CUDAMAT_E_STATE; }
    if (mode < 0 || mode > 2) { set_error("analyze: bad mode %d", mode); return CUDAMAT_E_INVALID; }
    DeviceGuard dg(s->device);
    if (!s->comm && (s->row0 != 0 || s->row1 != s->n_global)) {
        // the CSR of a true shard still holds GLOBAL column ids until cudamat_comm_init renumbers them [local | halo]
        set_error("analyze: handle owns the shard [%lld,%lld) of %lld rows but cudamat_comm_init was not called",
                  (long long)s->row0, (long long)s->row1, (long long)s->n_global);
        return CUDAMAT_E_STATE;
    }
    const double t0 = now_s();
    int hs[3] = {0, 0, 0};
    int rc = launch_row_stats(s, hs, &s->mean_row_len);
    if (rc) return rc;
    s->max_row_len = hs[0]; s->n_long_rows = hs[1]; s->max_slab_nnz = hs[2];
    if ((rc = plan_staged(s))) return rc;
    const bool aligned = (((uintptr_t)s->d_a) % 16 == 0) && (((uintptr_t)s->d_ja) % 16 == 0);
    if (!aligned) s->staged = StagedPlan();
    // Variant choice from the row statistics.  Measured on B200 (profiles/): with <= 32 entries per row the direct
    // row-per-lane kernel streams CSR at ~99 % of the copy roofline, ahead of the TMA-staged kernel whose
    // shared-memory ring caps occupancy (opt-in).  When the rows fall into a few classes up to translation
    // (stencil matrices) the dictionary variants move far fewer bytes and win outright: TILED (CLASS + x windows staged
    // in shared memory by TMA) > CLASS (offsets and values from the dictionary) > PATTERN (offsets only) > ROWLANE.
    if ((rc = rowclass_analyze(s))) return rc;
    int variant = s->opt_spmv_variant;
    if (variant == CUDAMAT_SPMV_AUTO)
        variant = march_available(s) ? CUDAMAT_SPMV_MARCH
                : (s->cls[1].h_tdict || s->cls[0].h_tdict) ? CUDAMAT_SPMV_TILED : s->cls[1].ncls > 0 ? CUDAMAT_SPMV_CLASS
                : CUDAMAT_SPMV_ROWLANE;      // PATTERN without the staged windows measured slower than CSR (0.36 vs 0.32 ms): explicit only
    // Row-length statistics (k_row_stats): the row-per-lane kernel walks every lane of a warp through the longest row of its
    // 32; when the fullest slab holds much more than 32 x the mean row length, or rows > 32 entries exist (the warp serialises
    // on them), the entry-parallel STREAM kernel takes over.
    if (variant == CUDAMAT_SPMV_ROWLANE && s->opt_spmv_variant == CUDAMAT_SPMV_AUTO && s->n > 0 &&
        (s->n_long_rows > 0 || s->max_row_len > 2.0 * s->mean_row_len + 4.0))
        variant = CUDAMAT_SPMV_STREAM;
    if (variant == CUDAMAT_SPMV_MARCH && !march_available(s)) variant = CUDAMAT_SPMV_TILED;
    if (variant == CUDAMAT_SPMV_TILED && !s->cls[1].h_tdict && !s->cls[0].h_tdict) variant = CUDAMAT_SPMV_CLASS;
    if (variant == CUDAMAT_SPMV_CLASS && s->cls[1].ncls == 0) variant = CUDAMAT_SPMV_PATTERN;
    if (variant == CUDAMAT_SPMV_PATTERN && s->cls[0].ncls == 0) variant = CUDAMAT_SPMV_ROWLANE;
    if (variant == CUDAMAT_SPMV_STAGED && s->staged.cap_nnz == 0) variant = CUDAMAT_SPMV_ROWLANE;
    s->spmv_variant = variant;
    if (variant == CUDAMAT_SPMV_STREAM) { if ((rc = stream_plan(s))) return rc; } else stream_release(s);
    if (st) st->t_analysis += now_s() - t0;
    if (mode == CUDAMAT_MODE_ILU0) {
        // borrowed device CSR was never validated: the ILU0 path (diagonal search, L/U split, level analysis, sync-free
        // sweeps) needs strictly ascending columns within a row, as the reference loader guarantees (mmio_wrapper.h:123-126)
        if (!s->comm && s->d_ia != s->own_ia && !s->csr_checked) {
            int bad[3] = {-1, -1, -1};
            if ((rc = launch_validate_csr(s->stream, s->d_ia, s->n, s->d_ja, s->nnz, s->n_global, bad))) return rc;
            if (bad[0] >= 0) { set_error("analyze: row pointers not monotone at row %d", bad[0]); return CUDAMAT_E_INVALID; }
            if (bad[1] >= 0) { set_error("analyze: column index out of range at entry %d", bad[1]); return CUDAMAT_E_INVALID; }
            if (bad[2] >= 0) { set_error("analyze: ILU0 needs strictly ascending column indices within a row (entry %d)", bad[2]); return CUDAMAT_E_INVALID; }
            s->csr_checked = true;
        }
        if ((rc = ilu0_analyze_and_factor(s, st))) return rc;
    }
    s->analyzed = true; s->analyzed_mode = mode; s->last_mode = -1;
    if (st) { st->spmv_variant = s->spmv_variant; st->kernel_launches = s->launches; }
    return CUDAMAT_OK;
}

int cudamat_solve_device(cudamat_solver *s, int mode, const double *d_b, const double *d_x0, const double *d_d,
                         double *d_x, int maxit, double tol, cudamat_stats *st) {
    if (!s || !d_b || !d_x) { set_error("solve_device: null argument"); return CUDAMAT_E_INVALID; }
    if (!s->analyzed || (mode == CUDAMAT_MODE_ILU0 && !s->d_M)) { set_error("solve_device: call cudamat_analyze(mode) first"); return CUDAMAT_E_STATE; }
    if (mode == CUDAMAT_MODE_SHIFTED && (!d_d || !d_x0)) { set_error("solve_device: shifted mode needs d and x0 (pbicgstab.h:116)"); return CUDAMAT_E_INVALID; }
    if (maxit < 0) maxit = 0;
    DeviceGuard dg(s->device);
    const double t0 = now_s();
    s->ev_used = 0; s->ev_slot.clear(); s->time_slot = -1;
    int rc;
    if (mode == CUDAMAT_MODE_ILU0) rc = solve_ilu0(s, d_b, d_x, maxit, tol);
    else rc = solve_unprec(s, mode, d_b, d_x0, d_d, d_x, maxit, tol);
    if (rc) return rc;
    CM_CUDA(cudaStreamSynchronize(s->stream));
    const double t1 = now_s();
    // residual history for the debug trace / cudamat_get_history
    const int nh = std::min(s->h_sc->half, s->h_sc->hist_cap);
    s->last_hist.assign((size_t)nh, 0.0);
    if (nh > 0) {
        CM_CUDA(cudaMemcpyAsync(s->last_hist.data(), s->d_hist, sizeof(double) * (size_t)nh, cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
    }
    if (s->opt_debug) print_debug_trace(s, mode);
    if (st) { fill_stats(s, st); st->t_loop = t1 - t0; }
    if (st) {
        for (int q = 0; q < 4; ++q) { st->t_kernel[q] = 0.0; st->n_kernel[q] = 0; }
        st->t_spmv = 0.0; st->n_spmv = 0;
        st->fused = mode != CUDAMAT_MODE_ILU0 ? s->last_fused : 0;
    }
    if (st && s->ev_used > 0) {
        for (int k = 0; k + 1 < s->ev_used; k += 2) {
            float ms = 0.f;
            const int slot = s->ev_slot[(size_t)k / 2];
            if (slot >= 0 && slot < 4 && cudaEventElapsedTime(&ms, s->ev_pool[k], s->ev_pool[k + 1]) == cudaSuccess) {
                st->t_kernel[slot] += ms * 1e-3; st->n_kernel[slot] += 1;
            }
        }
        st->t_spmv = st->t_kernel[0] + st->t_kernel[1]; st->n_spmv = st->n_kernel[0] + st->n_kernel[1];
    }
    return CUDAMAT_OK;
}

int cudamat_get_history(cudamat_solver *s, double *hist, int cap) {
    if (!s || (!hist && cap > 0)) return 0;
    const int m = std::min<int>(cap, (int)s->last_hist.size());
    for (int i = 0; i < m; ++i) hist[i] = s->last_hist[i];
    return m;
}

int cudamat_spmv_device(cudamat_solver *s, const double *d_x, const double *d_d, double *d_y, int variant) {
    if (!s || !d_x || !d_y) { set_error("spmv_device: null argument"); return CUDAMAT_E_INVALID; }
    if (!s->analyzed) { set_error("spmv_device: call cudamat_analyze first"); return CUDAMAT_E_STATE; }
    DeviceGuard dg(s->device);
    if (s->comm) {
        // sharded handle: stage x next to its halo, exchange, multiply (collective over all ranks)
        int rc = ensure_work(s, 8);
        if (rc) return rc;
        double *xs = wv(s, 7);
        CM_CUDA(cudaMemcpyAsync(xs, d_x, sizeof(double) * (size_t)s->n, cudaMemcpyDeviceToDevice, s->stream));
        if ((rc = comm_halo_exchange(s, xs))) return rc;
        return launch_spmv(s, spmv_args(s, xs, d_d, d_y, nullptr, 0, PH_NONE, 0), variant);
    }
    return launch_spmv(s, spmv_args(s, d_x, d_d, d_y, nullptr, 0, PH_NONE, 0), variant);
}

int cudamat_dot_device(cudamat_solver *s, const double *d_a, const double *d_b, double *result) {
    if (!s || !d_a || !d_b || !result) { set_error("dot_device: null argument"); return CUDAMAT_E_INVALID; }
    if (s->n == 0) { *result = 0.0; return CUDAMAT_OK; }
    DeviceGuard dg(s->device);
    int rc = launch_dot(s, d_a, d_b);
    if (rc) return rc;
    if ((rc = poll_status(s))) return rc;
    *result = s->h_sc->red[0];
    return CUDAMAT_OK;
}

int cudamat_get_ilu0_host(cudamat_solver *s, double *M_out) {
    if (!s || !M_out) return CUDAMAT_E_INVALID;
    if (!s->d_M) { set_error("get_ilu0_host: no factor (analyze with CUDAMAT_MODE_ILU0)"); return CUDAMAT_E_STATE; }
    DeviceGuard dg(s->device);
    if (s->d_perm) { set_error("get_ilu0_host: the factor belongs to the multicolour-permuted matrix (option ilu0_reorder)"); return CUDAMAT_E_STATE; }
    if (s->pre_nnz != s->nnz) { set_error("get_ilu0_host: sharded handles hold a block-Jacobi factor of the local block (%lld entries), not A's pattern", (long long)s->pre_nnz); return CUDAMAT_E_STATE; }
    int rc = copy_d2h(M_out, s->d_M, sizeof(double) * (size_t)s->nnz, s->stream);
    if (rc) return rc;
    CM_CUDA(cudaStreamSynchronize(s->stream));
    return CUDAMAT_OK;
}

int cudamat_sweep_blocks(cudamat_solver *s) { return s ? cudamat::sweepblk_blocks(s) : 0; }
int cudamat_sptrsv_device(cudamat_solver *s, int upper, const double *d_rhs, double *d_out) {
    if (!s || !d_rhs || !d_out) return CUDAMAT_E_INVALID;
    DeviceGuard dg(s->device);
    // kernel-level calls must not be gated by a finished solve
    int st = ST_RUNNING;
    CM_CUDA(cudaMemcpyAsync(&s->d_sc->status, &st, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    int rc = CUDAMAT_OK;
    if (s->opt_sptrsv_syncfree && (rc = sptrsv_arm(s, d_out))) return rc;
    rc = launch_sptrsv(s, upper != 0, d_rhs, d_out);
    if (rc) return rc;
    CM_CUDA(cudaStreamSynchronize(s->stream));
    return CUDAMAT_OK;
}

// ---- one-shot host entry points ---------------------------------------------------------------
int cudamat_bicgstab_host(int mode, int n, int nnz, const double *A, const int *iA, const int *jA,
                          const double *d, const double *x0, const double *b,
                          int maxit, double tol, int debug, double *x, double *dtAlg, cudamat_stats *st_out) {
    if (n < 0 || !iA || !b || !x) { set_error("bicgstab_host: null argument"); return CUDAMAT_E_INVALID; }
    if (mode == CUDAMAT_MODE_SHIFTED && (!d || !x0)) { set_error("bicgstab_host: shifted mode needs d and x0"); return CUDAMAT_E_INVALID; }
    cudamat_stats st{};
    cudamat_solver *s = nullptr;
    const bool tm = getenv("CUDAMAT_TIMING") != nullptr;      // stderr breakdown of the host entry point
    const double tt0 = now_s();
    int rc = require_device();
    if (rc) return rc;
    // one private non-blocking stream per thread for the one-shot entry points (CUDA graphs cannot be captured on the
    // legacy default stream); every operation of the call is enqueued on it or is synchronous
    static thread_local cudaStream_t host_streams[kMaxDev] = {};          // per thread AND per device
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= kMaxDev) { cudaGetLastError(); cur_dev = 0; }
    cudaStream_t &host_stream = host_streams[cur_dev];
    if (!host_stream && cudaStreamCreateWithFlags(&host_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); host_stream = nullptr; }
    rc = cudamat_create(&s, n, 0, n, host_stream);
    if (rc) return rc;
    s->opt_debug = debug;
    if (debug) printf("N=%d, nnz=%d\n", n, nnz);                                  // pbicgstab.cu:203
    double *d_b = nullptr, *d_x = nullptr, *d_x0 = nullptr, *d_d = nullptr;
    auto cleanup = [&]() {
        cudaDeviceSynchronize();
        dev_free(d_b); dev_free(d_x); dev_free(d_x0); dev_free(d_d);
        cudamat_destroy(s);
    };
#define HOST_TRY(expr) do { rc = (expr); if (rc) { cleanup(); return rc; } } while (0)
#define HOST_CUDA(call) do { if (!cuda_ok((call), #call, __FILE__, __LINE__)) { cleanup(); return CUDAMAT_E_CUDA; } } while (0)
    double t0 = now_s();
    HOST_TRY(cudamat_set_csr_host(s, nnz, A, iA, jA));
    const size_t nb = sizeof(double) * (size_t)std::max(n, 1);
    HOST_CUDA(dev_alloc((void **)&d_b, nb));
    HOST_CUDA(dev_alloc((void **)&d_x, nb));
    // uploads are ordered on the call's own stream (a legacy-stream cudaMemcpy from pageable memory is not ordered against
    // a cudaStreamNonBlocking stream); the host arrays stay untouched until the stream is synchronised by analyze / solve
    HOST_TRY(copy_h2d(d_b, b, sizeof(double) * (size_t)n, s->stream));
    if (mode == CUDAMAT_MODE_SHIFTED || (mode == CUDAMAT_MODE_PLAIN && x0)) {
        HOST_CUDA(dev_alloc((void **)&d_x0, nb));
        HOST_TRY(copy_h2d(d_x0, x0, sizeof(double) * (size_t)n, s->stream));
    }
    if (mode != CUDAMAT_MODE_ILU0 && d) {
        HOST_CUDA(dev_alloc((void **)&d_d, nb));
        HOST_TRY(copy_h2d(d_d, d, sizeof(double) * (size_t)n, s->stream));
    }
    HOST_CUDA(cudaStreamSynchronize(s->stream));
    st.t_h2d = now_s() - t0;
    const double tt1 = now_s();
    HOST_TRY(cudamat_analyze(s, mode, &st));
    const double tt2 = now_s();
    if (debug && mode == CUDAMAT_MODE_ILU0) {
        printf("analysis lower+upper %f (s), levels %d / %d\n", st.t_analysis, st.levels_l, st.levels_u);   // cf. pbicgstab.cu:349
        printf("ILU0 factorisation time(s) = %10.8f \n", st.t_ilu0);                                         // cf. :354-363
    }
    HOST_TRY(cudamat_solve_device(s, mode, d_b, d_x0, d_d, d_x, maxit, tol, &st));
    t0 = now_s();
    HOST_TRY(copy_d2h(x, d_x, sizeof(double) * (size_t)n, s->stream));
    HOST_CUDA(cudaStreamSynchronize(s->stream));
    st.t_d2h = now_s() - t0;
    st.kernel_launches = s->launches;
    if (dtAlg) *dtAlg = st.t_loop;
    if (st_out) *st_out = st;
    const double tt3 = now_s();
    cleanup();
    if (tm) fprintf(stderr, "cudamat_bicgstab_host: create+upload %.4f s, analyze %.4f s, solve+download %.4f s, cleanup %.4f s\n",
                    tt1 - tt0, tt2 - tt1, tt3 - tt2, now_s() - tt3);
#undef HOST_TRY
#undef HOST_CUDA
    return CUDAMAT_OK;
}

int cudamat_ilu0_host(int n, int nnz, const double *A, const int *iA, const int *jA,
                      double *M_out, int *levels, int *zero_pivot) {
    if (n < 0 || !iA || !M_out) { set_error("ilu0_host: null argument"); return CUDAMAT_E_INVALID; }
    cudamat_solver *s = nullptr;
    int rc = cudamat_create(&s, n, 0, n, nullptr);
    if (rc) return rc;
    cudamat_stats st{};
    rc = cudamat_set_csr_host(s, nnz, A, iA, jA);
    if (!rc) rc = cudamat_analyze(s, CUDAMAT_MODE_ILU0, &st);
    if (!rc) rc = cudamat_get_ilu0_host(s, M_out);
    if (!rc) {
        if (levels) { levels[0] = st.levels_l; levels[1] = st.levels_u; }
        if (zero_pivot) *zero_pivot = st.zero_pivot;
    }
    cudamat_destroy(s);
    return rc;
}

// ---- generators -------------------------------------------------------------------------------
int64_t cudamat_poisson3d_nnz(int N, int64_t row0, int64_t row1) {
    // closed form per row would do; a loop over planes is cheap enough on the host
    const int64_t nn = (int64_t)N * N;
    int64_t cnt = 0;
    for (int64_t r = row0; r < row1;) {
        // whole rows of the grid (N consecutive matrix rows) at once
        const int64_t i = r % N, j = (r / N) % N, k = r / nn;
        if (i == 0 && r + N <= row1) {
            const int nb = (k > 0) + (j > 0) + (j < N - 1) + (k < N - 1);
            cnt += (int64_t)N * (1 + nb) + 2 * (int64_t)(N - 1);
            r += N;
        } else {
            cnt += 1 + (k > 0) + (j > 0) + (i > 0) + (i < N - 1) + (j < N - 1) + (k < N - 1);
            r += 1;
        }
    }
    return cnt;
}
int cudamat_gen_poisson3d_device(int N, int64_t row0, int64_t row1, int *d_ia, int *d_ja, double *d_a, void *stream) {
    int rc = require_device();
    if (rc) return rc;
    if (N <= 0 || row0 < 0 || row1 < row0 || row1 > (int64_t)N * N * N || !d_ia) { set_error("gen_poisson3d: invalid argument"); return CUDAMAT_E_INVALID; }
    return gen_poisson3d(N, row0, row1, d_ia, d_ja, d_a, (cudaStream_t)stream);
}
int cudamat_gen_xtrue_device(uint64_t seed, int64_t i0, int64_t cnt, double *d_out, void *stream) {
    int rc = require_device();
    if (rc) return rc;
    return gen_xtrue(seed, i0, cnt, d_out, (cudaStream_t)stream);
}
int cudamat_gen_random_dd_device(int n, uint64_t seed, int *d_ia, int *d_ja, double *d_a, int64_t *nnz_out, void *stream) {
    int rc = require_device();
    if (rc) return rc;
    if (n <= 0 || !d_ia) { set_error("gen_random_dd: invalid argument"); return CUDAMAT_E_INVALID; }
    return gen_random_dd(n, seed, d_ia, d_ja, d_a, nnz_out, (cudaStream_t)stream);
}

// ---- multi-GPU: implemented in comm.cu --------------------------------------------------------

void cudamat_free(void *p) { free(p); }

}  // extern "C"
