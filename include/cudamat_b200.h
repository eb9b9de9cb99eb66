/*
 * cudamat_b200.h — C ABI of the B200-native BiCGSTAB path (libcudamat_b200.so).
 *
 * This is the drop-in boundary for the reference's pbicgstab path: plain pointers and sizes,
 * no C++/torch types.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference checkout).  The C++ mirror of the reference's own header
 * (cuda-mat_b200/host/pbicgstab.h) and the Python ctypes binding are thin wrappers over this.
 *
 * Conventions
 *   - return value: CUDAMAT_OK (0) or a negative CUDAMAT_E_* code; cudamat_last_error() returns
 *     a human readable message for the calling thread.  Nothing in this library calls exit()
 *     (the reference's checkCudaErrors does: helper_cuda.h:999-1010).
 *   - "host" entry points take host pointers owned by the caller (pbicgstab.h:96-110);
 *     "device" entry points borrow device pointers on the current CUDA device.
 *   - CSR is int32 / fp64; index base is read from iA[0] (pbicgstab.cu:201,782,953): 0 or 1.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     CUDAMAT_E_NO_DEVICE.
 */
#ifndef CUDAMAT_B200_H
#define CUDAMAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUDAMAT_ABI_VERSION 2

/* status codes */
#define CUDAMAT_OK              0
#define CUDAMAT_E_INVALID      -1   /* bad argument / malformed CSR                         */
#define CUDAMAT_E_NO_DEVICE    -2   /* no usable CUDA device (there is no CPU fallback)     */
#define CUDAMAT_E_CUDA         -3   /* a CUDA runtime call or kernel failed                 */
#define CUDAMAT_E_NO_DIAGONAL  -4   /* ILU0 needs a structurally non-zero diagonal (pbicgstab.h:118) */
#define CUDAMAT_E_IO           -5   /* Matrix Market file could not be read / unsupported   */
#define CUDAMAT_E_COMM         -6   /* NCCL not available / communicator failure            */
#define CUDAMAT_E_STATE        -7   /* call order violated (e.g. solve before analyze)      */

/* solver modes: the three reference entry points */
#define CUDAMAT_MODE_PLAIN    0     /* bicgstab(A,b)            pbicgstab.h:113 (intended algorithm) */
#define CUDAMAT_MODE_SHIFTED  1     /* bicgstab(A0,d,x0,b)      pbicgstab.h:116  */
#define CUDAMAT_MODE_ILU0     2     /* bicgstab_lu_precond      pbicgstab.h:119-120 */

/* breakdown codes in cudamat_stats.breakdown */
#define CUDAMAT_BRK_NONE      0
#define CUDAMAT_BRK_OMEGA     1     /* |omega| < 1e-5            pbicgstab.cu:559,735 */
#define CUDAMAT_BRK_NAN       2     /* isnan(omega)              pbicgstab.cu:559,735 */
#define CUDAMAT_BRK_MAXIT     3     /* maxit reached             pbicgstab.cu:575,751 */

/* SpMV kernel variants (cudamat_stats.spmv_variant, cudamat_set_option "spmv_variant") */
#define CUDAMAT_SPMV_AUTO      0    /* chosen from row-length statistics at analyze time     */
#define CUDAMAT_SPMV_ROWLANE   1    /* thread-per-row, direct global loads                   */
#define CUDAMAT_SPMV_STAGED    2    /* row-block staged through shared memory by TMA bulk copies */
#define CUDAMAT_SPMV_PATTERN   3    /* column offsets from a per-row class dictionary (1 B/row), values from CSR */
#define CUDAMAT_SPMV_CLASS     4    /* offsets AND values from the class dictionary: CSR arrays are not read   */
#define CUDAMAT_SPMV_TILED     5    /* CLASS + the x windows of a 2048-row tile staged in shared memory by TMA  */
#define CUDAMAT_SPMV_MARCH     6    /* plane-marching stencil kernel: 3 planes of x in a shared-memory ring, one new tile per step;
                                       in the unpreconditioned loop the p / s vector updates are folded into it (option "fuse") */
#define CUDAMAT_SPMV_STREAM    7    /* irregular rows: (col, val) streamed and x gathered entry-parallel per warp chunk, row
                                       chains out of shared memory (chosen from the row-length statistics)                   */

typedef struct cudamat_stats {
    int    iterations;      /* the reference's loop counter i at exit                          */
    int    converged;       /* 1 iff ||r|| < tol*||r0|| was observed                           */
    int    breakdown;       /* CUDAMAT_BRK_*                                                   */
    int    half_steps;      /* residual norms recorded (incl. ||r0||)                          */
    double nrm_r0;          /* ||b - A x0||_2                                                  */
    double nrm_r;           /* last residual norm                                              */
    double t_h2d;           /* seconds: host->device upload (host entry points only)           */
    double t_analysis;      /* seconds: SpMV plan + level analysis (csrsv_analysis pbicgstab.cu:335-350) */
    double t_ilu0;          /* seconds: ILU(0) factorisation (pbicgstab.cu:353-363)            */
    double t_loop;          /* seconds: iteration loop only == *dtAlg (pbicgstab.cu:365-374)   */
    double t_d2h;           /* seconds: device->host copy of x                                 */
    int    levels_l;        /* level count of the unit-lower factor (0 if not ILU0)            */
    int    levels_u;
    int    spmv_variant;    /* variant actually used                                           */
    int    zero_pivot;      /* 0; <0: -(1+row) of the first exact-zero ILU0 pivot              */
    int64_t kernel_launches;/* kernels of this library launched by the call                    */
    double t_spmv;          /* seconds: sum of the SpMV kernel durations inside the loop, measured with
                               CUDA events on the launching stream (option "time_spmv" = k > 0: the SpMVs of every
                               k-th iteration are timed), else 0 */
    int    n_spmv;          /* SpMV launches covered by t_spmv                                  */
    int    graph_replay;    /* 1 iff the iteration loop ran as CUDA-graph replays of poll_every-iteration batches */
    /* ABI 2: per-kernel event timings of the sampled iterations ("time_spmv"): [0] SpMV 1 (MARCH fused loop: with the
     * folded p update), [1] SpMV 2 (with the folded s update), [2] x / r update + dots, [3] separate p / s updates */
    double t_kernel[4];
    int    n_kernel[4];
    int    fused;           /* bit 0: the p update ran folded into SpMV 1, bit 1: the s update into SpMV 2 (MARCH);
                               4: the whole iteration ran inside the persistent cooperative kernel           */
} cudamat_stats;

typedef struct cudamat_solver cudamat_solver;   /* opaque per-matrix handle */

/* ---------------------------------------------------------------------------------------------
 * Library
 * ------------------------------------------------------------------------------------------- */
int         cudamat_abi_version(void);
const char *cudamat_last_error(void);
/* number of visible CUDA devices (0 if none / driver missing); never fails */
int         cudamat_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * One-shot host-pointer solves: what the reference's three C++ entry points bind to.
 *   replaces bicgstab            pbicgstab.cu:756-922   (mode PLAIN:   d = x0 = NULL, x0 := ones)
 *            bicgstab (shifted)  pbicgstab.cu:926-1088  (mode SHIFTED: d, x0 required)
 *            bicgstab_lu_precond pbicgstab.cu:157-409   (mode ILU0:    d = x0 = NULL, x0 := ones)
 * x: caller-allocated n doubles; dtAlg (may be NULL) gets the loop seconds; st may be NULL.
 * debug != 0 prints the reference's trace lines (pbicgstab.cu:76,113,144,203,349-363,484,550).
 * The host arrays may be pageable (malloc'ed, as example.cpp:96-104,252 passes them) or pinned: pageable arrays of 8 MB or
 * more travel through a threaded pinned staging area (csrc/hostcopy.cpp), pinned ones are copied directly.
 * ------------------------------------------------------------------------------------------- */
int cudamat_bicgstab_host(int mode, int n, int nnz, const double *A, const int *iA, const int *jA,
                          const double *d, const double *x0, const double *b,
                          int maxit, double tol, int debug, double *x, double *dtAlg,
                          cudamat_stats *st);

/* ILU(0) factor of a host CSR in A's pattern (replaces cusparseDcsrilu0 pbicgstab.cu:359 + the
 * analysis at :338,345). M_out: nnz doubles (unit-L strictly lower, U upper incl. diagonal).
 * levels[0..1] (may be NULL) get the L / U level counts. */
int cudamat_ilu0_host(int n, int nnz, const double *A, const int *iA, const int *jA,
                      double *M_out, int *levels, int *zero_pivot);

/* ---------------------------------------------------------------------------------------------
 * Handle API (device resident; used by bench.py, the tests and the multi-GPU path).
 * A handle owns one row shard [row0,row1) of an n_global x n_global matrix; a single-GPU solve
 * has row0 = 0, row1 = n_global.  stream is a cudaStream_t (NULL = legacy default stream).
 * ------------------------------------------------------------------------------------------- */
int cudamat_create(cudamat_solver **out, int64_t n_global, int64_t row0, int64_t row1, void *stream);
int cudamat_destroy(cudamat_solver *s);
/* option keys: "spmv_variant" (CUDAMAT_SPMV_*), "poll_every" (iterations between status polls), "sptrsv_syncfree" (0: one
 * launch per level), "sptrsv_no_smem" (1: never use the single-CTA shared-memory sweeps), "sptrsv_ring" (0: small systems use the round-1
 * barrier-per-level kernel instead of the ring kernel), "sptrsv_blocked" (0: no block-wavefront sweeps on grid stencils), "sptrsv_ctas_per_sm",
 * "ilu0_reorder" (1: multicolour ordering of the preconditioner matrix — few sweep levels, a different ILU(0), opt-in),
 * "host_analysis" (1: ILU0 level analysis on the host, cross-check), "graph" (-1 auto, 0 off, 1 force CUDA-graph replay),
 * "debug", "time_spmv" (k: event-time the main kernels of every k-th iteration), "fuse" (bit 0: fold the p update into MARCH SpMV 1,
 * bit 1: the s update into SpMV 2; 0: never), "resume" (1: the next solve continues the previous one for maxit more iterations), "march_grid" (CTAs), "stream_blocks" (column
 * blocks of the STREAM variant: 0 auto, 1 never), "persist" (unpreconditioned loop as one persistent cooperative kernel per batch of
 * poll_every iterations: -1 auto = small systems / shards, 0 off, 1 force) */
int cudamat_set_option(cudamat_solver *s, const char *key, int64_t value);

/* CSR rows of this shard with GLOBAL column indices (cusparseDcsrmv operand pbicgstab.cu:67).
 * _host uploads (and normalises base-1 to base-0); _device borrows base-0 device arrays, which
 * must stay alive and unmodified until destroy. */
int cudamat_set_csr_host(cudamat_solver *s, int nnz, const double *A, const int *iA, const int *jA);
int cudamat_set_csr_device(cudamat_solver *s, int64_t nnz, const double *dA, const int *dIA, const int *dJA);

/* SpMV plan from row-length statistics; for CUDAMAT_MODE_ILU0 also level analysis + factorisation
 * (cusparseDcsrsv_analysis pbicgstab.cu:338,345; cusparseDcsrilu0 :359).  On a sharded handle (after
 * cudamat_comm_init) MODE_ILU0 builds a block-Jacobi ILU(0) of the shard's diagonal block. */
int cudamat_analyze(cudamat_solver *s, int mode, cudamat_stats *st);

/* the iteration (gpu_pbicgstab pbicgstab.cu:45-154 / gpu_pbicgstab2 :581-754) on device vectors of
 * this shard's rows. d_x0 NULL = ones; d_d NULL = no diagonal shift. */
int cudamat_solve_device(cudamat_solver *s, int mode, const double *d_b, const double *d_x0,
                         const double *d_d, double *d_x, int maxit, double tol, cudamat_stats *st);
/* residual-norm history of the last solve (||r0|| first); returns entries copied */
int cudamat_get_history(cudamat_solver *s, double *hist, int cap);

/* ---- kernel-level entry points (parity tests and the SpMV GB/s metric) -------------------- */
/* y = A x (+ d.*x) with the planned (variant 0) or a forced variant; single-shard handles only */
int cudamat_spmv_device(cudamat_solver *s, const double *d_x, const double *d_d, double *d_y, int variant);
/* spec dot product (replaces cublasDdot / cublasDnrm2^2) */
int cudamat_dot_device(cudamat_solver *s, const double *d_a, const double *d_b, double *result);
/* copy the ILU(0) factor values (nnz doubles, A's pattern) to a host buffer */
int cudamat_get_ilu0_host(cudamat_solver *s, double *M_out);
/* out = L^{-1} rhs (upper = 0, unit diagonal) or U^{-1} rhs (upper = 1) on the analysed factor */
int cudamat_sptrsv_device(cudamat_solver *s, int upper, const double *d_rhs, double *d_out);
/* number of 16^3 blocks of the block-wavefront sweep plan (7-point grid factors, csrc/sweepblk.cu); 0: the generic sweeps run */
int cudamat_sweep_blocks(cudamat_solver *s);

/* ---- multi-GPU (one process per GPU; NCCL is dlopen()ed on first use) ----------------------- */
/* The reference is single-GPU; sharding follows BASELINE.json's north_star: contiguous row
 * shards, halo entries of the SpMV operand exchanged point-to-point, one fused small allreduce per
 * reduction point.  The two planners are pure host code. */
int cudamat_partition_rows(int64_t n_global, int world, int rank, int64_t *row0, int64_t *row1);
int cudamat_halo_plan_host(int64_t row0, int64_t row1, int64_t nnz, const int *ja_global, int world,
                           const int64_t *row_starts, int *nhalo, int **halo_cols, int *recv_cnt);
/* Host planner of the TILED SpMV variant (pure host code, no device needed; used by cudamat_analyze and exported for
 * the CPU tests).  Input: the row-class dictionary — ncls <= 64 classes of len[c] <= 16 entries, column offsets
 * off[c*16+q] = ja - row in storage order, values val[c*16+q] (may be NULL), rows per class hist[c] — and the row count.
 * Output (arrays sized 4, 4, 4, ncls*16, -, -, 8, 8, 64): the x windows staged per 2048-row tile (first offset, length
 * and shared-memory base in elements), the shared-memory index of every entry relative to its row, the bit mask of the
 * classes that fit the windows, and the superset pattern (sup_len = 0: none): byte offset and value per pattern entry
 * and the presence mask of every class.  *nseg = 0: no plan (too many / too wide windows).  No reference counterpart:
 * the reference hands CSR to cusparseDcsrmv (pbicgstab.cu:67,104,132). */
int cudamat_tiled_plan_host(int ncls, const int *len, const int *off, const double *val, const unsigned *hist, int n,
                            int with_vals, int *nseg, int *seg_lo, int *seg_len, int *seg_base, int *disp,
                            unsigned long long *ok_mask, int *sup_len, int *sup_boff, double *sup_val,
                            unsigned char *class_mask, long long *smem_bytes);
/* Host planner of the MARCH SpMV variant (pure host code; used by cudamat_analyze and exported for the CPU tests).  Input: the
 * superset pattern of the TILED plan (ascending column offsets, values) and the row count.  *ok = 1 when the offsets split into
 * planes -D / 0 / +D (D a multiple of the 2048-row tile, n a multiple of D, in-plane offsets within H <= 512): then D, H, tiles
 * per plane S, planes P and per entry the plane (dz[8] in {-1,0,1}) and in-plane offset (loff[8]).  No reference counterpart. */
int cudamat_march_plan_host(int sup_len, const int *sup_off, const double *sup_val, long long n, int *ok, int *D, int *H,
                            int *S, int *P, int *dz, int *loff);
#define CUDAMAT_UNIQUE_ID_BYTES 128
int cudamat_comm_unique_id(void *id128);
int cudamat_comm_init(cudamat_solver *s, const void *id128, int rank, int world);
/* 1 when the handle uses the peer-memory path (CUDA IPC over NVLink: halo rows stored into the neighbours' vectors
 * by the producing kernels, partial sums gathered by the reducing kernels — no NCCL call inside the iteration),
 * 0 when it uses ncclSend/Recv + ncclAllReduce (CUDAMAT_NO_P2P=1, IPC unavailable, non-contiguous halo sends). */
int cudamat_comm_p2p_enabled(cudamat_solver *s);

/* ---- device-side synthetic inputs ------------------------------- */
/* rows [row0,row1) of the N^3 7-point Dirichlet Poisson matrix, base-0, global column ids.
 * d_ia: row1-row0+1 ints (local offsets from 0). Pass d_ja = d_a = NULL to only fill d_ia. */
int cudamat_gen_poisson3d_device(int N, int64_t row0, int64_t row1, int *d_ia, int *d_ja, double *d_a, void *stream);
int64_t cudamat_poisson3d_nnz(int N, int64_t row0, int64_t row1);
/* x_true[i] in (-1,1) from a counter hash of (seed, i), i in [i0, i0+cnt) */
int cudamat_gen_xtrue_device(uint64_t seed, int64_t i0, int64_t cnt, double *d_out, void *stream);
/* generator.cpp-style random nonsymmetric diagonally-dominant CSR, irregular row lengths.
 * two-pass: with d_ja == NULL fills d_ia (n+1) and returns nnz through *nnz_out. */
int cudamat_gen_random_dd_device(int n, uint64_t seed, int *d_ia, int *d_ja, double *d_a, int64_t *nnz_out, void *stream);

/* ---- Matrix Market loading (replaces loadMMSparseMatrix mmio_wrapper.h:133-348) ------------ */
/* arrays are malloc()ed; release with cudamat_free. csr_format 0 = CSC like the reference. */
int  cudamat_load_mm(const char *filename, int csr_format, int *m, int *n, int *nnz,
                     double **aVal, int **aRowInd, int **aColInd);
void cudamat_free(void *p);
/* ---- Matrix Market writing (replaces mm_write_banner / mm_write_mtx_crd, mmio.c:405-445,447-510) ----
 * CSR (index base = rowptr[0], 0 or 1) -> coordinate real general, or (symmetric != 0) the lower triangle as
 * coordinate real symmetric; 1-based, %.17g (exact round trip).  comment may be NULL. */
int  cudamat_write_mm(const char *filename, int m, int n, int nnz, const double *val, const int *rowptr,
                      const int *colind, int symmetric, const char *comment);
/* dense vector as the n x 1 coordinate matrix the reference's -V switch reads (example.cpp:310-336) */
int  cudamat_write_mm_vector(const char *filename, int n, const double *x, const char *comment);

#ifdef __cplusplus
}
#endif
#endif /* CUDAMAT_B200_H */
