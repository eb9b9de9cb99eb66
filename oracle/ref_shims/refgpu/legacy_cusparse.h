/*
 * legacy_cusparse.h — TEST-ONLY shim.  Lets the reference's UNMODIFIED pbicgstab.cu compile against CUDA 12.9.
 *
 * The reference calls six cuSPARSE entry points that were removed in CUDA 11:
 *   cusparseSolveAnalysisInfo_t / cusparseCreate|DestroySolveAnalysisInfo   pbicgstab.cu:331-332,377-378
 *   cusparseDcsrsv_analysis                                                  pbicgstab.cu:338,345
 *   cusparseDcsrsv_solve                                                     pbicgstab.cu:94,98,123,127
 *   cusparseDcsrilu0                                                         pbicgstab.cu:359
 *   cusparseDcsrmv                                                           pbicgstab.cu:67,104,132,469,...,704
 * This header re-creates exactly those names on top of NVIDIA's own successors in the SAME library
 * (cusparseSpSV, cusparseDcsrilu02, cusparseSpMV), honouring the legacy descriptor's index base / fill mode /
 * diag type.  The arithmetic therefore still happens inside closed-source cuSPARSE/cuBLAS, as in the reference.
 *
 * Force-included (nvcc -include) when oracle/Makefile builds oracle/_ref/libref_pbicgstab.so from
 * /root/reference/pbicgstab.cu.  Nothing in the product includes, links or loads this.
 */
#ifndef ORACLE_LEGACY_CUSPARSE_H
#define ORACLE_LEGACY_CUSPARSE_H

#include <cuda_runtime.h>
#include <cusparse.h>
#include <stdlib.h>
#include <stdio.h>

struct legacyAnalysisInfo {
    /* SpSV state, built lazily at the first solve: the reference analyses with A's values (pbicgstab.cu:338)
     * and solves with the factor M (:94); the legacy analysis only looked at the pattern, SpSV_analysis
     * may look at values, so it has to see the array the solve will use. */
    cusparseSpMatDescr_t mat;
    cusparseSpSVDescr_t  spsv;
    cusparseDnVecDescr_t vx, vy;
    void                *buf;
    const double        *vals;          /* value array the SpSV analysis was run on */
    const void          *px, *py;
    cusparseFillMode_t   fill;
    cusparseDiagType_t   diag;
    /* csrilu02 state */
    csrilu02Info_t       ilu;
    void                *ilu_buf;
};
typedef struct legacyAnalysisInfo *cusparseSolveAnalysisInfo_t;

static inline cusparseStatus_t cusparseCreateSolveAnalysisInfo(cusparseSolveAnalysisInfo_t *info) {
    *info = (cusparseSolveAnalysisInfo_t)calloc(1, sizeof(struct legacyAnalysisInfo));
    return *info ? CUSPARSE_STATUS_SUCCESS : CUSPARSE_STATUS_ALLOC_FAILED;
}

static inline void legacy_release_spsv(cusparseSolveAnalysisInfo_t info) {
    if (info->spsv) cusparseSpSV_destroyDescr(info->spsv);
    if (info->mat)  cusparseDestroySpMat(info->mat);
    if (info->vx)   cusparseDestroyDnVec(info->vx);
    if (info->vy)   cusparseDestroyDnVec(info->vy);
    if (info->buf)  cudaFree(info->buf);
    info->spsv = 0; info->mat = 0; info->vx = 0; info->vy = 0; info->buf = 0; info->vals = 0;
}

static inline cusparseStatus_t cusparseDestroySolveAnalysisInfo(cusparseSolveAnalysisInfo_t info) {
    if (!info) return CUSPARSE_STATUS_SUCCESS;
    legacy_release_spsv(info);
    if (info->ilu)     cusparseDestroyCsrilu02Info(info->ilu);
    if (info->ilu_buf) cudaFree(info->ilu_buf);
    free(info);
    return CUSPARSE_STATUS_SUCCESS;
}

/* legacy analysis: pattern only.  Records the triangle the caller selected on the descriptor. */
static inline cusparseStatus_t cusparseDcsrsv_analysis(cusparseHandle_t, cusparseOperation_t, int, int,
                                                       const cusparseMatDescr_t descr, const double *, const int *,
                                                       const int *, cusparseSolveAnalysisInfo_t info) {
    info->fill = cusparseGetMatFillMode(descr);
    info->diag = cusparseGetMatDiagType(descr);
    legacy_release_spsv(info);
    return CUSPARSE_STATUS_SUCCESS;
}

#define LEGACY_TRY(call) do { cusparseStatus_t st__ = (call); if (st__ != CUSPARSE_STATUS_SUCCESS) return st__; } while (0)

/* op(A) x = alpha f on the triangle/diag type currently set on the descriptor */
static inline cusparseStatus_t cusparseDcsrsv_solve(cusparseHandle_t h, cusparseOperation_t op, int m, const double *alpha,
                                                    const cusparseMatDescr_t descr, const double *val, const int *rowptr,
                                                    const int *colind, cusparseSolveAnalysisInfo_t info,
                                                    const double *f, double *x) {
    cusparseFillMode_t fill = cusparseGetMatFillMode(descr);
    cusparseDiagType_t diag = cusparseGetMatDiagType(descr);
    if (!info->spsv || info->vals != val || info->fill != fill || info->diag != diag) {
        legacy_release_spsv(info);
        info->fill = fill; info->diag = diag;
        int nnz = 0, first = 0;
        if (cudaMemcpy(&nnz, rowptr + m, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return CUSPARSE_STATUS_EXECUTION_FAILED;
        if (cudaMemcpy(&first, rowptr, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return CUSPARSE_STATUS_EXECUTION_FAILED;
        nnz -= first;
        LEGACY_TRY(cusparseCreateCsr(&info->mat, m, m, nnz, (void *)rowptr, (void *)colind, (void *)val, CUSPARSE_INDEX_32I,
                                     CUSPARSE_INDEX_32I, cusparseGetMatIndexBase(descr), CUDA_R_64F));
        LEGACY_TRY(cusparseSpMatSetAttribute(info->mat, CUSPARSE_SPMAT_FILL_MODE, &fill, sizeof(fill)));
        LEGACY_TRY(cusparseSpMatSetAttribute(info->mat, CUSPARSE_SPMAT_DIAG_TYPE, &diag, sizeof(diag)));
        LEGACY_TRY(cusparseCreateDnVec(&info->vx, m, (void *)f, CUDA_R_64F));
        LEGACY_TRY(cusparseCreateDnVec(&info->vy, m, (void *)x, CUDA_R_64F));
        info->px = f; info->py = x;
        LEGACY_TRY(cusparseSpSV_createDescr(&info->spsv));
        size_t sz = 0;
        LEGACY_TRY(cusparseSpSV_bufferSize(h, op, alpha, info->mat, info->vx, info->vy, CUDA_R_64F, CUSPARSE_SPSV_ALG_DEFAULT, info->spsv, &sz));
        if (cudaMalloc(&info->buf, sz ? sz : 16) != cudaSuccess) return CUSPARSE_STATUS_ALLOC_FAILED;
        LEGACY_TRY(cusparseSpSV_analysis(h, op, alpha, info->mat, info->vx, info->vy, CUDA_R_64F, CUSPARSE_SPSV_ALG_DEFAULT, info->spsv, info->buf));
        info->vals = val;
    }
    if (info->px != f) { LEGACY_TRY(cusparseDnVecSetValues(info->vx, (void *)f)); info->px = f; }
    if (info->py != x) { LEGACY_TRY(cusparseDnVecSetValues(info->vy, (void *)x)); info->py = x; }
    return cusparseSpSV_solve(h, op, alpha, info->mat, info->vx, info->vy, CUDA_R_64F, CUSPARSE_SPSV_ALG_DEFAULT, info->spsv);
}

/* in-place ILU(0) on csrValM in the pattern (rowptr, colind); no pivoting, no boost — like the legacy routine */
static inline cusparseStatus_t cusparseDcsrilu0(cusparseHandle_t h, cusparseOperation_t, int m, const cusparseMatDescr_t descrA,
                                                double *csrValM, const int *rowptr, const int *colind,
                                                cusparseSolveAnalysisInfo_t info) {
    int nnz = 0, first = 0;
    if (cudaMemcpy(&nnz, rowptr + m, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return CUSPARSE_STATUS_EXECUTION_FAILED;
    if (cudaMemcpy(&first, rowptr, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return CUSPARSE_STATUS_EXECUTION_FAILED;
    nnz -= first;
    /* csrilu02 wants a GENERAL descriptor without fill/diag attributes: use a private one with the caller's base */
    cusparseMatDescr_t d = 0;
    LEGACY_TRY(cusparseCreateMatDescr(&d));
    cusparseSetMatType(d, CUSPARSE_MATRIX_TYPE_GENERAL);
    cusparseSetMatIndexBase(d, cusparseGetMatIndexBase(descrA));
    if (info->ilu) { cusparseDestroyCsrilu02Info(info->ilu); info->ilu = 0; }
    if (info->ilu_buf) { cudaFree(info->ilu_buf); info->ilu_buf = 0; }
    LEGACY_TRY(cusparseCreateCsrilu02Info(&info->ilu));
    int bs = 0;
    LEGACY_TRY(cusparseDcsrilu02_bufferSize(h, m, nnz, d, csrValM, rowptr, colind, info->ilu, &bs));
    if (cudaMalloc(&info->ilu_buf, bs > 0 ? bs : 16) != cudaSuccess) return CUSPARSE_STATUS_ALLOC_FAILED;
    LEGACY_TRY(cusparseDcsrilu02_analysis(h, m, nnz, d, csrValM, rowptr, colind, info->ilu, CUSPARSE_SOLVE_POLICY_USE_LEVEL, info->ilu_buf));
    cusparseStatus_t st = cusparseDcsrilu02(h, m, nnz, d, csrValM, rowptr, colind, info->ilu, CUSPARSE_SOLVE_POLICY_USE_LEVEL, info->ilu_buf);
    cusparseDestroyMatDescr(d);
    return st;
}

/* y = alpha op(A) x + beta y.  Descriptors and the work buffer are cached per (values, x, y) triple: the reference calls
 * this 2-3 times per iteration with a handful of operand combinations. */
struct legacySpmvSlot { const void *val, *rp, *ci, *x, *y; int m, n, nnz, base; cusparseSpMatDescr_t mat; cusparseDnVecDescr_t vx, vy; void *buf; size_t bufsz; };
static inline cusparseStatus_t cusparseDcsrmv(cusparseHandle_t h, cusparseOperation_t op, int m, int n, int nnz, const double *alpha,
                                              const cusparseMatDescr_t descr, const double *val, const int *rowptr, const int *colind,
                                              const double *x, const double *beta, double *y) {
    static struct legacySpmvSlot slots[16];
    static int used = 0, next = 0;
    struct legacySpmvSlot *s = 0;
    for (int i = 0; i < used; ++i)
        if (slots[i].val == val && slots[i].rp == rowptr && slots[i].ci == colind && slots[i].x == x && slots[i].y == y && slots[i].m == m &&
            slots[i].n == n && slots[i].nnz == nnz && slots[i].base == (int)cusparseGetMatIndexBase(descr)) { s = &slots[i]; break; }
    if (!s) {
        if (used < 16) s = &slots[used++];
        else {
            s = &slots[next]; next = (next + 1) % 16;
            cusparseDestroySpMat(s->mat); cusparseDestroyDnVec(s->vx); cusparseDestroyDnVec(s->vy); cudaFree(s->buf);
        }
        s->val = val; s->rp = rowptr; s->ci = colind; s->x = x; s->y = y; s->m = m; s->n = n; s->nnz = nnz;
        s->base = (int)cusparseGetMatIndexBase(descr); s->buf = 0; s->bufsz = 0;
        int xn = (op == CUSPARSE_OPERATION_NON_TRANSPOSE) ? n : m, yn = (op == CUSPARSE_OPERATION_NON_TRANSPOSE) ? m : n;
        LEGACY_TRY(cusparseCreateCsr(&s->mat, m, n, nnz, (void *)rowptr, (void *)colind, (void *)val, CUSPARSE_INDEX_32I,
                                     CUSPARSE_INDEX_32I, cusparseGetMatIndexBase(descr), CUDA_R_64F));
        LEGACY_TRY(cusparseCreateDnVec(&s->vx, xn, (void *)x, CUDA_R_64F));
        LEGACY_TRY(cusparseCreateDnVec(&s->vy, yn, (void *)y, CUDA_R_64F));
        LEGACY_TRY(cusparseSpMV_bufferSize(h, op, alpha, s->mat, s->vx, beta, s->vy, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, &s->bufsz));
        if (cudaMalloc(&s->buf, s->bufsz ? s->bufsz : 16) != cudaSuccess) return CUSPARSE_STATUS_ALLOC_FAILED;
    }
    return cusparseSpMV(h, op, alpha, s->mat, s->vx, beta, s->vy, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, s->buf);
}

#endif /* ORACLE_LEGACY_CUSPARSE_H */
