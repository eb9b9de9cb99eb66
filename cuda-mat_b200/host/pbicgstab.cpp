// pbicgstab.cpp — the reference's three C++ entry points (pbicgstab.h:113,116,119-120) and host
// utilities (pbicgstab.cu:413-423,1093-1115) as thin wrappers over the C ABI (include/cudamat_b200.h).
#include "pbicgstab.h"
#include "../../include/cudamat_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

double rand_float_0_1() { return static_cast<double>(rand()) / static_cast<double>(RAND_MAX); }
double rand_float(double min, double max) { return rand_float_0_1() * (max - min) + min; }

void gen_rand_vector(int n, double *vector, double probability_of_zero, double min, double max) {
    for (int k = 0; k < n; ++k) {
        const bool zero = rand_float_0_1() <= probability_of_zero;
        vector[k] = zero ? 0.0 : rand_float(min, max);
    }
}

void toDenseVector(int n, int /*nnz*/, double *A, int *IA, double *out) {
    int seen = IA[0], next = 0;
    for (int row = 0; row < n; ++row) {
        const bool has_entry = IA[row + 1] > seen;
        out[row] = has_entry ? A[next] : 0.0;
        if (has_entry) { ++next; seen = IA[row + 1]; }
    }
}

static bool strict_compat() {
    const char *e = getenv("CUDAMAT_STRICT_COMPAT");
    return e && *e && strcmp(e, "0") != 0;
}

static bool run(int mode, int n, int nnz, double *A, int *iA, int *jA, double *d, double *x0, double *b, int maxit,
                double tol, bool debug, double *x, double *dtAlg) {
    cudamat_stats st;
    double dt = 0.0;
    const int rc = cudamat_bicgstab_host(mode, n, nnz, A, iA, jA, d, x0, b, maxit, tol, debug ? 1 : 0, x, &dt, &st);
    if (dtAlg) *dtAlg = dt;
    if (rc != CUDAMAT_OK) {
        fprintf(stderr, "!!!! cudamat: %s\n", cudamat_last_error());
        return false;
    }
    if (mode == CUDAMAT_MODE_ILU0 && strict_compat()) return true;      // pbicgstab.cu:408
    return st.converged != 0;
}

bool bicgstab(int n, int nnz, double *A, int *iA, int *jA, double *b, int maxit, double tol, bool debug, double *x,
              double *dtAlg) {
    return run(CUDAMAT_MODE_PLAIN, n, nnz, A, iA, jA, nullptr, nullptr, b, maxit, tol, debug, x, dtAlg);
}

bool bicgstab(int n, int nnz, double *A0, int *iA0, int *jA0, double *d, double *x0, double *b, int maxit, double tol,
              bool debug, double *x, double *dtAlg) {
    return run(CUDAMAT_MODE_SHIFTED, n, nnz, A0, iA0, jA0, d, x0, b, maxit, tol, debug, x, dtAlg);
}

bool bicgstab_lu_precond(int n, int nnz, double *A, int *iA, int *jA, double *b, int maxit, double tol, bool debug,
                         double *x, double *dtAlg) {
    return run(CUDAMAT_MODE_ILU0, n, nnz, A, iA, jA, nullptr, nullptr, b, maxit, tol, debug, x, dtAlg);
}
