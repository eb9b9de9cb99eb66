// Shim that links the REFERENCE's OpenMP CPU solver (bicstab_omp/bicstab.cpp BiCG(), :93-196)
// in-process so bench.py can time it on the GPU box's host cores without going through its
// text file format.  BICSTAB_SRC is the throw-away patched temp copy made by
// oracle/Makefile.  Test / baseline infrastructure only.
#define main bicstab_omp_main
#include BICSTAB_SRC
#undef main
#include <omp.h>

// BiCG takes crsMatrix by value (shallow) and x0 = ones internally (bicstab.cpp:139-140).
extern "C" int ref_bicg(int n, int nz, double *val, int *col, int *rowindex,
                        double *b, double *x, int maxit, int *iters) {
    crsMatrix A;
    A.N = n; A.NZ = nz; A.Value = val; A.Col = col; A.RowIndex = rowindex;
    int it = 0;
    int rc = BiCG(A, b, x, maxit, it);
    *iters = it;
    return rc;
}
extern "C" int ref_omp_threads(void) { return omp_get_max_threads(); }
extern "C" void ref_omp_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
