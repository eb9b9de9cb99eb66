// Shim that exposes the REFERENCE's own loader (mmio_wrapper.h:133-348, compiled from
// /root/reference where it lies) through a C symbol, so tests/golden/make_golden.py can
// record what CSR the reference produces for each fixture.  Test infrastructure only.
#include <stdio.h>
#include <stdlib.h>
#include "mmio.h"
#include "mmio_wrapper.h"

extern "C" int ref_loadMMSparseMatrix(const char *filename, int *m, int *n, int *nnz,
                                      double **aVal, int **aRowInd, int **aColInd) {
    return loadMMSparseMatrix(const_cast<char *>(filename), 'd', true, m, n, nnz,
                              aVal, aRowInd, aColInd);
}
extern "C" void ref_free(void *p) { free(p); }
