// mmio_wrapper.h shim: same entry point as the reference's loader (mmio_wrapper.h:133-142),
// implemented on cudamat_load_mm (csrc/mmload.cpp). elem_type must be 'd' (fp64), as in every
// call site of the reference (example.cpp:40,48,56,116,124,252,313). Arrays are malloc()ed; the
// caller free()s them (example.cpp:96-104,372-374). Returns 0 on success, 1 on failure.
#pragma once
#include <cstdio>
#include "../../../include/cudamat_b200.h"

inline int loadMMSparseMatrix(char *filename, char elem_type, bool csrFormat, int *m, int *n, int *nnz,
                              double **aVal, int **aRowInd, int **aColInd) {
    if (elem_type != 'd' && elem_type != 'D') {
        fprintf(stderr, "!!!! only elem_type 'd' is supported\n");
        return 1;
    }
    const int rc = cudamat_load_mm(filename, csrFormat ? 1 : 0, m, n, nnz, aVal, aRowInd, aColInd);
    if (rc != CUDAMAT_OK) {
        fprintf(stderr, "%s\n", cudamat_last_error());
        return 1;
    }
    return 0;
}
