// C entry points around the REFERENCE's own GPU solver, compiled UNMODIFIED from /root/reference/pbicgstab.cu
// (see oracle/Makefile target _ref/libref_pbicgstab.so and ref_shims/refgpu/legacy_cusparse.h).
// Test / baseline infrastructure only: loaded by tests/ and by bench.py's reference legs, never by the product.
//
// The reference returns neither the iteration count nor the residuals (pbicgstab.cu:408), it only prints them under
// debug (pbicgstab.cu:76,113,144,484,550).  With trace_path != NULL the call runs with debug = true and fd 1 is pointed
// at that file for its duration, so the caller can parse the reference's own trace.
#include <cstdio>
#include <iostream>
#include <fcntl.h>
#include <unistd.h>
#include "pbicgstab.h"          // the reference's header (-I$(REF)); its <conio.h> resolves to ref_shims/refgpu/conio.h

namespace {
struct Redirect {
    int saved = -1;
    explicit Redirect(const char *path) {
        if (!path) return;
        fflush(stdout); std::cout.flush();
        int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd < 0) return;
        saved = dup(1);
        dup2(fd, 1);
        close(fd);
    }
    ~Redirect() {
        if (saved < 0) return;
        fflush(stdout); std::cout.flush();
        dup2(saved, 1);
        close(saved);
    }
};
}

// bicgstab_lu_precond  pbicgstab.h:119-120 / pbicgstab.cu:157
extern "C" int ref_bicgstab_lu_precond(int n, int nnz, double *A, int *iA, int *jA, double *b, int maxit, double tol,
                                       double *x, double *dtAlg, const char *trace_path) {
    Redirect r(trace_path);
    return bicgstab_lu_precond(n, nnz, A, iA, jA, b, maxit, tol, trace_path != nullptr, x, dtAlg) ? 1 : 0;
}
// bicgstab (A0 + diag(d), caller x0)  pbicgstab.h:116 / pbicgstab.cu:926
extern "C" int ref_bicgstab_shifted(int n, int nnz, double *A0, int *iA0, int *jA0, double *d, double *x0, double *b,
                                    int maxit, double tol, double *x, double *dtAlg, const char *trace_path) {
    Redirect r(trace_path);
    return bicgstab(n, nnz, A0, iA0, jA0, d, x0, b, maxit, tol, trace_path != nullptr, x, dtAlg) ? 1 : 0;
}
// bicgstab (plain)  pbicgstab.h:113 / pbicgstab.cu:756 — broken initial residual in the reference (pbicgstab.cu:469-478)
extern "C" int ref_bicgstab_plain(int n, int nnz, double *A, int *iA, int *jA, double *b, int maxit, double tol,
                                  double *x, double *dtAlg, const char *trace_path) {
    Redirect r(trace_path);
    return bicgstab(n, nnz, A, iA, jA, b, maxit, tol, trace_path != nullptr, x, dtAlg) ? 1 : 0;
}
