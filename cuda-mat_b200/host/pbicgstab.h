// pbicgstab.h — C++ mirror of the reference's public header (pbicgstab.h:23-120) on top of the
// C ABI in include/cudamat_b200.h.  Same names, argument meaning and defaults, so a caller of the
// reference (example.cpp:87,146,352) builds against this header unchanged.  Differences, all
// deliberate:
//   * no <conio.h> (Windows only) and no dependency on cuSPARSE/cuBLAS;
//   * bicgstab(A,b) implements the INTENDED algorithm (the reference's is broken: pbicgstab.cu:469-478);
//   * bicgstab_lu_precond returns true iff the iteration converged (the reference returns true
//     unconditionally, pbicgstab.cu:408); export CUDAMAT_STRICT_COMPAT=1 to get the old behaviour;
//   * a CUDA failure prints the message and returns false instead of exit(EXIT_FAILURE).
#pragma once

#include <cmath>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

enum Base { Base0 = 0, Base1 = 1 };

// rand()/RAND_MAX in [0,1] and its affine map to [min,max] (pbicgstab.cu:413-423); libc rand(), never seeded
double rand_float_0_1();
double rand_float(double min, double max);

// dense n x m sweep, entry (i,j) is zero with the given probability, |value| >= eps (pbicgstab.h:31-54)
template <Base base>
int gen_rand_csr_matrix(int n, int m, std::vector<double> *A, std::vector<int> *IA, std::vector<int> *JA,
                        double probability_of_zero, double min, double max, double eps) {
    int next = base;
    IA->push_back(next);
    for (int row = 0; row < n; ++row) {
        for (int col = 0; col < m; ++col) {
            if (rand_float_0_1() <= probability_of_zero) continue;
            double v = rand_float(min, max);
            while (std::fabs(v) < eps) v = rand_float(min, max);
            A->push_back(v);
            JA->push_back(col + base);
            ++next;
        }
        IA->push_back(next);
    }
    return static_cast<int>(A->size());
}

// dense n x m sweep over a callback, entries with |f(i,j)| <= eps are dropped (pbicgstab.h:56-76)
template <Base base>
int fill_csr_matrix(int n, int m, std::vector<double> *A, std::vector<int> *IA, std::vector<int> *JA,
                    std::function<double(int, int)> f, double eps) {
    int next = base;
    IA->push_back(next);
    for (int row = 0; row < n; ++row) {
        for (int col = 0; col < m; ++col) {
            const double v = f(row, col);
            if (std::fabs(v) > eps) {
                A->push_back(v);
                JA->push_back(col + base);
                ++next;
            }
        }
        IA->push_back(next);
    }
    return static_cast<int>(A->size());
}

void gen_rand_vector(int n, double *vector, double probability_of_zero, double min, double max);

// "(v0 v1 ... )" with std::to_string formatting (pbicgstab.h:79-88)
template <typename T>
void dump_vector(std::ostringstream &stream, int n, T *vector) {
    stream << "(";
    for (int k = 0; k < n; ++k) stream << std::to_string(vector[k]) << " ";
    stream << ")";
}

// m x 1 CSR column vector -> dense (pbicgstab.cu:1101-1115)
void toDenseVector(int n, int nnz, double *A, int *IA, double *out);

#define IN   // function argument marked as `input`
#define OUT  // function argument marked as `output`

// Common arguments (pbicgstab.h:96-110): n, nnz, CSR (A, iA, jA) with iA[0] = index base (0 or 1),
// b dense right-hand side, maxit, tol (||r|| < tol*||r0||), debug trace, x caller-allocated output,
// dtAlg = seconds spent in the iteration loop on the GPU (I/O not accounted).

// solve Ax = b, no preconditioner, x0 = ones
bool bicgstab(int n, int nnz, double IN(*A), int IN(*iA), int IN(*jA), double IN(*b), int maxit, double tol,
              bool debug, double OUT(*x), double OUT(*dtAlg));
// solve (A0 + I*d)x = b from the caller's x0, no preconditioner
bool bicgstab(int n, int nnz, double IN(*A0), int IN(*iA0), int IN(*jA0), double IN(*d), double IN(*x0),
              double IN(*b), int maxit, double tol, bool debug, double OUT(*x), double OUT(*dtAlg));
// solve Ax = b with the ILU0 right preconditioner; A[i,i] must be structurally non-zero; x0 = ones
bool bicgstab_lu_precond(int n, int nnz, double IN(*A), int IN(*iA), int IN(*jA), double IN(*b), int maxit,
                         double tol, bool debug, double OUT(*x), double OUT(*dtAlg));
