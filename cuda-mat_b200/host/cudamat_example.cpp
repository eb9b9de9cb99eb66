// cudamat_example.cpp — demo CLI with the reference's switches (example.cpp:185-223):
//   -M<matrix.mtx> -V<vector.mtx> -D (debug trace) -R<prob of zero> -N<dim> -P (print x) device=<n>
// plus -T<tol> -I<maxit> -U (unpreconditioned instead of ILU0) -O<x.mtx> (write the solution as an n x 1 Matrix
// Market file, the format -V reads) -W<A.mtx> (write the system matrix, e.g. the random default problem).  Unlike the reference's main
// (example.cpp:169,377) the exit status is 0 on success.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "helper_cuda.h"
#include "helper_cusolver.h"
#include "mmio_wrapper.h"
#include "pbicgstab.h"
#include "../../include/cudamat_b200.h"

int main(int argc, char *argv[]) {
    const char *matrix_file = nullptr, *vector_file = nullptr, *out_file = nullptr, *mat_out_file = nullptr;
    bool debug = false, print = false, unprec = false;
    double p_zero_mat = 0.99, p_zero_vec = 0.2, tol = 1e-6;
    int dim = 10000, maxit = 2000;
    for (int k = 1; k < argc; ++k) {
        const char *a = argv[k];
        if (a[0] != '-') continue;                       // "device=<n>" is handled by findCudaDevice
        switch (a[1]) {
        case 'M': matrix_file = a + 2; break;
        case 'V': vector_file = a + 2; break;
        case 'D': debug = true; break;
        case 'P': print = true; break;
        case 'U': unprec = true; break;
        case 'R': p_zero_mat = std::stod(a + 2); break;
        case 'N': dim = std::stoi(a + 2); break;
        case 'T': tol = std::stod(a + 2); break;
        case 'I': maxit = std::stoi(a + 2); break;
        case 'O': out_file = a + 2; break;
        case 'W': mat_out_file = a + 2; break;
        case 'd': break;                                 // -device=<n>
        default: fprintf(stderr, "Unknown switch '-%s'\n", a + 1); return EXIT_FAILURE;
        }
    }
    findCudaDevice(argc, (const char **)argv);

    int n = 0, nnz = 0;
    double *A = nullptr, *b = nullptr;
    int *iA = nullptr, *jA = nullptr;
    if (matrix_file) {
        printf("Using matrix input file [%s]\n", matrix_file);
        int m = 0;
        if (loadMMSparseMatrix(const_cast<char *>(matrix_file), 'd', true, &m, &n, &nnz, &A, &iA, &jA)) {
            fprintf(stderr, "!!!! loadMMSparseMatrix FAILED\n");
            return EXIT_FAILURE;
        }
        if (m != n) { fprintf(stderr, "!!!! square matrix is expected\n"); return EXIT_FAILURE; }
    } else {
        // example.cpp:274-285: diagonal in [1,10], off-diagonal in [1,10] with probability 1-R
        std::vector<double> va; std::vector<int> vi, vj;
        nnz = fill_csr_matrix<Base1>(dim, dim, &va, &vi, &vj, [&](int i, int j) {
            if (i == j) return rand_float(1, 10);
            return rand_float_0_1() >= p_zero_mat ? rand_float(1, 10) : 0.0;
        }, 1e-3);
        n = dim;
        A = (double *)malloc(sizeof(double) * nnz);
        iA = (int *)malloc(sizeof(int) * (n + 1));
        jA = (int *)malloc(sizeof(int) * nnz);
        memcpy(A, va.data(), sizeof(double) * nnz);
        memcpy(iA, vi.data(), sizeof(int) * (n + 1));
        memcpy(jA, vj.data(), sizeof(int) * nnz);
    }
    b = (double *)malloc(sizeof(double) * n);
    if (vector_file) {
        printf("Using vector input file [%s]\n", vector_file);
        int vm = 0, vn = 0, vnnz = 0; double *vA = nullptr; int *vIA = nullptr, *vJA = nullptr;
        if (loadMMSparseMatrix(const_cast<char *>(vector_file), 'd', true, &vm, &vn, &vnnz, &vA, &vIA, &vJA)) {
            fprintf(stderr, "!!!! loadMMSparseMatrix FAILED\n");
            return EXIT_FAILURE;
        }
        if (vn != 1) { fprintf(stderr, "b must be a vector !\n"); return EXIT_FAILURE; }
        if (vm != n) { fprintf(stderr, "incorrect dim\n"); return EXIT_FAILURE; }
        toDenseVector(vm, vnnz, vA, vIA, b);
        free(vA); free(vIA); free(vJA);
    } else {
        gen_rand_vector(n, b, p_zero_vec, 1, 5.0);       // example.cpp:338-340
    }
    if (mat_out_file && cudamat_write_mm(mat_out_file, n, n, nnz, A, iA, jA, 0, "written by cudamat_example")) {
        fprintf(stderr, "%s\n", cudamat_last_error());
        return EXIT_FAILURE;
    }
    double *x = (double *)malloc(sizeof(double) * n);
    std::cout << "nnz=" << nnz << std::endl;
    double dtAlg = 0.0;
    const double t1 = second();
    const bool solved = unprec ? bicgstab(n, nnz, A, iA, jA, b, maxit, tol, debug, x, &dtAlg)
                               : bicgstab_lu_precond(n, nnz, A, iA, jA, b, maxit, tol, debug, x, &dtAlg);
    const double t2 = second();
    if (solved) {
        std::cout << "success" << std::endl;
        if (print) {
            std::ostringstream s;
            dump_vector(s, n, x);
            std::cout << "result:" << std::endl << s.str() << std::endl;
        }
        if (out_file && cudamat_write_mm_vector(out_file, n, x, "solution written by cudamat_example")) {
            fprintf(stderr, "%s\n", cudamat_last_error());
            return EXIT_FAILURE;
        }
        std::cout << "algorithm delta time = " << dtAlg << " s" << std::endl;
        std::cout << "total delta time = " << t2 - t1 << " s" << std::endl;
    } else {
        std::cerr << "method failed" << std::endl;
    }
    free(x); free(b); free(A); free(iA); free(jA);
    return solved ? EXIT_SUCCESS : EXIT_FAILURE;
}
