// mmload.cpp — Matrix Market coordinate loader with the semantics of the reference's
// loadMMSparseMatrix (mmio_wrapper.h:133-348 on top of NIST mmio.c): real/integer entries,
// general/symmetric/skew-symmetric/hermitian storage (mirrored into the full pattern), sorted
// CSR (or CSC) output, index base auto-detected exactly like the reference (:266-289: any index 0 =>
// base-0, any row == m or col == n => base-1, both => error, neither => base-0), the same
// verify_pattern checks (:91-130), malloc()ed output arrays.  Host-only code, written from the
// format specification; shares no source with mmio.c.
#include "../../include/cudamat_b200.h"
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace cudamat { void set_error(const char *fmt, ...); }
using cudamat::set_error;

namespace {

struct Banner { bool coordinate, real, integer, complex_, pattern, general, symmetric, skew, hermitian; };

static std::string lower(std::string s) { for (auto &c : s) c = (char)tolower((unsigned char)c); return s; }

static bool parse_banner(const char *line, Banner &b) {
    char w[5][64];
    if (sscanf(line, "%63s %63s %63s %63s %63s", w[0], w[1], w[2], w[3], w[4]) != 5) return false;
    if (strcmp(w[0], "%%MatrixMarket") != 0) return false;
    if (lower(w[1]) != "matrix") return false;
    const std::string fmt = lower(w[2]), field = lower(w[3]), sym = lower(w[4]);
    b = Banner{};
    b.coordinate = fmt == "coordinate";
    if (!b.coordinate && fmt != "array") return false;
    b.real = field == "real"; b.integer = field == "integer"; b.complex_ = field == "complex"; b.pattern = field == "pattern";
    if (!(b.real || b.integer || b.complex_ || b.pattern)) return false;
    b.general = sym == "general"; b.symmetric = sym == "symmetric"; b.skew = sym == "skew-symmetric"; b.hermitian = sym == "hermitian";
    return b.general || b.symmetric || b.skew || b.hermitian;
}

struct Entry { int i, j; double v; };

}  // namespace

extern "C" int cudamat_load_mm(const char *filename, int csr_format, int *m, int *n, int *nnz,
                               double **aVal, int **aRowInd, int **aColInd) {
    if (!filename || !m || !n || !nnz || !aVal || !aRowInd || !aColInd) { set_error("load_mm: null argument"); return CUDAMAT_E_INVALID; }
    FILE *f = fopen(filename, "r");
    if (!f) { set_error("!!!! can not open file: '%s'", filename); return CUDAMAT_E_IO; }
    std::vector<char> buf(1 << 16);
    Banner bn{};
    if (!fgets(buf.data(), (int)buf.size(), f) || !parse_banner(buf.data(), bn)) {
        fclose(f); set_error("load_mm: '%s' has no valid MatrixMarket banner", filename); return CUDAMAT_E_IO;
    }
    if (bn.complex_) { fclose(f); set_error("!!!! complex matrix requires type 'z' or 'c'"); return CUDAMAT_E_IO; }
    if (!bn.coordinate || bn.pattern) { fclose(f); set_error("!!!! dense, array, pattern and integer matrices are not supported"); return CUDAMAT_E_IO; }
    // size line: first non-comment, non-blank line
    long M = 0, N = 0, NZ = 0;
    bool have_size = false;
    while (fgets(buf.data(), (int)buf.size(), f)) {
        const char *p = buf.data();
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '%' || *p == '\n' || *p == '\r' || *p == 0) continue;
        if (sscanf(p, "%ld %ld %ld", &M, &N, &NZ) == 3) have_size = true;
        break;
    }
    if (!have_size || M < 0 || N < 0 || NZ < 0 || M > 0x7fffffffL || N > 0x7fffffffL || NZ > 0x3fffffffL) {
        fclose(f); set_error("load_mm: '%s' has no valid size line", filename); return CUDAMAT_E_IO;
    }
    std::vector<Entry> e;
    e.reserve((size_t)NZ * ((bn.general) ? 1 : 2));
    for (long k = 0; k < NZ; ++k) {
        int i, j; double v;
        if (fscanf(f, "%d %d %lg", &i, &j, &v) != 3) { fclose(f); set_error("load_mm: premature end of file in '%s' (entry %ld)", filename, k); return CUDAMAT_E_IO; }
        e.push_back({i, j, v});
        if (!bn.general && i != j) e.push_back({j, i, bn.skew ? -v : v});     // mirrored entry (:197-223)
    }
    fclose(f);
    const long nz = (long)e.size();
    // sort by the major index, then the minor one (:253-258)
    if (csr_format) std::stable_sort(e.begin(), e.end(), [](const Entry &a, const Entry &b) { return a.i != b.i ? a.i < b.i : a.j < b.j; });
    else            std::stable_sort(e.begin(), e.end(), [](const Entry &a, const Entry &b) { return a.j != b.j ? a.j < b.j : a.i < b.i; });
    bool base0 = false, base1 = false;
    for (const Entry &x : e) {
        if (x.i == 0 || x.j == 0) base0 = true;
        if (x.i == M || x.j == N) base1 = true;
    }
    if (base0 && base1) { set_error("Error: input matrix is base-0 and base-1"); return CUDAMAT_E_IO; }
    const int base = base1 ? 1 : 0;
    const long major = csr_format ? M : N;
    int *ptr = (int *)malloc(sizeof(int) * (size_t)(major + 1));
    int *ind = (int *)malloc(sizeof(int) * (size_t)std::max<long>(nz, 1));
    double *val = (double *)malloc(sizeof(double) * (size_t)std::max<long>(nz, 1));
    if (!ptr || !ind || !val) { free(ptr); free(ind); free(val); set_error("!!!! allocation error, malloc failed"); return CUDAMAT_E_IO; }
    bool bad = false;
    for (long k = 0; k <= major; ++k) ptr[k] = 0;
    ptr[0] = base;
    for (const Entry &x : e) {
        const long r = (csr_format ? x.i : x.j) - base;
        if (r < 0 || r >= major) { bad = true; break; }
        ptr[r + 1]++;
    }
    if (!bad) {
        for (long k = 0; k < major; ++k) ptr[k + 1] += ptr[k];
        for (long k = 0; k < nz; ++k) { ind[k] = csr_format ? e[k].j : e[k].i; val[k] = e[k].v; }
        // verify_pattern (:91-130)
        if (nz != ptr[major] - ptr[0]) bad = true;
        for (long r = 0; !bad && r < major; ++r) {
            const int s = ptr[r] - base, t = ptr[r + 1] - base;
            if (s > t) bad = true;
            for (int c = s; !bad && c < t; ++c) {
                if (ind[c] < base) bad = true;
                if (c < t - 1 && ind[c] >= ind[c + 1]) bad = true;          // duplicates / unsorted
            }
        }
    }
    if (bad) { free(ptr); free(ind); free(val); set_error("!!!! verify_pattern failed"); return CUDAMAT_E_IO; }
    *m = (int)M; *n = (int)N; *nnz = (int)nz; *aVal = val;
    if (csr_format) { *aRowInd = ptr; *aColInd = ind; } else { *aColInd = ptr; *aRowInd = ind; }
    return CUDAMAT_OK;
}
