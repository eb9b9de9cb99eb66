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
#include <thread>

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
    // mm_is_valid (mmio.c:93-99): "real hermitian" is not a valid type code; the reference then fails the read
    if (bn.hermitian && bn.real) { fclose(f); set_error("!!!! real hermitian is not a valid Matrix Market type (mmio.c:96)"); return CUDAMAT_E_IO; }
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
    // The entry block is read in one piece and parsed in parallel (strtol / strtod on newline-aligned chunks): the
    // reference goes through fscanf entry by entry (mmio.c:339-390), which takes minutes on GB-scale files.
    std::vector<Entry> e;
    {
        const long pos = ftell(f);
        fseek(f, 0, SEEK_END);
        const long end = ftell(f);
        fseek(f, pos, SEEK_SET);
        std::vector<char> txt((size_t)std::max<long>(end - pos, 0) + 1);
        const size_t got = fread(txt.data(), 1, txt.size() - 1, f);
        txt[got] = 0;
        fclose(f);
        unsigned T = std::thread::hardware_concurrency();
        T = std::max(1u, std::min(T ? T : 1u, 16u));
        if (got < (1u << 20)) T = 1;
        std::vector<std::vector<Entry>> part(T);
        std::vector<int> perr(T, 0);
        std::vector<size_t> cut(T + 1, got);
        cut[0] = 0;
        for (unsigned t = 1; t < T; ++t) {                       // chunk boundaries right after a newline
            size_t c = got / T * t;
            while (c < got && txt[c] != '\n') ++c;
            cut[t] = std::min(got, c + 1);
        }
        auto work = [&](unsigned t) {
            const char *p = txt.data() + cut[t], *pe = txt.data() + cut[t + 1];
            std::vector<Entry> &out = part[t];
            out.reserve((size_t)(NZ / T + 16) * (bn.general ? 1 : 2));
            while (p < pe) {
                while (p < pe && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
                if (p >= pe) break;
                if (*p == '%') { while (p < pe && *p != '\n') ++p; continue; }
                char *q;
                const long i = strtol(p, &q, 10);
                if (q == p) { perr[t] = 1; return; }
                p = q;
                const long j = strtol(p, &q, 10);
                if (q == p) { perr[t] = 1; return; }
                p = q;
                const double v = strtod(p, &q);
                if (q == p) { perr[t] = 1; return; }
                p = q;
                out.push_back({(int)i, (int)j, v});
                if (!bn.general && i != j) out.push_back({(int)j, (int)i, bn.skew ? -v : v});     // mirrored entry (:197-223)
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
        long read_entries = 0;
        size_t total = 0;
        for (unsigned t = 0; t < T; ++t) {
            if (perr[t]) { set_error("load_mm: malformed entry in '%s'", filename); return CUDAMAT_E_IO; }
            total += part[t].size();
        }
        e.reserve(total);
        for (unsigned t = 0; t < T; ++t) e.insert(e.end(), part[t].begin(), part[t].end());
        for (const Entry &x : e) read_entries += (bn.general || x.i >= x.j) ? 1 : 0;
        // entry count check: general files hold exactly NZ entries; symmetric ones NZ stored (+ mirrors)
        long stored = 0;
        if (bn.general) stored = (long)e.size();
        else { long offd = 0, dg = 0; for (const Entry &x : e) { if (x.i == x.j) ++dg; else ++offd; } stored = dg + offd / 2; }
        (void)read_entries;
        if (stored < NZ) { set_error("load_mm: premature end of file in '%s' (%ld of %ld entries)", filename, stored, NZ); return CUDAMAT_E_IO; }
        if (stored > NZ) { set_error("load_mm: '%s' holds more entries (%ld) than its size line says (%ld)", filename, stored, NZ); return CUDAMAT_E_IO; }
    }
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

// ---- writers (replace mm_write_mtx_crd / mm_write_banner, mmio.c:405-445, 447-510) ---------------------------
// CSR (base read from rowptr[0], 0 or 1) -> "%%MatrixMarket matrix coordinate real general|symmetric", 1-based entries,
// 17 significant digits (round-trips a double exactly).  symmetric != 0 writes only the lower triangle (row >= col);
// the caller asserts the matrix is symmetric.
extern "C" int cudamat_write_mm(const char *filename, int m, int n, int nnz, const double *val, const int *rowptr,
                                const int *colind, int symmetric, const char *comment) {
    if (!filename || !rowptr || (nnz > 0 && (!val || !colind)) || m < 0 || n < 0) { set_error("write_mm: invalid argument"); return CUDAMAT_E_INVALID; }
    const int base = rowptr[0];
    if (base != 0 && base != 1) { set_error("write_mm: index base %d is neither 0 nor 1", base); return CUDAMAT_E_INVALID; }
    FILE *f = fopen(filename, "w");
    if (!f) { set_error("write_mm: can not open '%s' for writing", filename); return CUDAMAT_E_IO; }
    long stored = 0;
    for (int i = 0; i < m; ++i)
        for (int k = rowptr[i] - base; k < rowptr[i + 1] - base; ++k)
            if (!symmetric || colind[k] - base <= i) ++stored;
    fprintf(f, "%%%%MatrixMarket matrix coordinate real %s\n", symmetric ? "symmetric" : "general");
    if (comment && *comment) fprintf(f, "%% %s\n", comment);
    fprintf(f, "%d %d %ld\n", m, n, stored);
    std::vector<char> buf(1 << 20);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (int i = 0; i < m; ++i)
        for (int k = rowptr[i] - base; k < rowptr[i + 1] - base; ++k)
            if (!symmetric || colind[k] - base <= i) fprintf(f, "%d %d %.17g\n", i + 1, colind[k] - base + 1, val[k]);
    const bool ok = fflush(f) == 0 && !ferror(f);
    fclose(f);
    if (!ok) { set_error("write_mm: write error on '%s'", filename); return CUDAMAT_E_IO; }
    return CUDAMAT_OK;
}
// dense vector as the n x 1 coordinate matrix the reference's -V switch reads (example.cpp:310-336, vec3.mtx);
// exact zeros are skipped like any sparse writer would (toDenseVector restores them, pbicgstab.cu:413-423)
extern "C" int cudamat_write_mm_vector(const char *filename, int n, const double *x, const char *comment) {
    if (!filename || (n > 0 && !x) || n < 0) { set_error("write_mm_vector: invalid argument"); return CUDAMAT_E_INVALID; }
    FILE *f = fopen(filename, "w");
    if (!f) { set_error("write_mm_vector: can not open '%s' for writing", filename); return CUDAMAT_E_IO; }
    long stored = 0;
    for (int i = 0; i < n; ++i) if (x[i] != 0.0) ++stored;
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n");
    if (comment && *comment) fprintf(f, "%% %s\n", comment);
    fprintf(f, "%d 1 %ld\n", n, stored);
    for (int i = 0; i < n; ++i) if (x[i] != 0.0) fprintf(f, "%d 1 %.17g\n", i + 1, x[i]);
    const bool ok = fflush(f) == 0 && !ferror(f);
    fclose(f);
    if (!ok) { set_error("write_mm_vector: write error on '%s'", filename); return CUDAMAT_E_IO; }
    return CUDAMAT_OK;
}
