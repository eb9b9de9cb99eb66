// rowclass.cu — row-class dictionaries for the compressed SpMV variants (analysis side).
//
// Stencil matrices (the reference's mat900 / mat10000 fixtures, the N^3 Poisson systems of BASELINE.json) have
// only a handful of distinct rows up to translation: the column OFFSETS ja[k] - i of a row, and often its VALUES
// too, come from a tiny set.  cudamat_analyze detects this on the device and replaces the 4-byte column index
// (and, when the values repeat as well, the 8-byte value) of every entry by ONE byte per row:
//   * CUDAMAT_SPMV_PATTERN: cls[i] -> (len, offsets[len]); values still streamed from CSR;
//   * CUDAMAT_SPMV_CLASS:   cls[i] -> (len, offsets[len], values[len]); the CSR arrays are not read at all.
// The dictionary (<= kDictMax classes of <= kDictLen entries) is exact: every row is verified against its class,
// a single mismatch (or a hash collision) disables the variant and the CSR kernel is used.  Entry order is the
// row's storage order, so the row sums are bit-identical to the CSR kernels.
// This is what cusparseDcsrmv (pbicgstab.cu:67,104,132,646,676,704) cannot do: it must stream 12 B per entry.
#include "solver.h"
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <memory>
#include <vector>

namespace cudamat {

constexpr int kTab = 2048;                      // open-addressing table slots (power of two)

__device__ __forceinline__ uint64_t mixk(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ULL;
    return h ^ (h >> 29);
}
// key of a row: hash of (len, offsets[, value bits]); never 0
__device__ __forceinline__ uint64_t row_key(int row, const int *ia, const int *ja, const double *a, bool with_vals, int *len_out) {
    const int s = ia[row], e = ia[row + 1];
    *len_out = e - s;
    uint64_t h = mixk(0x243F6A8885A308D3ULL, (uint64_t)(e - s));
    for (int k = s; k < e; ++k) {
        h = mixk(h, (uint64_t)(uint32_t)(ja[k] - row));
        if (with_vals) h = mixk(h, (uint64_t)__double_as_longlong(a[k]));
    }
    return h ? h : 1ull;
}
__device__ __forceinline__ int tab_find(const unsigned long long *tab, uint64_t key) {
    int slot = (int)(key & (kTab - 1));
    for (int probes = 0; probes < kTab; ++probes) {
        const unsigned long long cur = tab[slot];
        if (cur == key) return slot;
        if (cur == 0ull) return -1;
        slot = (slot + 1) & (kTab - 1);
    }
    return -1;
}

// pass 1: insert every row's key, remember the smallest row of each key. fail[0] != 0: not representable
__global__ void k_cls_insert(int n, const int *ia, const int *ja, const double *a, int with_vals,
                             unsigned long long *tab, int *rep, int *fail) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int len;
    const uint64_t key = row_key(row, ia, ja, a, with_vals != 0, &len);
    if (len > kDictLen) { *fail = 1; return; }
    int slot = (int)(key & (kTab - 1));
    for (int probes = 0;; ++probes) {
        if (probes >= kTab) { *fail = 1; return; }
        unsigned long long cur = *(volatile unsigned long long *)(tab + slot);
        if (cur == 0ull) cur = atomicCAS(tab + slot, 0ull, (unsigned long long)key);
        if (cur == 0ull || cur == key) break;
        slot = (slot + 1) & (kTab - 1);
    }
    if (row < *(volatile int *)(rep + slot)) atomicMin(rep + slot, row);
}

// pass 2 (one thread): number the classes by their first row, build the dictionary from the representatives
__global__ void k_cls_number(const int *ia, const int *ja, const double *a, int with_vals,
                             const unsigned long long *tab, const int *rep, int *slot_id, RowDict *dict, int *ncls_out, int *fail) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int slots[kDictMax], m = 0;
    for (int sl = 0; sl < kTab; ++sl) slot_id[sl] = -1;
    *ncls_out = 0;
    if (*fail) return;                                       // pass 1 already gave up (too many classes / long rows)
    for (int sl = 0; sl < kTab; ++sl) {
        if (tab[sl] != 0ull) {
            if (m >= kDictMax) { *fail = 1; *ncls_out = 0; return; }
            int pos = m++;                                   // insertion sort by representative row
            while (pos > 0 && rep[slots[pos - 1]] > rep[sl]) { slots[pos] = slots[pos - 1]; --pos; }
            slots[pos] = sl;
        }
    }
    for (int id = 0; id < m; ++id) {
        const int sl = slots[id], row = rep[sl];
        slot_id[sl] = id;
        const int s = ia[row], len = ia[row + 1] - s;
        dict->len[id] = len;
        for (int k = 0; k < kDictLen; ++k) {
            dict->off[id * kDictLen + k] = (k < len) ? ja[s + k] - row : 0;
            dict->val[id * kDictLen + k] = (k < len && with_vals) ? a[s + k] : 0.0;
        }
    }
    *ncls_out = m;
}

// pass 3: class id of every row, verified entry by entry against the dictionary
__global__ void k_cls_assign(int n, const int *ia, const int *ja, const double *a, int with_vals,
                             const unsigned long long *tab, const int *slot_id, const RowDict *dict, unsigned char *cls, int *fail) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n || *fail) return;                           // an earlier pass gave up: the dictionary is not valid
    int len;
    const uint64_t key = row_key(row, ia, ja, a, with_vals != 0, &len);
    const int slot = tab_find(tab, key);
    const int id = slot >= 0 ? slot_id[slot] : -1;
    if (id < 0 || dict->len[id] != len) { *fail = 1; return; }
    const int s = ia[row];
    for (int k = 0; k < len; ++k) {
        if (dict->off[id * kDictLen + k] != ja[s + k] - row) { *fail = 1; return; }
        if (with_vals && __double_as_longlong(dict->val[id * kDictLen + k]) != __double_as_longlong(a[s + k])) { *fail = 1; return; }
    }
    cls[row] = (unsigned char)id;
}

__global__ void k_shift_cols(int64_t nnz, const int *ja_global, int row0, int *out) {     // global column id -> offset-preserving local frame
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) out[k] = ja_global[k] - row0;
}

__global__ void k_cls_hist(int n, const unsigned char *cls, unsigned *hist) {
    __shared__ unsigned h[kDictMax];
    for (int i = threadIdx.x; i < kDictMax; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&h[cls[i] & (kDictMax - 1)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < kDictMax; i += blockDim.x) if (h[i]) atomicAdd(hist + i, h[i]);
}
// tile_ok[t] = every row of tile t belongs to a class whose offsets all fall into the staged windows
__global__ void k_tile_ok(int n, const unsigned char *cls, unsigned long long ok_mask, unsigned char *tile_ok) {
    const int tile = blockIdx.x;
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int r = tile * kTile + threadIdx.x; r < min(n, (tile + 1) * kTile); r += blockDim.x)
        if (!((ok_mask >> (cls[r] & 63)) & 1ull)) bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) tile_ok[tile] = bad ? 0 : 1;
}

// presence mask of every row w.r.t. the superset pattern: tmask[r] = cmask[cls[r]]
struct ClassMasks { unsigned char m[kDictMax]; };
__global__ void k_cls_to_mask(int n, const unsigned char *cls, const ClassMasks cm, unsigned char *tmask) {
    __shared__ unsigned char tab[kDictMax];                        // per-thread indexing of a parameter array is slow
    if (threadIdx.x < kDictMax) tab[threadIdx.x] = cm.m[threadIdx.x];
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) tmask[r] = tab[cls[r] & 63];
}

// ---- TILED plan, host part (pure host code, also reachable through cudamat_tiled_plan_host for the CPU tests) -------
// From the class dictionary and the class histogram: the windows (clusters of the column offsets of the frequent classes),
// the shared-memory position of every entry, which classes fit the windows, the shared-memory records and — when it
// exists — the superset pattern with the presence mask of every class.
struct TiledPlanHost {
    std::unique_ptr<TiledDict> T;
    std::vector<TiledSmemClass> sd;
    ClassMasks cm;
    unsigned long long ok_mask = 0;
    size_t smem_windows = 0;
};
static bool tiled_plan_host(const DictParam &D, int ncls, const unsigned *hist, int n, bool with_vals, TiledPlanHost &P) {
    if (ncls <= 0 || ncls > kDictMax) return false;
    // offsets of the frequent classes (>= 1/64 of the rows, or the most frequent one)
    int top = 0;
    for (int c = 1; c < ncls; ++c) if (hist[c] > hist[top]) top = c;
    std::vector<int> offs;
    for (int c = 0; c < ncls; ++c)
        if (c == top || hist[c] >= (unsigned)std::max(1, n / 64))
            for (int k = 0; k < D.len[c]; ++k) offs.push_back(D.off[c * kDictLen + k]);
    if (offs.empty()) return false;
    std::sort(offs.begin(), offs.end());
    offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
    P.T.reset(new TiledDict());
    TiledDict *T = P.T.get();
    memset(T, 0, sizeof(TiledDict));
    int nseg = 0, base = 0;
    size_t k = 0;
    bool fits = true;
    while (k < offs.size()) {
        int lo = offs[k], hi = offs[k];
        while (k + 1 < offs.size() && (long long)offs[k + 1] - hi <= 4096) hi = offs[++k];
        ++k;
        if (nseg == kMaxSeg) { fits = false; break; }
        lo &= ~1;                                              // 16-byte aligned bulk copies (tile starts are even)
        int len = kTile + (hi - lo) + 1;
        len = (len + 1) & ~1;
        T->seg_lo[nseg] = lo; T->seg_len[nseg] = len; T->seg_base[nseg] = base;
        base += len;
        ++nseg;
    }
    P.smem_windows = sizeof(double) * (size_t)base;
    if (!fits || P.smem_windows > 100 * 1024) { P.T.reset(); return false; }
    T->nseg = nseg;
    P.ok_mask = 0;
    for (int c = 0; c < ncls; ++c) {
        T->len[c] = D.len[c];
        bool ok = true;
        for (int q = 0; q < kDictLen; ++q) {
            const int o = D.off[c * kDictLen + q];
            T->off[c * kDictLen + q] = o;
            T->val[c * kDictLen + q] = D.val[c * kDictLen + q];
            int disp = 0;
            if (q < D.len[c]) {
                int sg = -1;
                for (int g = 0; g < nseg; ++g)
                    if (o >= T->seg_lo[g] && o + kTile <= T->seg_lo[g] + T->seg_len[g]) { sg = g; break; }
                if (sg < 0) ok = false; else disp = T->seg_base[sg] + (o - T->seg_lo[sg]);
            }
            T->disp[c * kDictLen + q] = disp;
        }
        if (ok) P.ok_mask |= 1ull << c;
    }
    // shared-memory form of the dictionary (one TMA bulk copy per CTA, behind the windows)
    T->sdict_base = base;
    T->maxlen = 0; T->minlen = kDictLen;
    T->disp0 = -1;
    for (int g = 0; g < nseg; ++g)
        if (0 >= T->seg_lo[g] && kTile <= T->seg_lo[g] + T->seg_len[g]) { T->disp0 = T->seg_base[g] - T->seg_lo[g]; break; }
    P.sd.assign((size_t)ncls, TiledSmemClass());
    for (int c = 0; c < ncls; ++c) {
        memset(&P.sd[c], 0, sizeof(TiledSmemClass));
        T->maxlen = std::max(T->maxlen, T->len[c]);
        T->minlen = std::min(T->minlen, T->len[c]);
        for (int q = 0; q < kDictLen; ++q) {
            P.sd[c].boff[q] = q < T->len[c] ? T->disp[c * kDictLen + q] * 8 : -1;
            P.sd[c].val[q] = q < T->len[c] ? T->val[c * kDictLen + q] : 0.0;
        }
    }
    // superset pattern over the staged classes: sorted union of their offsets (every class ascending, so each is an
    // order-preserving subset); for the values dictionary the value at an offset must be the same in every class
    memset(&P.cm, 0, sizeof P.cm);
    std::vector<int> so; std::vector<double> sv;
    bool ok = true;
    for (int c = 0; c < ncls && ok; ++c) {
        if (!((P.ok_mask >> c) & 1ull)) continue;
        for (int q = 0; q < T->len[c] && ok; ++q) {
            const int o = T->off[c * kDictLen + q];
            if (q > 0 && o <= T->off[c * kDictLen + q - 1]) ok = false;
            const double v = T->val[c * kDictLen + q];
            size_t j = 0;
            while (j < so.size() && so[j] < o) ++j;
            if (j < so.size() && so[j] == o) { if (with_vals && memcmp(&sv[j], &v, sizeof v) != 0) ok = false; }
            else { so.insert(so.begin() + j, o); sv.insert(sv.begin() + j, v); }
        }
    }
    if (ok && !so.empty() && so.size() <= 8) {
        for (int c = 0; c < ncls; ++c) {
            if (!((P.ok_mask >> c) & 1ull)) continue;
            for (int q = 0; q < T->len[c]; ++q)
                for (size_t j = 0; j < so.size(); ++j)
                    if (so[j] == T->off[c * kDictLen + q]) { P.cm.m[c] |= (unsigned char)(1u << j); T->sup_boff[j] = T->disp[c * kDictLen + q] * 8; }
        }
        for (size_t j = 0; j < so.size(); ++j) { T->sup_val[j] = sv[j]; T->sup_off[j] = so[j]; }
        T->sup_len = (int)so.size();
    }
    return true;
}

// TILED plan: class histogram on the device, the host plan above, then the device-side tables (shared-memory records,
// presence mask per row, eligibility byte per tile)
static int tiled_plan(cudamat_solver *s, RowClasses &C, bool with_vals) {
    const int n = s->n;
    unsigned *d_hist = nullptr;
    CM_CUDA(dev_alloc((void **)&d_hist, sizeof(unsigned) * kDictMax));
    CM_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned) * kDictMax, s->stream));
    k_cls_hist<<<296, 256, 0, s->stream>>>(n, C.d_cls, d_hist);
    unsigned hist[kDictMax];
    CM_CUDA(cudaMemcpyAsync(hist, d_hist, sizeof hist, cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    dev_free(d_hist);
    s->launches++;
    TiledPlanHost P;
    if (!tiled_plan_host(*C.h_dict, C.ncls, hist, n, with_vals, P)) return CUDAMAT_OK;
    CM_CUDA(dev_alloc((void **)&C.d_sdict, sizeof(TiledSmemClass) * (size_t)C.ncls));
    CM_CUDA(cudaMemcpyAsync(C.d_sdict, P.sd.data(), sizeof(TiledSmemClass) * (size_t)C.ncls, cudaMemcpyHostToDevice, s->stream));
    if (P.T->sup_len > 0) {
        CM_CUDA(dev_alloc((void **)&C.d_tmask, (size_t)n + 16));
        k_cls_to_mask<<<(n + 255) / 256, 256, 0, s->stream>>>(n, C.d_cls, P.cm, C.d_tmask);
        CM_CUDA(cudaGetLastError());
        s->launches++;
    }
    const int ntile = (n + kTile - 1) / kTile;
    CM_CUDA(dev_alloc((void **)&C.d_tile_ok, (size_t)std::max(ntile, 1)));
    k_tile_ok<<<ntile, 256, 0, s->stream>>>(n, C.d_cls, P.ok_mask, C.d_tile_ok);
    CM_CUDA(cudaGetLastError());
    CM_CUDA(cudaStreamSynchronize(s->stream));                       // P.sd is a local
    s->launches++;
    C.h_tdict = P.T.release();
    C.tiled_smem = P.smem_windows + sizeof(TiledSmemClass) * (size_t)C.ncls;
    return CUDAMAT_OK;
}

void rowclass_release(cudamat_solver *s) {
    delete s->march; s->march = nullptr; s->march_tmask = nullptr;
    for (RowClasses *C : {&s->cls_g}) {
        dev_free(C->d_cls); delete C->h_dict; delete C->h_tdict; dev_free(C->d_tile_ok); dev_free(C->d_sdict); dev_free(C->d_tmask); dev_free(C->d_dict);
        *C = RowClasses();
    }
    for (int m = 0; m < 2; ++m) {
        dev_free(s->cls[m].d_cls);
        delete s->cls[m].h_dict;
        delete s->cls[m].h_tdict;
        dev_free(s->cls[m].d_tile_ok);
        dev_free(s->cls[m].d_sdict);
        dev_free(s->cls[m].d_tmask);
        dev_free(s->cls[m].d_dict);
        s->cls[m] = RowClasses();
    }
}

// one classification pass (m = 1: offsets + values, m = 0: offsets only) over the column ids `ja` into C
static int classify_rows(cudamat_solver *s, const int *ja, int m, RowClasses &C, unsigned long long *tab, int *rep, int *slot_id, int *flags) {
    const int n = s->n;
    cudaError_t e;
    if ((e = dev_alloc((void **)&C.d_cls, (size_t)n + 16)) != cudaSuccess || (e = dev_alloc((void **)&C.d_dict, sizeof(RowDict))) != cudaSuccess) {
        cuda_ok(e, "cudaMalloc(row classes)", __FILE__, __LINE__); return CUDAMAT_E_CUDA;
    }
    cudaMemsetAsync(tab, 0, sizeof(unsigned long long) * kTab, s->stream);
    cudaMemsetAsync(rep, 0x7f, sizeof(int) * kTab, s->stream);
    cudaMemsetAsync(flags, 0, sizeof(int) * 2, s->stream);
    const int grid = (n + 255) / 256;
    k_cls_insert<<<grid, 256, 0, s->stream>>>(n, s->d_ia, ja, s->d_a, m, tab, rep, flags);
    k_cls_number<<<1, 32, 0, s->stream>>>(s->d_ia, ja, s->d_a, m, tab, rep, slot_id, C.d_dict, flags + 1, flags);
    k_cls_assign<<<grid, 256, 0, s->stream>>>(n, s->d_ia, ja, s->d_a, m, tab, slot_id, C.d_dict, C.d_cls, flags);
    s->launches += 3;
    int h[2] = {1, 0};
    if ((e = cudaMemcpyAsync(h, flags, sizeof h, cudaMemcpyDeviceToHost, s->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s->stream)) != cudaSuccess) {
        cuda_ok(e, "row class analysis", __FILE__, __LINE__); return CUDAMAT_E_CUDA;
    }
    if (h[0] == 0 && h[1] > 0) {
        RowDict hd;
        if ((e = cudaMemcpy(&hd, C.d_dict, sizeof hd, cudaMemcpyDeviceToHost)) != cudaSuccess) {
            cuda_ok(e, "row class dictionary download", __FILE__, __LINE__); return CUDAMAT_E_CUDA;
        }
        C.ncls = h[1];
        C.h_dict = new DictParam();
        memset(C.h_dict, 0, sizeof(DictParam));
        for (int id = 0; id < C.ncls; ++id) {
            C.h_dict->len[id] = hd.len[id];
            C.h_dict->run[id] = -1;
            for (int k = 0; k < kDictLen; ++k) {
                C.h_dict->off[id * kDictLen + k] = hd.off[id * kDictLen + k];
                C.h_dict->val[id * kDictLen + k] = hd.val[id * kDictLen + k];
            }
            for (int k = 0; k + 2 < hd.len[id]; ++k)
                if (hd.off[id * kDictLen + k] == -1 && hd.off[id * kDictLen + k + 1] == 0 && hd.off[id * kDictLen + k + 2] == 1) { C.h_dict->run[id] = k; break; }
        }
    } else { dev_free(C.d_cls); dev_free(C.d_dict); C = RowClasses(); }
    return CUDAMAT_OK;
}

// builds s->cls[0] (offsets only) and s->cls[1] (offsets + values) when the matrix allows it
int rowclass_analyze(cudamat_solver *s) {
    rowclass_release(s);
    const int n = s->n;
    if (n <= 0 || s->nnz <= 0 || s->max_row_len > kDictLen) return CUDAMAT_OK;
    unsigned long long *tab = nullptr; int *rep = nullptr, *slot_id = nullptr, *flags = nullptr;
    CM_CUDA(dev_alloc((void **)&tab, sizeof(unsigned long long) * kTab));
    CM_CUDA(dev_alloc((void **)&rep, sizeof(int) * kTab));
    CM_CUDA(dev_alloc((void **)&slot_id, sizeof(int) * kTab));
    CM_CUDA(dev_alloc((void **)&flags, sizeof(int) * 2));
    int rc = CUDAMAT_OK;
    for (int m = 1; m >= 0 && rc == CUDAMAT_OK; --m) rc = classify_rows(s, s->d_ja, m, s->cls[m], tab, rep, slot_id, flags);   // m = 1: with values, m = 0: offsets only
    cudaStreamSynchronize(s->stream);
    dev_free(tab); dev_free(rep); dev_free(slot_id); dev_free(flags);
    for (int m = 1; m >= 0 && rc == CUDAMAT_OK; --m)
        if (s->cls[m].ncls > 0) rc = tiled_plan(s, s->cls[m], m == 1);
    // MARCH: needs the values dictionary with a superset pattern and EVERY tile inside the windows.  On a sharded handle the
    // local column ids of the boundary planes point into the halo region, which breaks the translation invariance the
    // classes rest on: there the classes are formed a second time over GLOBAL offsets (ja_global - row0), and the kernel
    // fetches the planes -1 / P from the halo region instead — if that region is exactly the two neighbours' planes.
    RowClasses *MC = &s->cls[1];
    if (rc == CUDAMAT_OK && s->comm && s->d_ja_global && s->cls[1].ncls > 0) {
        int *ja_g = nullptr;
        CM_CUDA(dev_alloc((void **)&ja_g, sizeof(int) * (size_t)std::max<int64_t>(s->nnz, 1)));
        k_shift_cols<<<1184, 256, 0, s->stream>>>(s->nnz, s->d_ja_global, (int)s->row0, ja_g);
        s->launches++;
        CM_CUDA(dev_alloc((void **)&tab, sizeof(unsigned long long) * kTab));
        CM_CUDA(dev_alloc((void **)&rep, sizeof(int) * kTab));
        CM_CUDA(dev_alloc((void **)&slot_id, sizeof(int) * kTab));
        CM_CUDA(dev_alloc((void **)&flags, sizeof(int) * 2));
        rc = classify_rows(s, ja_g, 1, s->cls_g, tab, rep, slot_id, flags);
        cudaStreamSynchronize(s->stream);
        dev_free(tab); dev_free(rep); dev_free(slot_id); dev_free(flags); dev_free(ja_g);
        if (rc == CUDAMAT_OK && s->cls_g.ncls > 0) rc = tiled_plan(s, s->cls_g, true);
        MC = &s->cls_g;
    }
    if (rc == CUDAMAT_OK && MC->h_tdict && MC->h_tdict->sup_len > 0 && MC->d_tmask) {
        const int ntile = (n + kTile - 1) / kTile;
        std::vector<unsigned char> ok((size_t)ntile);
        cudaError_t e = cudaMemcpyAsync(ok.data(), MC->d_tile_ok, (size_t)ntile, cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) { cuda_ok(e, "tile_ok download", __FILE__, __LINE__); return CUDAMAT_E_CUDA; }
        bool all_ok = true;
        for (unsigned char b : ok) if (!b) { all_ok = false; break; }
        MarchPlan M;
        if (all_ok && march_plan_host(*MC->h_tdict, n, M)) {
            M.lo_base = M.hi_base = -1; M.n_tot = n;
            bool usable = true;
            if (s->comm) {
                static const int env_ms = [] { const char *e = getenv("CUDAMAT_MARCH_SHARDS"); return e && *e ? atoi(e) : -1; }();   // 0 off, 2 force
                const int ms = env_ms >= 0 ? env_ms : s->opt_march_shards;
                usable = comm_halo_planes(s, M.D, &M.lo_base, &M.hi_base) && ms != 0;
                M.n_tot = n + s->nhalo;
                // MARCH beats TILED on a shard when its work items fill the CTA slots — 256^3 on 2 GPUs: 288 items for 296 slots,
                // 3415 vs 3300 it/s; 512^3 on 2 GPUs: 256 items, 511 vs 520 it/s (same box, A/B)
                int dev = 0, sms = 148;
                if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                const int G = 2 * sms;
                if ((long long)march_choose_zc(M.S, M.P, G) * M.S * 20 < (long long)G * 19 && ms < 2) usable = false;
            }
            if (usable) {
                s->march = new MarchPlan(M);
                s->march_tmask = MC->d_tmask;
                int dev = 0, sms = 148;
                if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                s->march_grid = 2 * sms;
            }
        }
    }
    return rc;
}

}  // namespace cudamat

// host planner of the TILED SpMV variant, exported for the CPU tests (no device needed)
extern "C" int cudamat_tiled_plan_host(int ncls, const int *len, const int *off, const double *val, const unsigned *hist, int n,
                                       int with_vals, int *nseg, int *seg_lo, int *seg_len, int *seg_base, int *disp,
                                       unsigned long long *ok_mask, int *sup_len, int *sup_boff, double *sup_val,
                                       unsigned char *class_mask, long long *smem_bytes) {
    using namespace cudamat;
    if (ncls <= 0 || ncls > kDictMax || !len || !off || !hist || n <= 0) { set_error("tiled_plan_host: invalid argument"); return CUDAMAT_E_INVALID; }
    std::unique_ptr<DictParam> D(new DictParam());
    memset(D.get(), 0, sizeof(DictParam));
    for (int c = 0; c < ncls; ++c) {
        if (len[c] < 0 || len[c] > kDictLen) { set_error("tiled_plan_host: class length %d out of range", len[c]); return CUDAMAT_E_INVALID; }
        D->len[c] = len[c];
        for (int q = 0; q < kDictLen; ++q) { D->off[c * kDictLen + q] = off[c * kDictLen + q]; D->val[c * kDictLen + q] = val ? val[c * kDictLen + q] : 0.0; }
    }
    TiledPlanHost P;
    const bool have = tiled_plan_host(*D, ncls, hist, n, with_vals != 0, P);
    if (nseg) *nseg = have ? P.T->nseg : 0;
    if (!have) return CUDAMAT_OK;
    for (int g = 0; g < kMaxSeg; ++g) {
        if (seg_lo) seg_lo[g] = P.T->seg_lo[g];
        if (seg_len) seg_len[g] = P.T->seg_len[g];
        if (seg_base) seg_base[g] = P.T->seg_base[g];
    }
    if (disp) for (int k = 0; k < ncls * kDictLen; ++k) disp[k] = P.T->disp[k];
    if (ok_mask) *ok_mask = P.ok_mask;
    if (sup_len) *sup_len = P.T->sup_len;
    for (int q = 0; q < 8; ++q) { if (sup_boff) sup_boff[q] = P.T->sup_boff[q]; if (sup_val) sup_val[q] = P.T->sup_val[q]; }
    if (class_mask) for (int c = 0; c < kDictMax; ++c) class_mask[c] = P.cm.m[c];
    if (smem_bytes) *smem_bytes = (long long)(P.smem_windows + sizeof(TiledSmemClass) * (size_t)ncls);
    return CUDAMAT_OK;
}

