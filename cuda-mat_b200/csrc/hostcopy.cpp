// hostcopy.cpp — host <-> device copies of the one-shot entry points for PAGEABLE host arrays.
//
// The reference's callers hand over malloc'ed arrays (example.cpp:96-104,252).  cudaMemcpyAsync from pageable memory is
// staged by the driver on one thread: 10.7 GB/s measured on the B200 box against 53 GB/s from pinned memory, i.e. 0.14 s
// of a 0.38 s 256^3 solve.  Here a few worker threads copy 4 MB chunks into a recycled pinned staging area and issue the
// DMA of each chunk themselves (two slots per thread, so a thread fills one slot while the other is on the wire).
// Pinned / registered / managed sources and small copies go straight to cudaMemcpyAsync.
#include "solver.h"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace cudamat {

namespace {
constexpr size_t kChunk = 4u << 20;
constexpr int kSlots = 2;
constexpr int kMaxThreads = 16;
constexpr size_t kStagedMin = 8u << 20;      // below this the plain call is as fast
constexpr int kMaxDevStage = 64;

struct Stager {
    std::mutex mu;                            // one staged copy at a time per device
    bool tried = false, ok = false;
    int nthreads = 0;
    char *pin = nullptr;                      // nthreads * kSlots * kChunk, portable pinned
    cudaStream_t st[kMaxThreads] = {};
    cudaEvent_t ev[kMaxThreads][kSlots] = {};
    cudaEvent_t fence = nullptr;
};
Stager g_stagers[kMaxDevStage];

int stage_threads() {
    const char *e = getenv("CUDAMAT_COPY_THREADS");       // 0 = never stage
    if (e && *e) return std::max(0, std::min(kMaxThreads, atoi(e)));
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(8u, hc / 2));
}

bool stager_init(Stager &S) {
    if (S.tried) return S.ok;
    S.tried = true;
    S.nthreads = stage_threads();
    if (S.nthreads <= 0) return false;
    if (cudaHostAlloc((void **)&S.pin, (size_t)S.nthreads * kSlots * kChunk, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
    bool good = cudaEventCreateWithFlags(&S.fence, cudaEventDisableTiming) == cudaSuccess;
    for (int t = 0; t < S.nthreads && good; ++t) {
        good = cudaStreamCreateWithFlags(&S.st[t], cudaStreamNonBlocking) == cudaSuccess;
        for (int k = 0; k < kSlots && good; ++k) good = cudaEventCreateWithFlags(&S.ev[t][k], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!good) { cudaGetLastError(); return false; }      // leaked pieces stay unused; the plain path takes over
    S.ok = true;
    return true;
}

bool is_pageable(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// to_device: host -> pinned slot -> device; otherwise device -> pinned slot -> host
int staged_copy(Stager &S, int dev, char *dst, const char *src, size_t bytes, bool to_device, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(S.mu);
    if (!stager_init(S)) return 1;                       // not available: the caller takes the plain path
    // the copies start after everything already enqueued on the caller's stream (allocation reuse, the solve before a download)
    CM_CUDA(cudaEventRecord(S.fence, stream));
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    const int nt = (int)std::min<size_t>((size_t)S.nthreads, nchunks);
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto worker = [&](int t) {
        if (cudaSetDevice(dev) != cudaSuccess || cudaStreamWaitEvent(S.st[t], S.fence, 0) != cudaSuccess) { failed = 1; return; }
        char *slot[kSlots];
        for (int k = 0; k < kSlots; ++k) slot[k] = S.pin + ((size_t)t * kSlots + k) * kChunk;
        size_t prev_c = 0; int prev_k = -1;
        for (int k = 0;; k ^= 1) {
            const size_t c = next.fetch_add(1);
            const bool have = c < nchunks && !failed.load();
            const size_t off = c * kChunk, len = have ? std::min(kChunk, bytes - off) : 0;
            if (to_device) {
                if (!have) break;
                // the slot's previous DMA must have left it
                if (cudaEventSynchronize(S.ev[t][k]) != cudaSuccess) { failed = 1; break; }
                memcpy(slot[k], src + off, len);
                if (cudaMemcpyAsync(dst + off, slot[k], len, cudaMemcpyHostToDevice, S.st[t]) != cudaSuccess ||
                    cudaEventRecord(S.ev[t][k], S.st[t]) != cudaSuccess) { failed = 1; break; }
            } else {
                if (have && (cudaMemcpyAsync(slot[k], src + off, len, cudaMemcpyDeviceToHost, S.st[t]) != cudaSuccess ||
                             cudaEventRecord(S.ev[t][k], S.st[t]) != cudaSuccess)) { failed = 1; break; }
                if (prev_k >= 0) {
                    if (cudaEventSynchronize(S.ev[t][prev_k]) != cudaSuccess) { failed = 1; break; }
                    memcpy(dst + prev_c * kChunk, slot[prev_k], std::min(kChunk, bytes - prev_c * kChunk));
                }
                if (!have) break;
                prev_c = c; prev_k = k;
            }
        }
    };
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 1; t < nt; ++t) {
        try { th.emplace_back(worker, t); }
        catch (...) { break; }                  // no more threads to be had: the chunk counter hands their work to the others
    }
    worker(0);
    for (auto &x : th) x.join();
    if (failed.load()) {
        cudaError_t e = cudaGetLastError();
        set_error("staged host copy failed (%s)", cudaGetErrorName(e));
        return CUDAMAT_E_CUDA;
    }
    if (to_device) {
        // the caller's stream continues after the last DMA of every worker stream
        for (int t = 0; t < nt; ++t) {
            CM_CUDA(cudaEventRecord(S.ev[t][0], S.st[t]));     // recorded after both slots' copies on that stream
            CM_CUDA(cudaStreamWaitEvent(stream, S.ev[t][0], 0));
        }
    }
    return CUDAMAT_OK;
}
}  // namespace

int copy_h2d(void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    if (!bytes) return CUDAMAT_OK;
    int dev = 0;
    if (bytes >= kStagedMin && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevStage && is_pageable(src)) {
        int rc = staged_copy(g_stagers[dev], dev, (char *)dst, (const char *)src, bytes, true, stream);
        if (rc <= 0) return rc;
    }
    CM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return CUDAMAT_OK;
}

int copy_d2h(void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    if (!bytes) return CUDAMAT_OK;
    int dev = 0;
    if (bytes >= kStagedMin && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevStage && is_pageable(dst)) {
        int rc = staged_copy(g_stagers[dev], dev, (char *)dst, (const char *)src, bytes, false, stream);
        if (rc <= 0) return rc;                // done (complete on the host at return) or failed
    }
    CM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    return CUDAMAT_OK;
}

}  // namespace cudamat
