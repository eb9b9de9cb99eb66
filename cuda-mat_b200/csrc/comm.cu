// comm.cu — multi-GPU plumbing: one process per GPU, row-sharded matrix.
//   * halo exchange of the SpMV operand (p and s each iteration) point-to-point between the ranks
//     that share matrix columns (NVLink/NVSwitch through ncclSend/ncclRecv in one group);
//   * one fused small allreduce per reduction point: every rank deposits the partial sums of ITS rows
//     into a zero-padded array indexed by the GLOBAL tile/group, the sum over ranks therefore only adds
//     zeros to each entry (bit-exact, order independent) and every rank then finishes the fixed-order
//     tree redundantly — the result is bit-identical to the single-GPU run.
// NCCL is dlopen()ed on first use: the single-GPU library has no link-time dependency on it.
// The reference has no counterpart ("Distributed communication backend: none").
#include "solver.h"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <mutex>

namespace cudamat {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
// One NCCL communicator per process, created by the first cudamat_comm_init and shared by every later handle of
// the same (rank, world): bootstrapping a communicator costs ~0.1-1 s, a solver handle should not.
static ncclComm_t g_comm = nullptr;
static int g_comm_rank = -1, g_comm_world = 0;

static int load_nccl() {
    if (g_nccl.lib) return CUDAMAT_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { set_error("NCCL not found: dlopen(libnccl.so.2) failed: %s", dlerror()); return CUDAMAT_E_COMM; }
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return CUDAMAT_E_COMM; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = h;
    return CUDAMAT_OK;
}
#define CM_NCCL(call)                                                                         \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) { set_error("NCCL error %d (%s) at %s:%d", (int)r_, g_nccl.GetErrorString(r_), __FILE__, __LINE__); return CUDAMAT_E_COMM; } \
    } while (0)

// peer-memory path (CUDA IPC over NVLink), see internal.cuh
struct P2P {
    bool on = false;
    unsigned char *arena = nullptr;            // [flags 4 KB][gather parity 0][gather parity 1]
    unsigned long long *flags = nullptr;       // [0, kMaxWorld): reduction arrival per source rank; [kMaxWorld*(1+slot) + src]: halo slot
    double *gather[2] = {nullptr, nullptr};
    std::vector<unsigned char *> peer_arena;   // per rank (self: own pointer)
    std::vector<double *> peer_work;           // per rank, nullptr unless a halo neighbour
    std::vector<unsigned long long> peer_work_elems;
    std::vector<long long> peer_n;
    std::vector<int> peer_recv_off_me;         // offset of MY rows inside rank r's halo region
    PeerRed *d_peers = nullptr;
    unsigned *d_local_cnt = nullptr;           // [2 slots][kMaxPeer]
    unsigned char *d_tile_wait = nullptr;
    unsigned long long red_epoch = 0, halo_epoch[2] = {0, 0};
    std::vector<int> nbr;                      // ranks I exchange halo rows with (send side)
    std::vector<int> src;                      // ranks I receive halo rows from
};

// Mappings of the neighbours' work arenas are cached for the life of the process (keyed by the IPC handle): a recycled
// arena keeps its handle, so the next solver handle finds the mapping instead of paying cudaIpcOpenMemHandle again.
struct PeerMap { cudaIpcMemHandle_t h; void *ptr; };
static std::vector<PeerMap> g_peer_maps;
static std::mutex g_peer_maps_mu;
static void *peer_map_open(const cudaIpcMemHandle_t &h) {
    std::lock_guard<std::mutex> lk(g_peer_maps_mu);
    for (const PeerMap &m : g_peer_maps) if (!memcmp(&m.h, &h, sizeof h)) return m.ptr;
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_peer_maps.push_back(PeerMap{h, p});
    return p;
}

struct Comm {
    P2P p2p;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    std::vector<int64_t> row_starts;
    std::vector<int> recv_cnt, recv_off, send_cnt, send_off, send_contig;   // per peer
    int nsend = 0;
    int *d_send_idx = nullptr;
    double *d_sendbuf = nullptr;
    int *d_ja_local = nullptr;
    int exch_level = 2;
    int exch_count = 0;          // doubles in the exchange arrays
    double *d_exch_local = nullptr, *d_exch_glob = nullptr;
    int ntile_global = 0;
};

__global__ void k_remap_cols(int64_t nnz, const int *ja_global, int *ja_local, int64_t row0, int64_t row1, int n,
                             const int *halo_cols, int nhalo) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) {
        const int c = ja_global[k];
        if (c >= row0 && c < row1) { ja_local[k] = (int)(c - row0); continue; }
        int lo = 0, hi = nhalo - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (halo_cols[mid] < c) lo = mid + 1; else hi = mid; }
        ja_local[k] = n + lo;
    }
}
// out-of-shard column ids: warp-aggregated append (out == nullptr: count only)
__global__ void k_collect_halo(int64_t nnz, const int *ja_global, int64_t row0, int64_t row1, int *out, unsigned long long *cnt) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t k0 = k - lane; k0 < nnz; k0 += stride) {             // warp-uniform trip count
        const int64_t kk = k0 + lane;
        const int c = kk < nnz ? ja_global[kk] : 0;
        const bool outside = kk < nnz && (c < row0 || c >= row1);
        const unsigned m = __ballot_sync(0xffffffffu, outside);
        if (!m) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cnt, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (outside && out) out[base + __popc(m & ((1u << lane) - 1u))] = c;
    }
}
__global__ void k_pack(const double *vec, const int *idx, double *out, int cnt) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < cnt) out[k] = vec[idx[k]];
}

// MARCH on a slab shard reads the neighbours' boundary planes straight out of the operand's halo region: that needs the
// halo to be exactly one whole plane (D entries) of rank - 1 and / or of rank + 1 — nothing else.  The halo columns are
// sorted by global id, so D distinct ids below row0 within reach D are the plane [row0 - D, row0) in natural order.
bool comm_halo_planes(const cudamat_solver *s, int D, int *lo_base, int *hi_base) {
    const Comm *c = s->comm;
    *lo_base = *hi_base = -1;
    if (!c || c->world < 2 || (int)c->recv_cnt.size() != c->world) return false;
    for (int p = 0; p < c->world; ++p) {
        if (c->recv_cnt[p] == 0) continue;
        if (c->recv_cnt[p] != D) return false;
        if (p == c->rank - 1) *lo_base = s->n + c->recv_off[p];
        else if (p == c->rank + 1) *hi_base = s->n + c->recv_off[p];
        else return false;
    }
    if (c->rank > 0 && *lo_base < 0 && s->row0 > 0) return false;
    if (c->rank + 1 < c->world && *hi_base < 0 && s->row1 < s->n_global) return false;
    return true;
}

// ---- hooks used by solver.cu ---------------------------------------------------------------------
int comm_halo_exchange(cudamat_solver *s, double *vec) {
    Comm *c = s->comm;
    if (!c || c->world == 1) return CUDAMAT_OK;
    if (c->nsend > 0) {
        bool need_pack = false;
        for (int p = 0; p < c->world; ++p) if (c->send_cnt[p] > 0 && c->send_contig[p] < 0) need_pack = true;
        if (need_pack) {
            k_pack<<<(c->nsend + 255) / 256, 256, 0, s->stream>>>(vec, c->d_send_idx, c->d_sendbuf, c->nsend);
            s->launches++;
            CM_CUDA(cudaGetLastError());
        }
    }
    CM_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < c->world; ++p) {
        if (c->send_cnt[p] > 0) {
            const double *src = c->send_contig[p] >= 0 ? vec + c->send_contig[p] : c->d_sendbuf + c->send_off[p];
            CM_NCCL(g_nccl.Send(src, (size_t)c->send_cnt[p], ncclDouble, p, c->comm, s->stream));
        }
        if (c->recv_cnt[p] > 0)
            CM_NCCL(g_nccl.Recv(vec + s->n + c->recv_off[p], (size_t)c->recv_cnt[p], ncclDouble, p, c->comm, s->stream));
    }
    CM_NCCL(g_nccl.GroupEnd());
    return CUDAMAT_OK;
}

// Every reducing kernel is followed by this: groups, cross-rank exchange (if sharded), final sum, recurrence.
// `rc` is the reduction context the reducing kernel was launched with (it carries the p2p epoch).
int finish_reduction(cudamat_solver *s, const RedCtx &rc, int phase, int nq) {
    Comm *c = s->comm;
    if (s->n == 0 && (!c || c->world == 1)) return CUDAMAT_OK;
    if (!c || c->world == 1 || c->p2p.on)
        return launch_reduce_finish(s, rc, nq, phase, 0, nullptr, c ? c->p2p.flags : nullptr);
    // NCCL path: local tile (and group) partials -> allreduce of the zero-padded partials -> final
    int r;
    if ((r = launch_reduce_finish(s, rc, nq, phase, 1, nullptr, nullptr))) return r;
    CM_NCCL(g_nccl.AllReduce(c->d_exch_local, c->d_exch_glob, (size_t)c->exch_count, ncclDouble, ncclSum, c->comm, s->stream));
    return launch_reduce_finish(s, rc, nq, phase, 2, c->d_exch_glob, nullptr);
}

bool comm_p2p(const cudamat_solver *s) { return s->comm && s->comm->world > 1 && s->comm->p2p.on; }

void comm_begin_reduction(cudamat_solver *s, RedCtx &rc) {
    if (!comm_p2p(s)) return;
    rc.p2p.epoch = ++s->comm->p2p.red_epoch;
}

// kernel that writes `vec` (one of the work vectors): where its halo rows go. slot 0 / 1 = the two vectors pushed
// per iteration (p, s); returns false when the peer-memory path is off
bool comm_halo_push(cudamat_solver *s, double *vec, int slot, HaloPush *hp) {
    *hp = HaloPush{};
    if (!comm_p2p(s)) return false;
    Comm *c = s->comm; P2P &P = c->p2p;
    const size_t k = (size_t)(vec - s->work) / s->work_elems;            // which work vector
    hp->epoch = ++P.halo_epoch[slot];
    hp->local_cnt = P.d_local_cnt + slot * kMaxPeer;
    for (int r : P.nbr) {
        const int q = hp->npeer++;
        hp->dst[q] = P.peer_work[r] + k * P.peer_work_elems[r] + P.peer_n[r] + P.peer_recv_off_me[r];
        hp->first[q] = c->send_contig[r];
        hp->cnt[q] = c->send_cnt[r];
        hp->ntiles[q] = (hp->first[q] + hp->cnt[q] - 1) / kTile - hp->first[q] / kTile + 1;
        hp->flag[q] = reinterpret_cast<unsigned long long *>(P.peer_arena[r]) + (size_t)kMaxWorld * (1 + slot) + c->rank;
    }
    return true;
}
void comm_halo_wait(cudamat_solver *s, int slot, HaloWait *hw) {
    *hw = HaloWait{};
    if (!comm_p2p(s)) return;
    P2P &P = s->comm->p2p;
    for (int r : P.src) hw->flag[hw->nsrc++] = P.flags + (size_t)kMaxWorld * (1 + slot) + r;
    hw->tile_wait = P.d_tile_wait;
    hw->epoch = P.halo_epoch[slot];
}

__global__ void k_tile_wait(int n, const int *ia, const int *ja, unsigned char *tile_wait) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    for (int k = ia[row]; k < ia[row + 1]; ++k)
        if (ja[k] >= n) { tile_wait[row / kTile] = 1; return; }
}

static void p2p_release(Comm *c) {
    P2P &P = c->p2p;
    for (size_t r = 0; r < P.peer_arena.size(); ++r)
        if ((int)r != c->rank && P.peer_arena[r]) cudaIpcCloseMemHandle(P.peer_arena[r]);
    // peer work-arena mappings stay cached (g_peer_maps)
    if (P.arena) cudaFree(P.arena);
    if (P.d_peers) cudaFree(P.d_peers);
    if (P.d_local_cnt) cudaFree(P.d_local_cnt);
    if (P.d_tile_wait) cudaFree(P.d_tile_wait);
    P = P2P();
}

// Peer-memory set-up (collective). Any failure on any rank switches the path off everywhere (NCCL path stays).
static int p2p_setup(cudamat_solver *s, const std::vector<int> &W) {
    Comm *c = s->comm; P2P &P = c->p2p;
    const int world = c->world, rank = c->rank;
    const char *env = getenv("CUDAMAT_NO_P2P");
    int ok = !(env && *env && *env != '0') && world <= kMaxWorld;
    // halo sends must be contiguous row ranges (row slabs of a banded matrix are), at most kMaxPeer neighbours each way
    for (int r = 0; r < world && ok; ++r) {
        if (c->send_cnt[r] > 0) { if (c->send_contig[r] < 0) ok = 0; P.nbr.push_back(r); }
        if (c->recv_cnt[r] > 0) P.src.push_back(r);
    }
    if ((int)P.nbr.size() > kMaxPeer || (int)P.src.size() > kMaxPeer) ok = 0;
    struct Xchg { cudaIpcMemHandle_t work, arena; unsigned long long work_elems; long long n; int ok; int pad; };
    Xchg mine{}; memset(&mine, 0, sizeof mine);
    const size_t gbytes = sizeof(double) * (size_t)c->exch_count;
    // rounded to whole 2 MB blocks: an IPC handle names the cudaMalloc BLOCK, which must not be shared with any other buffer (solver.cu)
    const size_t abytes = ((4096 + 2 * ((gbytes + 255) / 256) * 256) + (2u << 20) - 1) / (2u << 20) * (2u << 20);
    if (s->work && s->work_pooled) {                  // an arena from the memory pool cannot be shared over IPC
        cudaStreamSynchronize(s->stream); dev_free(s->work); s->work = nullptr; s->work_nvec = 0;
    }
    // the arena is sized ONCE for the largest solve (its IPC handle goes to the neighbours: it must never be reallocated
    // while the peer-memory path is on, see ensure_work)
    if (ok && ensure_work(s, kWorkVecsShared) != CUDAMAT_OK) ok = 0;
    if (ok && cudaMalloc(&P.arena, abytes) != cudaSuccess) { ok = 0; P.arena = nullptr; }
    if (ok) {
        cudaMemsetAsync(P.arena, 0, abytes, s->stream);
        P.flags = reinterpret_cast<unsigned long long *>(P.arena);
        P.gather[0] = reinterpret_cast<double *>(P.arena + 4096);
        P.gather[1] = reinterpret_cast<double *>(P.arena + 4096 + ((gbytes + 255) / 256) * 256);
        if (cudaIpcGetMemHandle(&mine.work, s->work) != cudaSuccess || cudaIpcGetMemHandle(&mine.arena, P.arena) != cudaSuccess) ok = 0;
        mine.work_elems = s->work_elems; mine.n = s->n;
    }
    cudaGetLastError();
    mine.ok = ok;
    // all-gather the handles
    unsigned char *d_x = nullptr;
    CM_CUDA(cudaMalloc(&d_x, sizeof(Xchg) * (size_t)(world + 1)));
    CM_CUDA(cudaMemcpyAsync(d_x + sizeof(Xchg) * (size_t)world, &mine, sizeof mine, cudaMemcpyHostToDevice, s->stream));
    CM_NCCL(g_nccl.AllGather(d_x + sizeof(Xchg) * (size_t)world, d_x, sizeof(Xchg), ncclUint8, c->comm, s->stream));
    std::vector<Xchg> all((size_t)world);
    CM_CUDA(cudaMemcpyAsync(all.data(), d_x, sizeof(Xchg) * (size_t)world, cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    for (int r = 0; r < world; ++r) if (!all[r].ok) ok = 0;
    int opened = ok;
    if (ok) {
        P.peer_arena.assign(world, nullptr); P.peer_work.assign(world, nullptr);
        P.peer_work_elems.assign(world, 0); P.peer_n.assign(world, 0); P.peer_recv_off_me.assign(world, 0);
        for (int r = 0; r < world && opened; ++r) {
            P.peer_work_elems[r] = all[r].work_elems; P.peer_n[r] = all[r].n;
            int off = 0;                                           // rank r's halo region: sources in rank order
            for (int q = 0; q < rank; ++q) off += W[(size_t)r * world + q];
            P.peer_recv_off_me[r] = off;
            if (r == rank) { P.peer_arena[r] = P.arena; P.peer_work[r] = s->work; continue; }
            void *pa = nullptr;
            if (cudaIpcOpenMemHandle(&pa, all[r].arena, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened = 0; break; }
            P.peer_arena[r] = (unsigned char *)pa;
            if (c->send_cnt[r] > 0) {
                void *pw = peer_map_open(all[r].work);
                if (!pw) { opened = 0; break; }
                P.peer_work[r] = (double *)pw;
            }
        }
        cudaGetLastError();
    }
    // agree on the outcome
    int *d_ok = reinterpret_cast<int *>(d_x);
    CM_CUDA(cudaMemcpyAsync(d_ok, &opened, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CM_NCCL(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->comm, s->stream));
    CM_CUDA(cudaMemcpyAsync(&opened, d_ok, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    CM_CUDA(cudaFree(d_x));
    if (!opened) { p2p_release(c); return CUDAMAT_OK; }
    // device tables
    std::vector<PeerRed> pr((size_t)world);
    for (int r = 0; r < world; ++r) {
        unsigned char *pa = P.peer_arena[r];
        pr[r].gather[0] = reinterpret_cast<double *>(pa + 4096);
        pr[r].gather[1] = reinterpret_cast<double *>(pa + 4096 + ((gbytes + 255) / 256) * 256);
        pr[r].flag = reinterpret_cast<unsigned long long *>(pa) + rank;
    }
    CM_CUDA(cudaMalloc(&P.d_peers, sizeof(PeerRed) * (size_t)world));
    CM_CUDA(cudaMemcpyAsync(P.d_peers, pr.data(), sizeof(PeerRed) * (size_t)world, cudaMemcpyHostToDevice, s->stream));
    CM_CUDA(cudaMalloc(&P.d_local_cnt, sizeof(unsigned) * 2 * kMaxPeer));
    CM_CUDA(cudaMemsetAsync(P.d_local_cnt, 0, sizeof(unsigned) * 2 * kMaxPeer, s->stream));
    const int ntile = (s->n + kTile - 1) / kTile;
    CM_CUDA(cudaMalloc(&P.d_tile_wait, (size_t)std::max(ntile, 1)));
    CM_CUDA(cudaMemsetAsync(P.d_tile_wait, 0, (size_t)std::max(ntile, 1), s->stream));
    if (s->n > 0) k_tile_wait<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->n, s->d_ia, s->d_ja, P.d_tile_wait);
    CM_CUDA(cudaGetLastError());
    CM_CUDA(cudaStreamSynchronize(s->stream));
    RedCtx &rcx = s->rc;
    rcx.p2p.world = world;
    rcx.p2p.me = rank;
    rcx.p2p.peers = P.d_peers;
    rcx.p2p.stride = c->exch_count / kMaxQ;
    if (c->exch_level == 2) { rcx.p2p.my_off = rcx.group0; rcx.p2p.my_cnt = rcx.ngroup_loc; }
    else { rcx.p2p.my_off = rcx.tile0; rcx.p2p.my_cnt = rcx.ntile; }
    P.on = true;
    return CUDAMAT_OK;
}

void comm_epochs_get(const cudamat_solver *s, unsigned long long e[3]) {
    e[0] = e[1] = e[2] = 0;
    if (!comm_p2p(s)) return;
    const P2P &P = s->comm->p2p;
    e[0] = P.halo_epoch[0]; e[1] = P.halo_epoch[1]; e[2] = P.red_epoch;
}
void comm_epochs_set(cudamat_solver *s, const unsigned long long e[3]) {
    if (!comm_p2p(s)) return;
    P2P &P = s->comm->p2p;
    P.halo_epoch[0] = e[0]; P.halo_epoch[1] = e[1]; P.red_epoch = e[2];
}
const unsigned long long *comm_red_flags(const cudamat_solver *s) { return comm_p2p(s) ? s->comm->p2p.flags : nullptr; }

void comm_release(cudamat_solver *s) {
    Comm *c = s->comm;
    if (!c) return;
    p2p_release(c);
    s->rc.p2p = P2PRed{};
    if (c->d_send_idx) cudaFree(c->d_send_idx);
    if (c->d_sendbuf) cudaFree(c->d_sendbuf);
    if (c->d_ja_local) cudaFree(c->d_ja_local);
    if (c->d_exch_local) cudaFree(c->d_exch_local);
    if (c->d_exch_glob) cudaFree(c->d_exch_glob);
    delete c;
    s->comm = nullptr;
}

}  // namespace cudamat
using namespace cudamat;

extern "C" {

// ---- pure host planners (no GPU / NCCL needed; exercised by the gloo CPU tests) ----------------------
// Row shard of `rank`: contiguous, boundaries aligned to the reduction group (2 Mi rows) when every rank
// can get at least one group, else to the reduction tile (2048 rows).
int cudamat_partition_rows(int64_t n_global, int world, int rank, int64_t *row0, int64_t *row1) {
    if (n_global < 0 || world < 1 || rank < 0 || rank >= world || !row0 || !row1) { set_error("partition_rows: invalid argument"); return CUDAMAT_E_INVALID; }
    const int64_t group = (int64_t)kTile * kGroupTiles;
    const int64_t gran = (n_global >= group * world) ? group : kTile;
    const int64_t per = ((n_global + world - 1) / world + gran - 1) / gran * gran;
    *row0 = std::min<int64_t>((int64_t)rank * per, n_global);
    *row1 = std::min<int64_t>(*row0 + per, n_global);
    return CUDAMAT_OK;
}

// Halo plan of one shard: the sorted unique global columns outside [row0,row1) (malloc()ed into
// *halo_cols, release with cudamat_free) and how many of them each owner rank holds (recv_cnt[world]).
// Local column numbering is then: c in [row0,row1) -> c-row0, else n_local + position in halo_cols.
int cudamat_halo_plan_host(int64_t row0, int64_t row1, int64_t nnz, const int *ja_global, int world,
                           const int64_t *row_starts, int *nhalo, int **halo_cols, int *recv_cnt) {
    if (!ja_global && nnz > 0) { set_error("halo_plan: null argument"); return CUDAMAT_E_INVALID; }
    if (!row_starts || !nhalo || !halo_cols || !recv_cnt || world < 1) { set_error("halo_plan: null argument"); return CUDAMAT_E_INVALID; }
    std::vector<int> halo;
    for (int64_t k = 0; k < nnz; ++k) if (ja_global[k] < row0 || ja_global[k] >= row1) halo.push_back(ja_global[k]);
    std::sort(halo.begin(), halo.end());
    halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
    for (int p = 0; p < world; ++p) recv_cnt[p] = 0;
    int p = 0;
    for (size_t k = 0; k < halo.size(); ++k) {
        if (halo[k] < 0 || halo[k] >= row_starts[world]) { set_error("halo_plan: column %d out of range", halo[k]); return CUDAMAT_E_INVALID; }
        while (halo[k] >= row_starts[p + 1]) ++p;
        recv_cnt[p]++;
    }
    *nhalo = (int)halo.size();
    *halo_cols = (int *)malloc(sizeof(int) * std::max<size_t>(halo.size(), 1));
    if (!*halo_cols) { set_error("halo_plan: out of memory"); return CUDAMAT_E_INVALID; }
    std::copy(halo.begin(), halo.end(), *halo_cols);
    return CUDAMAT_OK;
}

int cudamat_comm_p2p_enabled(cudamat_solver *s) { return (s && comm_p2p(s)) ? 1 : 0; }

int cudamat_comm_unique_id(void *id128) {
    if (!id128) return CUDAMAT_E_INVALID;
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    CM_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == CUDAMAT_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(id128, &id, sizeof id);
    return CUDAMAT_OK;
}

// Collective over all ranks; call after cudamat_set_csr_* and before cudamat_analyze.
int cudamat_comm_init(cudamat_solver *s, const void *id128, int rank, int world) {
    if (!s || !id128 || world < 1 || rank < 0 || rank >= world) { set_error("comm_init: invalid argument"); return CUDAMAT_E_INVALID; }
    if (!s->d_ia) { set_error("comm_init: set the CSR shard first"); return CUDAMAT_E_STATE; }
    DeviceGuard dg(s->device);
    int rc = load_nccl();
    if (rc) return rc;
    comm_release(s);
    Comm *c = new Comm();
    s->comm = c;
    c->rank = rank; c->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    if (g_comm && g_comm_rank == rank && g_comm_world == world) c->comm = g_comm;       // id128 is ignored on reuse
    else {
        if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
        CM_NCCL(g_nccl.CommInitRank(&c->comm, world, id, rank));
        g_comm = c->comm; g_comm_rank = rank; g_comm_world = world;
    }
    // 1. everybody's row range
    int64_t *d_rng = nullptr;
    CM_CUDA(cudaMalloc(&d_rng, sizeof(int64_t) * 2 * (size_t)(world + 1)));
    const int64_t mine[2] = {s->row0, s->row1};
    CM_CUDA(cudaMemcpyAsync(d_rng + 2 * world, mine, sizeof mine, cudaMemcpyHostToDevice, s->stream));
    CM_NCCL(g_nccl.AllGather(d_rng + 2 * world, d_rng, 2, ncclInt64, c->comm, s->stream));
    std::vector<int64_t> rng(2 * (size_t)world);
    CM_CUDA(cudaMemcpyAsync(rng.data(), d_rng, sizeof(int64_t) * 2 * (size_t)world, cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    CM_CUDA(cudaFree(d_rng));
    c->row_starts.assign(world + 1, 0);
    for (int p = 0; p < world; ++p) {
        c->row_starts[p] = rng[2 * p];
        if (p > 0 && rng[2 * p] != rng[2 * p - 1]) { set_error("comm_init: row shards are not contiguous in rank order"); return CUDAMAT_E_INVALID; }
        if (rng[2 * p] % kTile != 0) { set_error("comm_init: shard boundaries must be multiples of %d rows", kTile); return CUDAMAT_E_INVALID; }
    }
    c->row_starts[world] = rng[2 * world - 1];
    if (c->row_starts[0] != 0 || c->row_starts[world] != s->n_global) { set_error("comm_init: shards do not cover [0,n)"); return CUDAMAT_E_INVALID; }
    // 2. halo columns of this shard: the out-of-range column ids are collected on the device (two passes: count,
    //    append), only that short list travels to the host where cudamat_halo_plan_host sorts / uniquifies it
    unsigned long long *d_cntx = nullptr;
    CM_CUDA(cudaMalloc(&d_cntx, sizeof(unsigned long long)));
    CM_CUDA(cudaMemsetAsync(d_cntx, 0, sizeof(unsigned long long), s->stream));
    if (s->nnz > 0) k_collect_halo<<<1184, 256, 0, s->stream>>>(s->nnz, s->d_ja_global, s->row0, s->row1, nullptr, d_cntx);
    unsigned long long nout = 0;
    CM_CUDA(cudaMemcpyAsync(&nout, d_cntx, sizeof nout, cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<int> ja((size_t)nout);
    if (nout > 0) {
        int *d_out = nullptr;
        CM_CUDA(cudaMalloc(&d_out, sizeof(int) * (size_t)nout));
        CM_CUDA(cudaMemsetAsync(d_cntx, 0, sizeof(unsigned long long), s->stream));
        k_collect_halo<<<1184, 256, 0, s->stream>>>(s->nnz, s->d_ja_global, s->row0, s->row1, d_out, d_cntx);
        CM_CUDA(cudaMemcpyAsync(ja.data(), d_out, sizeof(int) * (size_t)nout, cudaMemcpyDeviceToHost, s->stream));
        CM_CUDA(cudaStreamSynchronize(s->stream));
        CM_CUDA(cudaFree(d_out));
    }
    CM_CUDA(cudaFree(d_cntx));
    int nhalo = 0; int *halo_raw = nullptr;
    c->recv_cnt.assign(world, 0); c->recv_off.assign(world, 0);
    int prc = cudamat_halo_plan_host(s->row0, s->row1, (int64_t)nout, ja.data(), world, c->row_starts.data(), &nhalo, &halo_raw, c->recv_cnt.data());
    if (prc) return prc;
    std::vector<int> halo(halo_raw, halo_raw + nhalo);
    free(halo_raw);
    std::vector<int>().swap(ja);
    for (int p = 1; p < world; ++p) c->recv_off[p] = c->recv_off[p - 1] + c->recv_cnt[p - 1];
    // 3. who wants what from whom: allgather of the world x world count matrix
    int *d_cnt = nullptr;
    CM_CUDA(cudaMalloc(&d_cnt, sizeof(int) * (size_t)world * (size_t)(world + 1)));
    CM_CUDA(cudaMemcpyAsync(d_cnt + (size_t)world * world, c->recv_cnt.data(), sizeof(int) * (size_t)world, cudaMemcpyHostToDevice, s->stream));
    CM_NCCL(g_nccl.AllGather(d_cnt + (size_t)world * world, d_cnt, (size_t)world, ncclInt32, c->comm, s->stream));
    std::vector<int> W((size_t)world * world);
    CM_CUDA(cudaMemcpyAsync(W.data(), d_cnt, sizeof(int) * (size_t)world * world, cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    CM_CUDA(cudaFree(d_cnt));
    c->send_cnt.assign(world, 0); c->send_off.assign(world, 0); c->send_contig.assign(world, -1);
    for (int q = 0; q < world; ++q) c->send_cnt[q] = W[(size_t)q * world + rank];
    for (int q = 1; q < world; ++q) c->send_off[q] = c->send_off[q - 1] + c->send_cnt[q - 1];
    c->nsend = c->send_off[world - 1] + c->send_cnt[world - 1];
    // 4. exchange the wanted global column lists; owners turn them into local row indices
    int *d_halo = nullptr, *d_want = nullptr;
    CM_CUDA(cudaMalloc(&d_halo, sizeof(int) * (size_t)std::max(nhalo, 1)));
    CM_CUDA(cudaMalloc(&d_want, sizeof(int) * (size_t)std::max(c->nsend, 1)));
    CM_CUDA(cudaMemcpyAsync(d_halo, halo.data(), sizeof(int) * (size_t)nhalo, cudaMemcpyHostToDevice, s->stream));
    CM_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < world; ++p) {
        if (c->recv_cnt[p] > 0) CM_NCCL(g_nccl.Send(d_halo + c->recv_off[p], (size_t)c->recv_cnt[p], ncclInt32, p, c->comm, s->stream));
        if (c->send_cnt[p] > 0) CM_NCCL(g_nccl.Recv(d_want + c->send_off[p], (size_t)c->send_cnt[p], ncclInt32, p, c->comm, s->stream));
    }
    CM_NCCL(g_nccl.GroupEnd());
    std::vector<int> want((size_t)c->nsend);
    CM_CUDA(cudaMemcpyAsync(want.data(), d_want, sizeof(int) * (size_t)c->nsend, cudaMemcpyDeviceToHost, s->stream));
    CM_CUDA(cudaStreamSynchronize(s->stream));
    for (int k = 0; k < c->nsend; ++k) want[k] -= (int)s->row0;
    for (int q = 0; q < world; ++q) {
        bool contig = c->send_cnt[q] > 0;
        for (int k = 1; k < c->send_cnt[q] && contig; ++k)
            if (want[c->send_off[q] + k] != want[c->send_off[q]] + k) contig = false;
        if (contig) c->send_contig[q] = want[c->send_off[q]];
    }
    CM_CUDA(cudaMalloc(&c->d_send_idx, sizeof(int) * (size_t)std::max(c->nsend, 1)));
    CM_CUDA(cudaMemcpyAsync(c->d_send_idx, want.data(), sizeof(int) * (size_t)c->nsend, cudaMemcpyHostToDevice, s->stream));
    CM_CUDA(cudaMalloc(&c->d_sendbuf, sizeof(double) * (size_t)std::max(c->nsend, 1)));
    CM_CUDA(cudaFree(d_want));
    // 5. local/halo column numbering
    CM_CUDA(cudaMalloc(&c->d_ja_local, sizeof(int) * (size_t)std::max<int64_t>(s->nnz, 1) + 16));
    if (s->nnz > 0) {
        k_remap_cols<<<1184, 256, 0, s->stream>>>(s->nnz, s->d_ja_global, c->d_ja_local, s->row0, s->row1, s->n, d_halo, nhalo);
        CM_CUDA(cudaGetLastError());
    }
    CM_CUDA(cudaStreamSynchronize(s->stream));
    CM_CUDA(cudaFree(d_halo));
    s->d_ja = c->d_ja_local;
    s->nhalo = nhalo;
    // 6. reduction exchange level
    bool group_aligned = true;
    for (int p = 0; p < world; ++p) if (c->row_starts[p] % ((int64_t)kTile * kGroupTiles) != 0) group_aligned = false;
    c->ntile_global = (int)((s->n_global + kTile - 1) / kTile);
    RedCtx &rcx = s->rc;
    if (group_aligned) {
        c->exch_level = 2;
        c->exch_count = kMaxQ * rcx.slot_stride;
    } else {
        if (c->ntile_global > 64 * kGroupTiles) { set_error("comm_init: unaligned shards need n <= %lld", (long long)64 * kGroupTiles * kTile); return CUDAMAT_E_INVALID; }
        c->exch_level = 1;
        c->exch_count = kMaxQ * c->ntile_global;
    }
    CM_CUDA(cudaMalloc(&c->d_exch_local, sizeof(double) * (size_t)c->exch_count));
    CM_CUDA(cudaMalloc(&c->d_exch_glob, sizeof(double) * (size_t)c->exch_count));
    CM_CUDA(cudaMemsetAsync(c->d_exch_local, 0, sizeof(double) * (size_t)c->exch_count, s->stream));
    CM_CUDA(cudaMemsetAsync(c->d_exch_glob, 0, sizeof(double) * (size_t)c->exch_count, s->stream));
    rcx.exch_level = c->exch_level;
    rcx.tile0 = (int)(s->row0 / kTile);
    if (c->exch_level == 2) {
        // group partials go straight into the exchange array (same layout as the slots buffer,
        // which stays allocated as s->slots_own and is freed at destroy)
        rcx.slots = c->d_exch_local;
        rcx.exch = nullptr; rcx.exch_stride = 0;
    } else {
        rcx.exch = c->d_exch_local; rcx.exch_stride = c->ntile_global;
    }
    CM_CUDA(cudaStreamSynchronize(s->stream));
    s->analyzed = false;
    // 7. peer-memory path (CUDA IPC): fused halo pushes and reduction gathers without NCCL calls in the loop
    return p2p_setup(s, W);
}

}  // extern "C"
