// hb_comm.cu — multi-GPU data plane (see hb_comm.cuh).
//
// NCCL is resolved with dlopen at hb_comm_init: the library has no link-time dependency on it, a single-GPU process
// never loads it, and a process that already holds torch's libnccl.so.2 shares that copy.
//
// Search exchange over peer memory.  Every rank owns a window (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by
// every peer): [2 parities][nranks slots][slot_bytes] of (distance, id) blocks + [2][nranks][kFlagCtas] epoch flags.
// exchange_merge_kernel<true>, one launch per search: CTA c owns the queries q = c, c + grid, ...; it writes their local
// top-k into slot [rank] of EVERY rank's window (plain stores to peer addresses travel over NVLink), publishes
// flag[rank][c] = epoch with st.release.sys in every window, waits until flag[r][c] == epoch for all r in its own window
// (ld.acquire.sys) and merges the nranks sorted lists of its queries.  Two parities: a rank can be at most one search
// ahead of a peer (it cannot finish search e + 1 before the peer has pushed e + 1, which the peer does after it has merged e).
// The wait is bounded (kWaitNs): a peer that never arrives turns into HB_ERR_CUDA, not a hung GPU.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "hb_comm.cuh"

namespace hb {
namespace {

// ---- the few NCCL declarations we need (stable since NCCL 2.0; no header dependency) ------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { kNcclInt8 = 0, kNcclUint8 = 1, kNcclInt64 = 4, kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclMax = 2 };
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

void load_nccl() {
    if (g_nccl.h) return;
    const char *env = getenv("HB_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) throw Error(HB_ERR_UNSUPPORTED, std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "?"));
    auto sym = [&](const char *name) {
        void *p = dlsym(h, name);
        if (!p) throw Error(HB_ERR_UNSUPPORTED, std::string("NCCL lacks ") + name);
        return p;
    };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    g_nccl.h = h;
}
#define HB_NCCL(expr)                                                                                         \
    do {                                                                                                      \
        int _r = (expr);                                                                                      \
        if (_r != 0) throw Error(HB_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));         \
    } while (0)

constexpr int kFlagCtas = 256;              // flag slots per (parity, source rank)
constexpr int kExchangeGrid = 128;          // CTAs of the fused kernel: all co-resident on 148 SMs
constexpr int kExchangeThreads = 256;
constexpr unsigned long long kWaitNs = 8ull * 1000 * 1000 * 1000;  // bound of the wait for a peer's flag

struct CommState {
    CommInfo info;
    ncclComm_t comm = nullptr;
    int device = 0;
    bool p2p_ok = false;  // IPC mapping of peer windows works on this box (agreed between all ranks)
    // peer windows
    void *win_local = nullptr;
    void *win[kCommMaxRanks] = {nullptr};
    size_t slot_bytes = 0;
    uint32_t epoch = 0;
    DevBuf pack, gathered, small;
    int32_t *err_dev = nullptr;
};
CommState g_comm;

size_t window_bytes(size_t slot_bytes, int nranks) { return 2 * (size_t)nranks * slot_bytes + 2 * (size_t)nranks * kFlagCtas * 4; }

// ---- device code ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct ExchangeParams {
    int nranks, rank;
    char *win[kCommMaxRanks];  // every rank's window as mapped here (win[rank] = the local one); NCCL path: only [rank]
    size_t slot_bytes;         // stride between the rank slots
    size_t parity_off;         // byte offset of this search's parity half
    size_t flag_off;           // byte offset of the flags of this parity
    uint32_t epoch;
    int64_t nq;
    int k;
    const double *loc_dist;
    const int64_t *loc_ids;
    int64_t id_base;
    double *out_dist;
    int64_t *out_ids;
    int32_t *err;
    int stage_doubles;  // shared-memory doubles per warp for the merge (nranks * k), 0: search in global memory
};

// elements of a sorted list whose key precedes `key` (or_equal: or equals it)
__device__ __forceinline__ int count_before(const double *lst, int k, uint64_t key, bool or_equal) {
    int lo = 0, hi = k;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint64_t km = dist_key(__ldcg(lst + mid));
        if (or_equal ? km <= key : km < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Merge of the nranks sorted (distance, id) lists of query q: the k smallest by (distance, rank, position) — the stable
// sort of the concatenation in rank order, then take k (partitioned_hnsw.clj:187-196, ivf_flat.clj:291-294).
// `stage` (optional): nranks * k doubles of shared memory owned by this warp — the query's lists are copied there once and
// the binary searches run on shared memory (a search step from L2 is ~300 clocks, and a candidate makes nranks * log2 k of them)
__device__ __forceinline__ void merge_query(const ExchangeParams &P, const char *base, int64_t q, int lane, double *stage) {
    const int k = P.k, total = P.nranks * k;
    const size_t ids_off = (size_t)P.nq * k * 8;
    if (stage) {
        __syncwarp();
        for (int c = lane; c < total; c += 32) {
            const int r = c / k, j = c - r * k;
            stage[c] = __ldcg((const double *)(base + (size_t)r * P.slot_bytes) + q * k + j);
        }
        __syncwarp();
    }
    for (int c = lane; c < total; c += 32) {
        const int r = c / k, j = c - r * k;
        const double *lr = (const double *)(base + (size_t)r * P.slot_bytes) + q * k;
        const double dv = stage ? stage[c] : __ldcg(lr + j);
        const uint64_t key = dist_key(dv);
        int pos = j;
        for (int r2 = 0; r2 < P.nranks; ++r2) {
            if (r2 == r) continue;
            if (stage) {
                const double *l2 = stage + r2 * k;
                int lo = 0, hi = k;
                const bool or_equal = r2 < r;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const uint64_t km = dist_key(l2[mid]);
                    if (or_equal ? km <= key : km < key) lo = mid + 1;
                    else hi = mid;
                }
                pos += lo;
            } else {
                const double *l2 = (const double *)(base + (size_t)r2 * P.slot_bytes) + q * k;
                pos += count_before(l2, k, key, r2 < r);
            }
        }
        if (pos < k) {
            const int64_t *ir = (const int64_t *)(base + (size_t)r * P.slot_bytes + ids_off) + q * k;
            P.out_dist[q * k + pos] = dv;
            P.out_ids[q * k + pos] = __ldcg(ir + j);
        }
    }
}

template <bool P2P>
__global__ void __launch_bounds__(kExchangeThreads) exchange_merge_kernel(const ExchangeParams P) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, grid = gridDim.x;
    const int k = P.k;
    __shared__ int s_fail;
    if (P2P) {
        if (tid == 0) s_fail = 0;
        const size_t ids_off = (size_t)P.nq * k * 8;
        // push: this CTA's queries, to every rank's window (slot = my rank)
        const int64_t nmine = P.nq > cta ? (P.nq - cta + grid - 1) / grid : 0;
        const int64_t elems = nmine * k;
        for (int p = 0; p < P.nranks; ++p) {
            char *dst = P.win[p] + P.parity_off + (size_t)P.rank * P.slot_bytes;
            double *dd = (double *)dst;
            int64_t *di = (int64_t *)(dst + ids_off);
            for (int64_t e = tid; e < elems; e += kExchangeThreads) {
                const int64_t qi = e / k, j = e - qi * k;
                const int64_t o = (cta + qi * grid) * k + j;
                const int64_t id = P.loc_ids[o];
                dd[o] = P.loc_dist[o];
                di[o] = id < 0 ? id : id + P.id_base;
            }
        }
        __threadfence_system();
        __syncthreads();
        if (tid < P.nranks) {  // publish: flag [my rank][cta] of rank tid's window
            uint32_t *f = (uint32_t *)(P.win[tid] + P.flag_off) + (size_t)P.rank * kFlagCtas + cta;
            st_release_sys(f, P.epoch);
        }
        if (tid < P.nranks) {  // wait for rank tid's CTA `cta`
            const uint32_t *f = (const uint32_t *)(P.win[P.rank] + P.flag_off) + (size_t)tid * kFlagCtas + cta;
            const unsigned long long t0 = global_ns();
            while (ld_acquire_sys(f) != P.epoch) {
                if (global_ns() - t0 > kWaitNs) {
                    s_fail = 1;
                    atomicExch(P.err, 1);
                    break;
                }
                __nanosleep(200);
            }
        }
        __syncthreads();
        if (s_fail) return;
    }
    const char *base = P.win[P.rank] + P.parity_off;
    extern __shared__ double s_stage[];
    double *stage = P.stage_doubles > 0 ? s_stage + (size_t)warp * P.stage_doubles : nullptr;
    for (int64_t q = cta + (int64_t)warp * grid; q < P.nq; q += (int64_t)grid * (kExchangeThreads / 32)) merge_query(P, base, q, lane, stage);
}

// NCCL path: [dist nq*k][ids nq*k] block to all-gather
__global__ void pack_topk_kernel(const double *__restrict__ dist, const int64_t *__restrict__ ids, int64_t count, int64_t id_base,
                                 double *__restrict__ out_dist, int64_t *__restrict__ out_ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t id = ids[i];
    out_dist[i] = dist[i];
    out_ids[i] = id < 0 ? id : id + id_base;
}

__global__ void divide_centroids_kernel(const double *__restrict__ sums, const int64_t *__restrict__ counts, int nlist, int d,
                                        double *__restrict__ cents) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nlist * d) return;
    const int64_t c = counts[i / d];
    if (c > 0) cents[i] = __ddiv_rn(sums[i], (double)c);
}

template <typename T>
__global__ void init_centroids_sharded_kernel(const T *__restrict__ rows, int d, const int64_t *__restrict__ seed_rows, int nlist,
                                              int64_t first_row, int64_t n_local, double *__restrict__ cents) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nlist * d) return;
    const int c = (int)(i / d), j = (int)(i % d);
    const int64_t r = seed_rows[c] - first_row;
    cents[i] = (r >= 0 && r < n_local) ? to_f64(rows[r * d + j]) : 0.0;
}

void close_windows() {
    CommState &C = g_comm;
    for (int r = 0; r < C.info.nranks; ++r) {
        if (r != C.info.rank && C.win[r]) cudaIpcCloseMemHandle(C.win[r]);
        C.win[r] = nullptr;
    }
    cudaGetLastError();
}

// (Re)allocates the peer windows for slots of at least `need` bytes.  Collective: every rank reaches it with the same need.
void ensure_windows(size_t need) {
    CommState &C = g_comm;
    if (C.win_local && need <= C.slot_bytes) return;
    const int nr = C.info.nranks;
    HB_CUDA(cudaStreamSynchronize(g_stream));
    close_windows();
    comm_barrier();  // nobody maps my old window any more
    if (C.win_local) {
        HB_CUDA(cudaFree(C.win_local));
        C.win_local = nullptr;
    }
    C.slot_bytes = ((need + need / 4 + 4095) / 4096) * 4096;
    const size_t bytes = window_bytes(C.slot_bytes, nr);
    HB_CUDA(cudaMalloc(&C.win_local, bytes));
    HB_CUDA(cudaMemsetAsync(C.win_local, 0, bytes, g_stream));
    C.epoch = 0;
    cudaIpcMemHandle_t mine;
    HB_CUDA(cudaIpcGetMemHandle(&mine, C.win_local));
    char *stage = (char *)C.small.get((size_t)(nr + 1) * sizeof(mine));
    HB_CUDA(cudaMemcpyAsync(stage, &mine, sizeof(mine), cudaMemcpyHostToDevice, g_stream));
    comm_allgather_bytes(stage, stage + sizeof(mine), sizeof(mine));  // also orders every rank's memset before any push
    std::vector<cudaIpcMemHandle_t> all((size_t)nr);
    HB_CUDA(cudaMemcpyAsync(all.data(), stage + sizeof(mine), (size_t)nr * sizeof(mine), cudaMemcpyDeviceToHost, g_stream));
    HB_CUDA(cudaStreamSynchronize(g_stream));
    bool ok = true;
    for (int r = 0; r < nr && ok; ++r) {
        if (r == C.info.rank) {
            C.win[r] = C.win_local;
            continue;
        }
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
        } else {
            C.win[r] = p;
        }
    }
    // all ranks must agree on the path: max over ranks of "failed"
    double *flag = (double *)C.small.get(64);
    const double mine_failed = ok ? 0.0 : 1.0;
    HB_CUDA(cudaMemcpyAsync(flag, &mine_failed, 8, cudaMemcpyHostToDevice, g_stream));
    comm_allreduce_max_f64(flag, 1);
    double any_failed = 0.0;
    HB_CUDA(cudaMemcpyAsync(&any_failed, flag, 8, cudaMemcpyDeviceToHost, g_stream));
    HB_CUDA(cudaStreamSynchronize(g_stream));
    if (any_failed != 0.0) {
        close_windows();
        C.p2p_ok = false;
        C.info.p2p = false;
    } else {
        C.info.p2p = true;
    }
}

}  // namespace

const CommInfo &comm_info() { return g_comm.info; }

void comm_unique_id(void *out128) {
    load_nccl();
    ncclUniqueId id;
    HB_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == kCommIdBytes, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, sizeof(id));
}

void comm_init(const void *id128, int nranks, int rank, int device) {
    CommState &C = g_comm;
    HB_REQUIRE(!C.info.inited, "hb_comm_init: already initialised (call hb_comm_shutdown first)");
    HB_REQUIRE(nranks >= 1 && nranks <= kCommMaxRanks && rank >= 0 && rank < nranks, "hb_comm_init: bad nranks / rank");
    load_nccl();
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    HB_NCCL(g_nccl.CommInitRank(&C.comm, nranks, id, rank));
    C.info.nranks = nranks;
    C.info.rank = rank;
    C.info.inited = true;
    C.info.p2p = false;
    C.device = device;
    const char *e = getenv("HB_COMM_P2P");
    C.p2p_ok = !(e && atoi(e) == 0);
    C.err_dev = nullptr;
    HB_CUDA(cudaMalloc((void **)&C.err_dev, 4));
    HB_CUDA(cudaMemsetAsync(C.err_dev, 0, 4, g_stream));
    comm_barrier();
}

void comm_shutdown() {
    CommState &C = g_comm;
    if (!C.info.inited) return;
    cudaStreamSynchronize(g_stream);
    close_windows();
    if (C.comm) {
        // peers must have unmapped this window before it is freed
        try {
            comm_barrier();
        } catch (...) {
        }
        g_nccl.CommDestroy(C.comm);
    }
    if (C.win_local) cudaFree(C.win_local);
    if (C.err_dev) cudaFree(C.err_dev);
    C.pack.release();
    C.gathered.release();
    C.small.release();
    C = CommState();
}

static void require_comm() { HB_REQUIRE(g_comm.info.inited, "multi-GPU call before hb_comm_init"); }

void comm_allreduce_sum_f64(double *buf, int64_t count) {
    require_comm();
    if (count == 0) return;
    HB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, g_comm.comm, g_stream));
}
void comm_allreduce_sum_i64(int64_t *buf, int64_t count) {
    require_comm();
    if (count == 0) return;
    HB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclInt64, kNcclSum, g_comm.comm, g_stream));
}
void comm_allreduce_max_f64(double *buf, int64_t count) {
    require_comm();
    if (count == 0) return;
    HB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclMax, g_comm.comm, g_stream));
}
void comm_broadcast_bytes(void *buf, int64_t bytes, int root) {
    require_comm();
    HB_REQUIRE(root >= 0 && root < g_comm.info.nranks, "broadcast root out of range");
    if (bytes == 0) return;
    HB_NCCL(g_nccl.Broadcast(buf, buf, (size_t)bytes, kNcclUint8, root, g_comm.comm, g_stream));
}
void comm_allgather_bytes(const void *send, void *recv, int64_t bytes_per_rank) {
    require_comm();
    if (bytes_per_rank == 0) return;
    HB_NCCL(g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, kNcclUint8, g_comm.comm, g_stream));
}
void comm_barrier() {
    require_comm();
    double *x = (double *)g_comm.small.get(64);
    HB_CUDA(cudaMemsetAsync(x, 0, 8, g_stream));
    comm_allreduce_max_f64(x, 1);
    HB_CUDA(cudaStreamSynchronize(g_stream));
}

void comm_topk_exchange_merge(const double *loc_dist, const int64_t *loc_ids, int64_t id_base, int64_t nq, int k, double *out_dist,
                              int64_t *out_ids, int use_p2p, ExchangeStats *st) {
    require_comm();
    CommState &C = g_comm;
    if (nq == 0 || k == 0) return;
    const int nr = C.info.nranks;
    const size_t block = (size_t)nq * k * 16;  // [dist][ids] of one rank
    ExchangeParams P{};
    P.nranks = nr;
    P.rank = C.info.rank;
    P.nq = nq;
    P.k = k;
    P.loc_dist = loc_dist;
    P.loc_ids = loc_ids;
    P.id_base = id_base;
    P.out_dist = out_dist;
    P.out_ids = out_ids;
    P.err = C.err_dev;
    const int grid = (int)std::min<int64_t>(nq, kExchangeGrid);
    size_t smem = (size_t)nr * k * 8 * (kExchangeThreads / 32);
    if (smem > 96 * 1024) smem = 0;
    P.stage_doubles = smem ? nr * k : 0;
    if (smem > 48 * 1024) {
        HB_CUDA(cudaFuncSetAttribute(exchange_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HB_CUDA(cudaFuncSetAttribute(exchange_merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const bool want_p2p = use_p2p && C.p2p_ok && (nr > 1 || use_p2p == 2);  // 2: also with one rank (single-GPU tests of the kernel)
    if (want_p2p) ensure_windows(block);  // may turn p2p_ok off (IPC not available)
    if (st && st->e0) HB_CUDA(cudaEventRecord(st->e0, g_stream));
    if (want_p2p && C.p2p_ok && C.info.p2p) {
        ++C.epoch;
        const int parity = (int)(C.epoch & 1u);
        for (int r = 0; r < nr; ++r) P.win[r] = (char *)C.win[r];
        P.slot_bytes = C.slot_bytes;
        P.parity_off = (size_t)parity * nr * C.slot_bytes;
        P.flag_off = 2 * (size_t)nr * C.slot_bytes + (size_t)parity * nr * kFlagCtas * 4;
        P.epoch = C.epoch;
        if (st && st->e1) HB_CUDA(cudaEventRecord(st->e1, g_stream));
        exchange_merge_kernel<true><<<grid, kExchangeThreads, smem, g_stream>>>(P);
        HB_LAUNCH_CHECK();
    } else {
        char *pack = (char *)C.pack.get(block);
        char *gath = (char *)C.gathered.get(block * nr);
        const int64_t count = nq * k;
        pack_topk_kernel<<<(int)ceil_div(count, 256), 256, 0, g_stream>>>(loc_dist, loc_ids, count, id_base, (double *)pack,
                                                                          (int64_t *)(pack + (size_t)count * 8));
        HB_LAUNCH_CHECK();
        if (nr > 1) comm_allgather_bytes(pack, gath, (int64_t)block);
        else HB_CUDA(cudaMemcpyAsync(gath, pack, block, cudaMemcpyDeviceToDevice, g_stream));
        if (st && st->e1) HB_CUDA(cudaEventRecord(st->e1, g_stream));
        P.win[P.rank] = gath;
        P.slot_bytes = block;
        P.parity_off = 0;
        exchange_merge_kernel<false><<<grid, kExchangeThreads, smem, g_stream>>>(P);
        HB_LAUNCH_CHECK();
    }
    if (st && st->e2) HB_CUDA(cudaEventRecord(st->e2, g_stream));
}

void comm_check_exchange() {
    CommState &C = g_comm;
    if (!C.info.inited || !C.err_dev) return;
    int32_t h = 0;
    HB_CUDA(cudaMemcpy(&h, C.err_dev, 4, cudaMemcpyDeviceToHost));
    if (h) {
        HB_CUDA(cudaMemset(C.err_dev, 0, 4));
        throw Error(HB_ERR_CUDA, "hb_sharded_search: timed out waiting for a peer rank's top-k (did every rank make the same call?)");
    }
}

void launch_divide_centroids(const double *sums, const int64_t *counts, int nlist, int d, double *cents) {
    const int64_t tot = (int64_t)nlist * d;
    if (tot == 0) return;
    divide_centroids_kernel<<<(int)ceil_div(tot, 256), 256, 0, g_stream>>>(sums, counts, nlist, d, cents);
    HB_LAUNCH_CHECK();
}

void launch_init_centroids_sharded(const void *rows, int dtype, int d, const int64_t *seed_rows, int nlist, int64_t first_row,
                                   int64_t n_local, double *cents) {
    const int64_t tot = (int64_t)nlist * d;
    if (tot == 0) return;
    const int grid = (int)ceil_div(tot, 256);
    if (dtype == HB_F32) init_centroids_sharded_kernel<float><<<grid, 256, 0, g_stream>>>((const float *)rows, d, seed_rows, nlist, first_row, n_local, cents);
    else if (dtype == HB_BF16) init_centroids_sharded_kernel<__nv_bfloat16><<<grid, 256, 0, g_stream>>>((const __nv_bfloat16 *)rows, d, seed_rows, nlist, first_row, n_local, cents);
    else init_centroids_sharded_kernel<double><<<grid, 256, 0, g_stream>>>((const double *)rows, d, seed_rows, nlist, first_row, n_local, cents);
    HB_LAUNCH_CHECK();
}

}  // namespace hb
