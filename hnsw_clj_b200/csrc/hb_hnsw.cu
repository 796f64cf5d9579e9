// hb_hnsw.cu — batched HNSW search on the device (sm_100a): one warp walks one query through the layers.
//
// Restates search-knn / search-layer-ultra of the reference (src/hnsw/ultra_fast.clj:151-212, 346-374) for many
// concurrent queries.  What the reference does per query on one JVM thread — poll the closest candidate, look up
// its neighbour set, score every unvisited neighbour with distance-fn, push the admitted ones onto two
// java.util.PriorityQueues — a warp does here:
//   * the neighbour-candidate scoring (:185-204, the hot 70-95 % of the reference's query time) is a gather-dot
//     over the neighbour id list: the <= 32 unvisited rows are streamed through shared memory with full-line
//     128-bit loads (each row is 3 KB somewhere in HBM: every line is prefetched into L2 up front, the chunks are
//     software-pipelined through registers), and lane s advances ONE sequential fp64 accumulator over row s in
//     index order, so every distance has the reference's bits;
//   * both priority queues live in shared memory and are driven by lane 0 with java.util.PriorityQueue's exact
//     siftUp / siftDown, in the neighbour set's iteration order, so admissions, evictions and the heap-array order
//     the reference leaks into its result (:207-212) are reproduced, ties included;
//   * the visited set (:189-190, marked BEFORE scoring) is a per-warp bitmap in HBM, cleared through the list of
//     ids that were set.
// Quirks kept: expansion does not stop at the first candidate that fails `current-dist <= worst`
// (it keeps polling), upper layers run with ef = 1, layer 0 with ef = max(k, 50) unless the caller passes ef,
// the final top-k is a stable sort of the `nearest` heap array by distance.
// The candidate queue is unbounded in the reference; here it has `cand_cap` slots in shared memory and a query that
// would overflow them is flagged and re-run by the same kernel with its queue in global memory (capacity n).
#include <float.h>

#include "hb_hnsw.cuh"

namespace hb {
namespace {

constexpr int HN_WARPS = 4;
constexpr int HN_THREADS = HN_WARPS * 32;
constexpr int STAGE_ROW = 128 + 16;  // bytes per staged row chunk (+16: conflict-free 128-bit reads per quarter warp)

struct Heap {  // java.util.PriorityQueue over (dist, id); REV = Collections.reverseOrder (the `nearest` queue)
    double *d;
    int32_t *id;
    int n;
};

template <bool REV>
__device__ __forceinline__ bool before(double x, double y) { return REV ? (y < x) : (x < y); }  // compare(x, y) < 0

// PriorityQueue.offer -> siftUp: stop as soon as compare(x, parent) >= 0
template <bool REV>
__device__ __forceinline__ void heap_offer(Heap &h, double xd, int32_t xi) {
    int k = h.n++;
    while (k > 0) {
        const int parent = (k - 1) >> 1;
        const double pd = h.d[parent];
        if (!before<REV>(xd, pd)) break;
        h.d[k] = pd;
        h.id[k] = h.id[parent];
        k = parent;
    }
    h.d[k] = xd;
    h.id[k] = xi;
}
// PriorityQueue.poll -> siftDown of the last element: child = right if compare(left, right) > 0; stop if
// compare(x, child) <= 0
template <bool REV>
__device__ __forceinline__ void heap_poll(Heap &h, double &rd, int32_t &ri) {
    rd = h.d[0];
    ri = h.id[0];
    const int n = --h.n;
    if (n > 0) {
        const double xd = h.d[n];
        const int32_t xi = h.id[n];
        int k = 0;
        const int half = n >> 1;
        while (k < half) {
            int child = 2 * k + 1;
            double cd = h.d[child];
            const int right = child + 1;
            if (right < n) {
                const double rdv = h.d[right];
                if (before<REV>(rdv, cd)) {  // compare(left, right) > 0
                    child = right;
                    cd = rdv;
                }
            }
            if (!before<REV>(cd, xd)) break;  // compare(x, child) <= 0
            h.d[k] = cd;
            h.id[k] = h.id[child];
            k = child;
        }
        h.d[k] = xd;
        h.id[k] = xi;
    }
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// distance-fn(query, row[my_row]) for the m rows held by lanes 0..m-1 (my_row < 0 on the others): the gather-dot.
// Row chunks (128 bytes of each of the <= 32 rows) are staged through shared memory with full-line loads; the matching
// query chunk is loaded once per warp (one coalesced load, software-pipelined with the rows), widened to fp64 once and
// broadcast from shared memory -- a per-element global load of the query inside the sequential sum left the warp waiting
// on its scoreboard for half of the kernel (ncu, profiles/r01f_hnsw_1m_*).
template <typename TRow, typename TQry, int ARITH>
__device__ __forceinline__ double warp_gather_dot(const TRow *__restrict__ rows, int d, bool vec, const TQry *__restrict__ qp,
                                                  int my_row, int m, unsigned char *stage, double *qstage, int lane, int pf_window) {
    constexpr int CH = 128 / (int)sizeof(TRow);  // elements per 128-byte chunk
    constexpr int PER = 16 / (int)sizeof(TRow);  // elements per 16-byte part
    constexpr int QPL = (CH + 31) / 32;          // query elements a lane stages per chunk
    const int nch = (d + CH - 1) / CH;
    const int nr = (m + 3) >> 2;  // staging rounds: 4 rows per round, 8 lanes x 16 B per row chunk
    const int sub = lane >> 3, part = lane & 7;
    // Rolling L2 prefetch of the lane's own row, pf_window lines ahead of the chunk being staged.  Prefetching whole rows
    // up front put 32 rows x 3 KB per warp in flight -- 227 MB over the 2368 resident warps, more than the 126 MB L2, so
    // lines were evicted before use and fetched twice (ncu: 416 GB of DRAM reads for 262 GB of rows at configs[4]).
    const char *line = reinterpret_cast<const char *>(rows + (int64_t)max(my_row, 0) * d);
    const int nline = my_row >= 0 ? (int)(((size_t)d * sizeof(TRow) + 127) >> 7) : 0;
    for (int i = 0; i < min(pf_window, nline); ++i) prefetch_l2(line + (size_t)i * 128);
    // the rows this lane helps to stage: round r -> slot r*4 + sub
    int64_t src_row[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int rj = __shfl_sync(0xffffffffu, my_row, (r * 4 + sub) & 31);
        src_row[r] = (r < nr && rj >= 0) ? (int64_t)rj : -1;
    }
    auto load_chunk = [&](int c, uint4(&buf)[8], TQry(&qbuf)[QPL]) {
        const int e0 = c * CH + part * PER;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint4 val = make_uint4(0, 0, 0, 0);
            if (src_row[r] >= 0) {
                const TRow *src = rows + src_row[r] * d + e0;
                if (vec && e0 + PER <= d) val = __ldg(reinterpret_cast<const uint4 *>(src));
                else {
                    alignas(16) TRow tmp[PER];
#pragma unroll
                    for (int e = 0; e < PER; ++e) tmp[e] = (e0 + e < d) ? src[e] : TRow(0.0f);
                    val = *reinterpret_cast<uint4 *>(tmp);
                }
            }
            buf[r] = val;
        }
#pragma unroll
        for (int t = 0; t < QPL; ++t) {
            const int i = t * 32 + lane;
            qbuf[t] = (i < CH && c * CH + i < d) ? __ldg(qp + c * CH + i) : TQry(0);
        }
    };
    uint4 nxt[8];
    TQry qnxt[QPL];
    load_chunk(0, nxt, qnxt);
    double s = 0.0;
    for (int c = 0; c < nch; ++c) {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < nr) *reinterpret_cast<uint4 *>(stage + (r * 4 + sub) * STAGE_ROW + part * 16) = nxt[r];
#pragma unroll
        for (int t = 0; t < QPL; ++t)
            if (t * 32 + lane < CH) qstage[t * 32 + lane] = to_f64(qnxt[t]);
        __syncwarp();
        if (c + 1 < nch) load_chunk(c + 1, nxt, qnxt);
        if (c + pf_window < nline) prefetch_l2(line + (size_t)(c + pf_window) * 128);
        if (lane < m) {
            const uint4 *mine = reinterpret_cast<const uint4 *>(stage + lane * STAGE_ROW);
            const int kmax = min(CH, d - c * CH);
            if (kmax == CH) {  // full chunk: no per-element predicates in the sequential sum
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    const uint4 bits = mine[w];
                    const TRow *el = reinterpret_cast<const TRow *>(&bits);
#pragma unroll
                    for (int e = 0; e < PER; e += 2) {
                        const double2 qq = *reinterpret_cast<const double2 *>(qstage + w * PER + e);  // warp-uniform: broadcast
                        s = mac_seq<ARITH>(qq.x, to_f64(el[e]), s);
                        s = mac_seq<ARITH>(qq.y, to_f64(el[e + 1]), s);
                    }
                }
            } else {
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    const uint4 bits = mine[w];
                    const TRow *el = reinterpret_cast<const TRow *>(&bits);
#pragma unroll
                    for (int e = 0; e < PER; ++e)
                        if (w * PER + e < kmax) s = mac_seq<ARITH>(qstage[w * PER + e], to_f64(el[e]), s);
                }
            }
        }
    }
    return s;
}

template <typename TRow, typename TQry, int ARITH>
__global__ void __launch_bounds__(HN_THREADS, 4) hnsw_search_kernel(const HnswSearchParams P) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int efc = P.ef_cap;  // slots of the nearest queue / entry list (>= ef + 1)
    unsigned char *wb = smem_dyn + (size_t)warp * P.warp_smem;
    unsigned char *stage = wb;
    double *qstage = reinterpret_cast<double *>(wb + 32 * STAGE_ROW);  // the query chunk in fp64, 64 slots
    double *near_d = qstage + 64;
    double *sd = near_d + efc;
    double *cand_d_s = sd + 32;
    int32_t *near_i = reinterpret_cast<int32_t *>(cand_d_s + P.cand_cap_smem);
    int32_t *eps = near_i + efc;
    int32_t *sid = eps + efc;
    int32_t *cand_i_s = sid + 32;

    const int slot = blockIdx.x * HN_WARPS + warp;
    uint32_t *vis = P.visited + (int64_t)slot * P.vwords;
    int32_t *vlist = P.vlist + (int64_t)slot * P.vcap;
    const bool global_cand = P.g_cand_d != nullptr;
    const int cand_cap = global_cand ? P.g_cand_cap : P.cand_cap_smem;
    double *cand_d = global_cand ? P.g_cand_d + (int64_t)slot * P.g_cand_cap : cand_d_s;
    int32_t *cand_i = global_cand ? P.g_cand_i + (int64_t)slot * P.g_cand_cap : cand_i_s;

    const TRow *rows = reinterpret_cast<const TRow *>(P.rows);
    const int d = P.d;
    const bool vec = (((size_t)d * sizeof(TRow)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0);

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(P.next_work, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= P.nwork) break;
        const int q = P.work_list ? P.work_list[w] : w;
        const TQry *qp = reinterpret_cast<const TQry *>(P.queries) + (int64_t)q * d;
        const double qn = P.q_norm ? P.q_norm[q] : 0.0;

        int neps = 1;
        if (lane == 0) eps[0] = P.entry;
        __syncwarp();
        int overflow = 0;
        unsigned long long scored = 0;
        Heap cand{cand_d, cand_i, 0}, nearq{near_d, near_i, 0};

        for (int level = P.max_level; level >= 0 && !overflow; --level) {
            const int nc = level > 0 ? 1 : P.ef;  // num-closest (:372-374)
            const int64_t *__restrict__ aoff = P.adj_off[level];
            const int32_t *__restrict__ aids = P.adj_ids[level];
            int vcount = 0;
            cand.n = 0;
            nearq.n = 0;
            // ---- entry points (:162-167): score, mark visited, offer to both queues ----
            for (int base = 0; base < neps; base += 32) {
                const int m = min(32, neps - base);
                const int id = lane < m ? eps[base + lane] : -1;
                if (id >= 0) {
                    atomicOr(&vis[id >> 5], 1u << (id & 31));
                    const int p = vcount + lane;
                    if (p < P.vcap) vlist[p] = id;
                }
                vcount += m;
                const double acc = warp_gather_dot<TRow, TQry, ARITH>(rows, d, vec, qp, id, m, stage, qstage, lane, P.pf_window);
                if (lane < m) {
                    sd[lane] = apply_epi(P.epi, acc, qn, P.row_norm ? P.row_norm[id] : 0.0);
                    sid[lane] = id;
                }
                scored += (unsigned)m;
                __syncwarp();
                if (lane == 0) {
                    for (int s = 0; s < m; ++s) {
                        if (cand.n >= cand_cap || nearq.n >= efc) {
                            overflow = 1;
                            break;
                        }
                        heap_offer<false>(cand, sd[s], sid[s]);
                        heap_offer<true>(nearq, sd[s], sid[s]);
                    }
                }
                cand.n = __shfl_sync(0xffffffffu, cand.n, 0);
                nearq.n = __shfl_sync(0xffffffffu, nearq.n, 0);
                overflow = __shfl_sync(0xffffffffu, overflow, 0);
                __syncwarp();
            }
            // ---- expansion loop (:170-204) ----
            while (!overflow) {
                int cur = -1;
                if (lane == 0) {
                    while (cand.n > 0) {
                        double cd;
                        int32_t ci;
                        heap_poll<false>(cand, cd, ci);
                        // :175-178 — a candidate farther than the current worst is skipped, NOT a loop exit
                        if (nearq.n < nc || cd <= (nearq.n == 0 ? DBL_MAX : near_d[0])) {
                            cur = ci;
                            break;
                        }
                    }
                }
                cur = __shfl_sync(0xffffffffu, cur, 0);
                cand.n = __shfl_sync(0xffffffffu, cand.n, 0);
                if (cur < 0) break;
                const int64_t a0 = aoff[cur], a1 = aoff[cur + 1];
                for (int64_t base = a0; base < a1 && !overflow; base += 32) {
                    const int nb = (base + lane < a1) ? aids[base + lane] : -1;
                    bool fresh = false;
                    if (nb >= 0) {  // :189-190 visited is marked before scoring
                        const uint32_t bit = 1u << (nb & 31);
                        fresh = !(atomicOr(&vis[nb >> 5], bit) & bit);
                    }
                    const unsigned fm = __ballot_sync(0xffffffffu, fresh);
                    if (!fm) continue;
                    const int m = __popc(fm);
                    const int rank = __popc(fm & ((1u << lane) - 1u));
                    if (fresh) {
                        const int p = vcount + rank;
                        if (p < P.vcap) vlist[p] = nb;
                        sid[rank] = nb;  // compaction keeps the neighbour set's iteration order
                    }
                    vcount += m;
                    __syncwarp();
                    const int my = lane < m ? sid[lane] : -1;
                    const double acc = warp_gather_dot<TRow, TQry, ARITH>(rows, d, vec, qp, my, m, stage, qstage, lane, P.pf_window);
                    scored += (unsigned)m;
                    double dist = 0.0;
                    if (lane < m) dist = apply_epi(P.epi, acc, qn, P.row_norm ? P.row_norm[my] : 0.0);
                    // The worst of a full `nearest` only decreases while this batch is inserted, so a neighbour that
                    // fails `dist < worst` now fails it later too: only the others go through lane 0.
                    const double worst0 = nearq.n > 0 ? near_d[0] : DBL_MAX;
                    const bool pass = lane < m && (nearq.n < nc || dist < worst0);
                    unsigned pm = __ballot_sync(0xffffffffu, pass);
                    if (!pm) continue;
                    sd[lane] = dist;
                    __syncwarp();
                    if (lane == 0) {
                        while (pm) {
                            const int s = __ffs(pm) - 1;
                            pm &= pm - 1;
                            const double ds = sd[s];
                            if (nearq.n < nc || ds < (nearq.n == 0 ? DBL_MAX : near_d[0])) {  // :195-198
                                if (cand.n >= cand_cap) {
                                    overflow = 1;
                                    break;
                                }
                                const int32_t is = sid[s];
                                heap_offer<false>(cand, ds, is);
                                heap_offer<true>(nearq, ds, is);
                                if (nearq.n > nc) {
                                    double td;
                                    int32_t ti;
                                    heap_poll<true>(nearq, td, ti);
                                }
                            }
                        }
                    }
                    cand.n = __shfl_sync(0xffffffffu, cand.n, 0);
                    nearq.n = __shfl_sync(0xffffffffu, nearq.n, 0);
                    overflow = __shfl_sync(0xffffffffu, overflow, 0);
                    __syncwarp();
                }
            }
            // ---- `nearest` in heap-array order becomes the next layer's entry list (:207-212) ----
            __syncwarp();
            neps = nearq.n;
            for (int i = lane; i < neps; i += 32) eps[i] = near_i[i];
            // ---- clear the visited bits of this layer ----
            if (vcount <= P.vcap) {
                for (int i = lane; i < vcount; i += 32) vis[vlist[i] >> 5] = 0u;
            } else {
                for (int64_t i = lane; i < P.vwords; i += 32) vis[i] = 0u;
            }
            __syncwarp();
        }

        // ---- :364-370 re-score (same bits as the queue's distances), stable sort-by :distance, take k ----
        int64_t *oi = P.out_ids + (int64_t)q * P.k;
        double *od = P.out_dist + (int64_t)q * P.k;
        for (int j = lane; j < P.k; j += 32) {
            oi[j] = -1;
            od[j] = INFINITY;
        }
        __syncwarp();
        if (!overflow) {
            const int nn = nearq.n;
            for (int i = lane; i < nn; i += 32) {
                const double di = near_d[i];
                int rank = 0;
                for (int j = 0; j < nn; ++j) {
                    const double dj = near_d[j];
                    rank += (dj < di || (dj == di && j < i)) ? 1 : 0;
                }
                if (rank < P.k) {
                    oi[rank] = near_i[i];
                    od[rank] = di;
                }
            }
        } else if (lane == 0) {
            const int p = atomicAdd(P.n_overflow, 1);
            P.overflow_list[p] = q;
        }
        if (lane == 0 && P.n_scored) atomicAdd(P.n_scored, scored);
        __syncwarp();
    }
}

template <typename TRow, typename TQry, int ARITH>
void hnsw_go(const HnswSearchParams &P, int grid, size_t smem) {
    auto kernel = hnsw_search_kernel<TRow, TQry, ARITH>;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, HN_THREADS, smem, g_stream>>>(P);
    HB_LAUNCH_CHECK();
}
template <typename TRow, typename TQry>
void hnsw_arith(const HnswSearchParams &P, bool l2, int grid, size_t smem) {
    if (l2) hnsw_go<TRow, TQry, ARITH_L2>(P, grid, smem);
    else if (is_f32_repr<TRow>::value && is_f32_repr<TQry>::value) hnsw_go<TRow, TQry, ARITH_FMA>(P, grid, smem);
    else hnsw_go<TRow, TQry, ARITH_MULADD>(P, grid, smem);
}

}  // namespace

int hnsw_warps_per_cta() { return HN_WARPS; }

size_t hnsw_warp_smem(int ef_cap, int cand_cap_smem) {
    size_t b = 32 * STAGE_ROW + 64 * 8;
    b += (size_t)(ef_cap + 32 + cand_cap_smem) * 8;
    b += (size_t)(ef_cap + ef_cap + 32 + cand_cap_smem) * 4;
    return (b + 15) & ~(size_t)15;
}

void launch_hnsw_search(const HnswSearchParams &P, int rdtype, int qdtype, bool l2, int grid) {
    const size_t smem = (size_t)HN_WARPS * P.warp_smem;
    if (rdtype == HB_F32 && qdtype == HB_F32) hnsw_arith<float, float>(P, l2, grid, smem);
    else if (rdtype == HB_F32 && qdtype == HB_F64) hnsw_arith<float, double>(P, l2, grid, smem);
    else if (rdtype == HB_BF16 && qdtype == HB_F32) hnsw_arith<__nv_bfloat16, float>(P, l2, grid, smem);
    else if (rdtype == HB_BF16 && qdtype == HB_F64) hnsw_arith<__nv_bfloat16, double>(P, l2, grid, smem);
    else if (rdtype == HB_F64 && qdtype == HB_F32) hnsw_arith<double, float>(P, l2, grid, smem);
    else if (rdtype == HB_F64 && qdtype == HB_F64) hnsw_arith<double, double>(P, l2, grid, smem);
    else throw Error(HB_ERR_INVALID, "unsupported dtype combination");
}

}  // namespace hb
