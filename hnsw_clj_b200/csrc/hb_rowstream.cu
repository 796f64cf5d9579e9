// hb_rowstream.cu — the HBM-bound flat scan for small batches (1..8 queries against every row of one list: exact flat
// search `compute-exact-knn`, src/hnsw/bench.clj:72-84, and the coarse stage of `search-ivf-flat`,
// src/hnsw/ann/partition/ivf_flat.clj:261-269, when the reference's callers ask one query at a time).
//
// The arithmetic is the reference's: ONE fp64 accumulator per (row, query) pair advanced in index order by one thread
// (mac_seq, hb_common.cuh) — bit-identical to pairscan_kernel / smallscan_kernel.  What differs is how the bytes arrive:
// smallscan_kernel lets every thread stream its own row 128 bytes at a time, i.e. ~75 k concurrent 128-byte streams 3 KB
// apart, which HBM serves at ~60 % of its copy bandwidth (ncu: profiles/r01l_smallscan_*).  Here a WARP owns 32 consecutive
// rows and a ring of shared-memory stages; per stage every lane issues one bulk copy (cp.async.bulk, completion on an
// mbarrier) of SEG contiguous bytes of its row, so DRAM sees 512-byte bursts and the bytes in flight are bounded by
// shared memory, not by registers.  Lanes then read their own row segment back with conflict-free 128-bit loads
// (row pitch SEG + 16 bytes) and the queries by broadcast.  Warps run independently (no CTA barrier in the loop); the
// (row block, segment) steps of a warp form one continuous stream so the ring stays full across row blocks.
#include "hb_kernels.cuh"

namespace hb {
namespace {

__device__ __forceinline__ uint32_t rs_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool rs_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void rs_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
    static constexpr int PER = 4;
    __device__ static __forceinline__ double at(const uint4 &v, int e) {
        return (double)__uint_as_float(e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w);
    }
};
template <>
struct Vec16<__nv_bfloat16> {
    static constexpr int PER = 8;
    __device__ static __forceinline__ double at(const uint4 &v, int e) {
        const uint32_t w = (e >> 1) == 0 ? v.x : (e >> 1) == 1 ? v.y : (e >> 1) == 2 ? v.z : v.w;
        return (double)__uint_as_float((e & 1) ? (w & 0xFFFF0000u) : (w << 16));
    }
};
template <>
struct Vec16<double> {
    static constexpr int PER = 2;
    __device__ static __forceinline__ double at(const uint4 &v, int e) {
        return e == 0 ? __hiloint2double((int)v.y, (int)v.x) : __hiloint2double((int)v.w, (int)v.z);
    }
};

constexpr int kMaxStages = 4;
constexpr int kMaxWarps = 16;

// Dynamic shared memory: [d][NQ] fp64 queries | per warp: stages x 32 rows x (SEG + 16) bytes.
template <typename TRow, typename TQry, int ARITH, int NQ, int SEG>
__global__ void __launch_bounds__(32 * kMaxWarps) rowstream_kernel(const ScanParams P, int warps, int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int PITCH = SEG + 16;
    constexpr int PER = Vec16<TRow>::PER;
    __shared__ __align__(8) unsigned long long s_bar[kMaxWarps][kMaxStages];
    __shared__ int s_qidx[NQ];
    __shared__ long long s_qout[NQ];
    __shared__ double s_qn[NQ];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = P.d;
    double *qs = reinterpret_cast<double *>(smem_raw);
    const size_t qbytes = ((size_t)d * NQ * sizeof(double) + 127) & ~(size_t)127;
    unsigned char *ring = smem_raw + qbytes + (size_t)warp * stages * 32 * PITCH;
    const TRow *rows = static_cast<const TRow *>(P.rows);
    const TQry *queries = static_cast<const TQry *>(P.queries);
    const int64_t row_begin = P.list_off[0];
    const int64_t n = P.list_off[1] - row_begin;
    const int nqt = (int)min((int64_t)NQ, P.lq_off[1] - P.lq_off[0]);

    if (tid < NQ) {
        int qi = -1;
        long long ob = 0;
        double qn = 0.0;
        if (tid < nqt) {
            const int64_t sel0 = P.lq_off[0];
            const int64_t p = P.qsel ? (int64_t)P.qsel[sel0 + tid] : sel0 + tid;
            qi = P.pair_query ? P.pair_query[p] : (P.pair_div > 0 ? (int)(p / P.pair_div) : (int)p);
            ob = P.pair_out ? P.pair_out[p] : p * P.out_stride;
            qn = P.q_norm ? P.q_norm[qi] : 0.0;
        }
        s_qidx[tid] = qi;
        s_qout[tid] = ob;
        s_qn[tid] = qn;
    }
    if (lane == 0)
        for (int s = 0; s < stages; ++s) rs_mbar_init(rs_smem_u32(&s_bar[warp][s]), 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        const int qi = s_qidx[j];
        const TQry *qp = queries + (int64_t)max(qi, 0) * d;
        for (int k = tid; k < d; k += blockDim.x) qs[(int64_t)k * NQ + j] = qi >= 0 ? to_f64(qp[k]) : 0.0;
    }
    __syncthreads();  // the only CTA-wide barriers: from here on every warp runs on its own

    const int row_bytes = d * (int)sizeof(TRow);
    const int nseg = (row_bytes + SEG - 1) / SEG;
    const int64_t nblk = (n + 31) / 32;
    const int64_t gw = (int64_t)blockIdx.x * warps + warp, gstride = (int64_t)gridDim.x * warps;
    const int64_t my_blocks = gw < nblk ? (nblk - gw + gstride - 1) / gstride : 0;
    const int64_t steps = my_blocks * nseg;

    // issue cursor: the next (row block, segment, stage) to request
    int64_t i_blk = gw, i_step = 0;
    int i_seg = 0, i_stage = 0;
    auto issue = [&]() {
        const int64_t r = i_blk * 32 + lane;
        const uint32_t bar = rs_smem_u32(&s_bar[warp][i_stage]);
        if (r < n) {
            const uint32_t bytes = (uint32_t)min(SEG, row_bytes - i_seg * SEG);
            rs_mbar_expect_tx(bar, bytes);
            rs_bulk_g2s(rs_smem_u32(ring + ((size_t)i_stage * 32 + lane) * PITCH),
                        reinterpret_cast<const unsigned char *>(rows + (row_begin + r) * (int64_t)d) + (size_t)i_seg * SEG, bytes, bar);
        } else {
            rs_mbar_arrive(bar);
        }
        ++i_step;
        if (++i_seg == nseg) {
            i_seg = 0;
            i_blk += gstride;
        }
        if (++i_stage == stages) i_stage = 0;
    };

    while (i_step < min((int64_t)stages, steps)) issue();
    double acc[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) acc[j] = 0.0;
    int64_t c_blk = gw;
    int c_seg = 0, c_stage = 0;
    uint32_t c_parity = 0;
    for (int64_t step = 0; step < steps; ++step) {
        const int64_t r = c_blk * 32 + lane;
        while (!rs_mbar_try_wait(rs_smem_u32(&s_bar[warp][c_stage]), c_parity)) {
        }
        if (r < n) {
            const uint4 *mine = reinterpret_cast<const uint4 *>(ring + ((size_t)c_stage * 32 + lane) * PITCH);
            const int bytes = min(SEG, row_bytes - c_seg * SEG);
            const int nvec = bytes >> 4;
            const double *qk = qs + (int64_t)(c_seg * (SEG / (int)sizeof(TRow))) * NQ;
#pragma unroll 4
            for (int v = 0; v < nvec; ++v) {
                const uint4 bits = mine[v];
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    const double x = Vec16<TRow>::at(bits, e);
                    const double *qe = qk + (int64_t)(v * PER + e) * NQ;
#pragma unroll
                    for (int j = 0; j < NQ; ++j) acc[j] = mac_seq<ARITH>(qe[j], x, acc[j]);
                }
            }
            if (c_seg == nseg - 1) {
                const double rn = P.row_norm ? P.row_norm[row_begin + r] : 0.0;
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    if (j < nqt) P.out[s_qout[j] + r] = apply_epi(P.epi, acc[j], s_qn[j], rn);
                    acc[j] = 0.0;
                }
            }
        }
        if (++c_seg == nseg) {
            c_seg = 0;
            c_blk += gstride;
        }
        if (++c_stage == stages) {
            c_stage = 0;
            c_parity ^= 1u;
        }
        __syncwarp();  // every lane has read its segment: the stage may be overwritten
        if (i_step < steps) issue();
    }
}

// ---- the same ring for MANY short lists: the probed lists of a small-batch IVF scan (search-partition,
// src/hnsw/ann/partition/ivf_flat.clj:217-234) --------------------------------------------------------------------------
// ScanParams as pairscan_kernel reads them: tile t of tile_prefix = 128 rows of a list with >= 1 selection.  A warp takes
// quarter tiles (32 rows); for every selection of the list it streams the 32 rows through its ring once (a second
// selection of the same list finds them in L2) and advances one sequential fp64 sum per row against that selection's
// query.  All queries of the batch (<= 8) sit in shared memory as fp64 [q][d].  The issue cursor runs `stages` steps
// ahead of the consume cursor over the same (unit, selection, segment) sequence.
struct ListCursor {
    int64_t unit;     // quarter-tile index (tile * 4 + quarter); >= total: done
    int64_t r0;       // first slab row of the 32-row block
    int64_t sel0;     // first selection of the list
    int64_t blk_off;  // offset of the block inside its list (output position)
    int nrows;        // valid rows in the block (1..32)
    int nsel;         // selections of the list
    int j, seg;       // current selection / segment
};

template <typename TRow, typename TQry, int ARITH, int SEG>
__global__ void __launch_bounds__(32 * kMaxWarps) liststream_kernel(const ScanParams P, int warps, int stages, int nq) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int PITCH = SEG + 16;
    constexpr int PER = Vec16<TRow>::PER;
    __shared__ __align__(8) unsigned long long s_bar[kMaxWarps][kMaxStages];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = P.d;
    double *qs = reinterpret_cast<double *>(smem_raw);  // [nq][d]
    const size_t qbytes = ((size_t)d * nq * sizeof(double) + 127) & ~(size_t)127;
    unsigned char *ring = smem_raw + qbytes + (size_t)warp * stages * 32 * PITCH;
    const TRow *rows = static_cast<const TRow *>(P.rows);
    const TQry *queries = static_cast<const TQry *>(P.queries);

    if (lane == 0)
        for (int s = 0; s < stages; ++s) rs_mbar_init(rs_smem_u32(&s_bar[warp][s]), 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int64_t i = tid; i < (int64_t)nq * d; i += blockDim.x) qs[i] = to_f64(queries[i]);
    __syncthreads();  // the only CTA-wide barrier

    const int row_bytes = d * (int)sizeof(TRow);
    const int nseg = (row_bytes + SEG - 1) / SEG;
    const int64_t total_units = P.tile_prefix[P.nlist] * 4;
    const int64_t gw = (int64_t)blockIdx.x * warps + warp, gstride = (int64_t)gridDim.x * warps;

    // moves the cursor to the next non-empty unit at or after `unit` (warp-uniform: every lane computes the same)
    auto seek = [&](ListCursor &c, int64_t unit) {
        for (;; unit += gstride) {
            c.unit = unit;
            if (unit >= total_units) return;
            const int64_t t = unit >> 2;
            const int quarter = (int)(unit & 3);
            int lo = 0, hi = P.nlist;  // last l with tile_prefix[l] <= t
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (P.tile_prefix[mid] <= t) lo = mid;
                else hi = mid;
            }
            const int64_t l0 = P.list_off[lo], len = P.list_off[lo + 1] - l0;
            c.blk_off = (t - P.tile_prefix[lo]) * 128 + quarter * 32;
            const int64_t left = len - c.blk_off;
            c.sel0 = P.lq_off[lo];
            c.nsel = (int)(P.lq_off[lo + 1] - c.sel0);
            if (left <= 0 || c.nsel <= 0) continue;
            c.nrows = (int)min((int64_t)32, left);
            c.r0 = l0 + c.blk_off;
            c.j = 0;
            c.seg = 0;
            return;
        }
    };
    auto advance = [&](ListCursor &c) {
        if (++c.seg < nseg) return;
        c.seg = 0;
        if (++c.j < c.nsel) return;
        seek(c, c.unit + gstride);
    };

    ListCursor I, C;
    seek(I, gw);
    C = I;
    int i_stage = 0, c_stage = 0;
    uint32_t c_parity = 0;
    auto issue = [&]() {
        const uint32_t bar = rs_smem_u32(&s_bar[warp][i_stage]);
        if (lane < I.nrows) {
            const uint32_t bytes = (uint32_t)min(SEG, row_bytes - I.seg * SEG);
            rs_mbar_expect_tx(bar, bytes);
            rs_bulk_g2s(rs_smem_u32(ring + ((size_t)i_stage * 32 + lane) * PITCH),
                        reinterpret_cast<const unsigned char *>(rows + (I.r0 + lane) * (int64_t)d) + (size_t)I.seg * SEG, bytes, bar);
        } else {
            rs_mbar_arrive(bar);
        }
        if (++i_stage == stages) i_stage = 0;
        advance(I);
    };
    for (int s = 0; s < stages && I.unit < total_units; ++s) issue();

    double acc = 0.0;
    const double *qv = qs;
    int qi = 0;
    int64_t ob = 0;
    while (C.unit < total_units) {
        if (C.seg == 0) {  // a new (block, selection): which query, where the distances go
            const int64_t p = P.qsel ? (int64_t)P.qsel[C.sel0 + C.j] : C.sel0 + C.j;
            qi = P.pair_query ? P.pair_query[p] : (P.pair_div > 0 ? (int)(p / P.pair_div) : (int)p);
            ob = P.pair_out ? P.pair_out[p] : p * P.out_stride;
            qv = qs + (int64_t)qi * d;
            acc = 0.0;
        }
        while (!rs_mbar_try_wait(rs_smem_u32(&s_bar[warp][c_stage]), c_parity)) {
        }
        if (lane < C.nrows) {
            const uint4 *mine = reinterpret_cast<const uint4 *>(ring + ((size_t)c_stage * 32 + lane) * PITCH);
            const int bytes = min(SEG, row_bytes - C.seg * SEG);
            const int nvec = bytes >> 4;
            const double *qk = qv + C.seg * (SEG / (int)sizeof(TRow));
#pragma unroll 4
            for (int v = 0; v < nvec; ++v) {
                const uint4 bits = mine[v];
#pragma unroll
                for (int e = 0; e < PER; ++e) acc = mac_seq<ARITH>(qk[v * PER + e], Vec16<TRow>::at(bits, e), acc);
            }
            if (C.seg == nseg - 1) {
                const double rn = P.row_norm ? P.row_norm[C.r0 + lane] : 0.0;
                P.out[ob + C.blk_off + lane] = apply_epi(P.epi, acc, P.q_norm ? P.q_norm[qi] : 0.0, rn);
            }
        }
        if (++c_stage == stages) {
            c_stage = 0;
            c_parity ^= 1u;
        }
        advance(C);
        __syncwarp();  // every lane has read its segment: the stage may be overwritten
        if (I.unit < total_units) issue();
    }
}

// defaults from the sweep in profiles/r01m_rowstream_sweep.txt: 256-byte segments, 2 stages, as many warps as fit (12 at d = 768)
int g_stream_seg = 256, g_stream_stages = 2, g_stream_warps = 0;  // hb_set_option("stream_seg" / "stream_stages" / "stream_warps")

// opt-in shared memory per block of the process's device (one process per GPU: asked once)
int max_optin_smem() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0;
        HB_CUDA(cudaGetDevice(&dev));
        HB_CUDA(cudaDeviceGetAttribute(&cached, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    }
    return cached;
}

template <typename TRow, typename TQry, int ARITH, int NQ, int SEG>
bool rowstream_go(const ScanParams &P) {
    const int max_smem = max_optin_smem();
    const int stages = std::min(std::max(g_stream_stages, 2), kMaxStages);
    const size_t qbytes = ((size_t)P.d * NQ * sizeof(double) + 127) & ~(size_t)127;
    const size_t per_warp = (size_t)stages * 32 * (SEG + 16);
    const size_t budget = (size_t)max_smem - 1024;  // static shared memory of the kernel
    if (qbytes + per_warp > budget) return false;
    int warps = (int)std::min<size_t>(kMaxWarps, (budget - qbytes) / per_warp);
    if (g_stream_warps > 0) warps = std::min(warps, g_stream_warps);
    const size_t smem = qbytes + (size_t)warps * per_warp;
    auto kernel = rowstream_kernel<TRow, TQry, ARITH, NQ, SEG>;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<g_num_sms, warps * 32, smem, g_stream>>>(P, warps, stages);
    HB_LAUNCH_CHECK();
    return true;
}

template <typename TRow, typename TQry, int ARITH, int NQ>
bool rowstream_seg(const ScanParams &P) {
    if (g_stream_seg == 256) return rowstream_go<TRow, TQry, ARITH, NQ, 256>(P);
    if (g_stream_seg == 384) return rowstream_go<TRow, TQry, ARITH, NQ, 384>(P);
    return rowstream_go<TRow, TQry, ARITH, NQ, 512>(P);
}

template <typename TRow, typename TQry, int ARITH>
bool rowstream_nq(const ScanParams &P, int max_sel) {
    if (max_sel <= 1) return rowstream_seg<TRow, TQry, ARITH, 1>(P);
    if (max_sel <= 4) return rowstream_seg<TRow, TQry, ARITH, 4>(P);
    return rowstream_seg<TRow, TQry, ARITH, 8>(P);
}

template <typename TRow, typename TQry>
bool rowstream_arith(const ScanParams &P, bool l2, int max_sel) {
    if (l2) return rowstream_nq<TRow, TQry, ARITH_L2>(P, max_sel);
    if (is_f32_repr<TRow>::value && is_f32_repr<TQry>::value) return rowstream_nq<TRow, TQry, ARITH_FMA>(P, max_sel);
    return rowstream_nq<TRow, TQry, ARITH_MULADD>(P, max_sel);
}

template <typename TRow, typename TQry, int ARITH, int SEG>
bool liststream_go(const ScanParams &P, int nq) {
    const int max_smem = max_optin_smem();
    const int stages = std::min(std::max(g_stream_stages, 2), kMaxStages);
    const size_t qbytes = ((size_t)P.d * nq * sizeof(double) + 127) & ~(size_t)127;
    const size_t per_warp = (size_t)stages * 32 * (SEG + 16);
    const size_t budget = (size_t)max_smem - 1024;
    if (qbytes + per_warp > budget) return false;
    int warps = (int)std::min<size_t>(kMaxWarps, (budget - qbytes) / per_warp);
    if (g_stream_warps > 0) warps = std::min(warps, g_stream_warps);
    const size_t smem = qbytes + (size_t)warps * per_warp;
    auto kernel = liststream_kernel<TRow, TQry, ARITH, SEG>;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<g_num_sms, warps * 32, smem, g_stream>>>(P, warps, stages, nq);
    HB_LAUNCH_CHECK();
    return true;
}

template <typename TRow, typename TQry>
bool liststream_arith(const ScanParams &P, bool l2, int nq) {
    if (l2) return liststream_go<TRow, TQry, ARITH_L2, 256>(P, nq);
    if (is_f32_repr<TRow>::value && is_f32_repr<TQry>::value) return liststream_go<TRow, TQry, ARITH_FMA, 256>(P, nq);
    return liststream_go<TRow, TQry, ARITH_MULADD, 256>(P, nq);
}

}  // namespace

// Many lists (P.nlist > 1), a batch of nq <= kSmallScanQ queries that P.queries holds in rows 0..nq-1.
bool launch_liststream(const ScanParams &P, int rdtype, int qdtype, bool l2, int nq) {
    if (nq < 1 || nq > kSmallScanQ) return false;
    if (rdtype == HB_F32 && qdtype == HB_F32) return liststream_arith<float, float>(P, l2, nq);
    if (rdtype == HB_F32 && qdtype == HB_F64) return liststream_arith<float, double>(P, l2, nq);
    if (rdtype == HB_BF16 && qdtype == HB_F32) return liststream_arith<__nv_bfloat16, float>(P, l2, nq);
    if (rdtype == HB_BF16 && qdtype == HB_F64) return liststream_arith<__nv_bfloat16, double>(P, l2, nq);
    if (rdtype == HB_F64 && qdtype == HB_F32) return liststream_arith<double, float>(P, l2, nq);
    if (rdtype == HB_F64 && qdtype == HB_F64) return liststream_arith<double, double>(P, l2, nq);
    return false;
}

void set_rowstream_option(const char *name, int value) {
    const std::string n(name);
    if (n == "stream_seg") {
        HB_REQUIRE(value == 256 || value == 384 || value == 512, "stream_seg must be 256, 384 or 512");
        g_stream_seg = value;
    } else if (n == "stream_stages") {
        HB_REQUIRE(value >= 2 && value <= kMaxStages, "stream_stages must be 2..4");
        g_stream_stages = value;
    } else {
        HB_REQUIRE(value >= 0 && value <= kMaxWarps, "stream_warps must be 0..16");
        g_stream_warps = value;
    }
}

// One list (P.nlist == 1) with at most kSmallScanQ selections, 16-byte aligned rows of a multiple of 16 bytes.
// Returns false when the configuration does not fit (the caller falls back to smallscan_kernel).
bool launch_rowstream(const ScanParams &P, int rdtype, int qdtype, bool l2, int max_sel) {
    if (P.nlist != 1 || max_sel < 1 || max_sel > kSmallScanQ) return false;
    if (rdtype == HB_F32 && qdtype == HB_F32) return rowstream_arith<float, float>(P, l2, max_sel);
    if (rdtype == HB_F32 && qdtype == HB_F64) return rowstream_arith<float, double>(P, l2, max_sel);
    if (rdtype == HB_BF16 && qdtype == HB_F32) return rowstream_arith<__nv_bfloat16, float>(P, l2, max_sel);
    if (rdtype == HB_BF16 && qdtype == HB_F64) return rowstream_arith<__nv_bfloat16, double>(P, l2, max_sel);
    if (rdtype == HB_F64 && qdtype == HB_F32) return rowstream_arith<double, float>(P, l2, max_sel);
    if (rdtype == HB_F64 && qdtype == HB_F64) return rowstream_arith<double, double>(P, l2, max_sel);
    return false;
}

}  // namespace hb
