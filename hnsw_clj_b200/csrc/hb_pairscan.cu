// hb_pairscan.cu — exact distance kernels of libhnswb200 (sm_100a).
//
// Every (row, query) pair owns ONE fp64 accumulator that is advanced in strict index order
// (acc = acc + v[i]*q[i], i = 0..d-1), which is what the reference's Clojure loops compute
// (src/hnsw/ultra_fast.clj:53-95, src/hnsw/ann/partition/ivf_flat.clj:224-225, src/hnsw/bench.clj:80-83;
// SURVEY Appendix A.1).  Parallelism comes from running many independent pairs at once: a CTA owns a
// 128-row x 64-query tile, a thread an 8x4 register tile, and the k loop walks the dimension in order.
// Operands are staged to shared memory as fp64, k-major, so the inner loop is 6 LDS.128 + 32 DFMA.
// The bound is the fp64 pipe (64 DFMA/clk/SM), not HBM: every row byte is reused by 64 queries.
#include <float.h>
#include <limits.h>

#include "hb_kernels.cuh"

namespace hb {

namespace {

constexpr int TR = kTileRows, TQ = kTileQ, KC = 16, NT = 128, RT = 8, QT = 8;
constexpr int kSmemBytes = 2 * KC * (TR + TQ) * (int)sizeof(double);  // 49,152 B
// Thread (tx = tid & 7, ty = tid >> 3) owns rows  {2*ty + 32*i + {0,1}}, i < 4  and queries {2*tx + 16*j + {0,1}}, j < 4:
// every LDS.128 of a warp then reads one contiguous run of 16-byte chunks (no bank conflicts; identical
// quarter-warps are served by one wavefront), 8 LDS.128 feed 64 DFMA.
__device__ __forceinline__ int row_of(int ty, int i) { return 2 * ty + 32 * (i >> 1) + (i & 1); }
__device__ __forceinline__ int qry_of(int tx, int j) { return 2 * tx + 16 * (j >> 1) + (j & 1); }

template <typename T>
__device__ __forceinline__ T zero_of() { return T(0); }
template <>
__device__ __forceinline__ __nv_bfloat16 zero_of<__nv_bfloat16>() { return __float2bfloat16(0.0f); }

// v[0..N) = p[k0 .. k0+N), zero beyond d or when !valid.  VEC: 16-byte aligned rows.
template <typename T, int N>
struct alignas(16) Run {
    T v[N];
    __device__ __forceinline__ T operator[](int i) const { return v[i]; }
};

template <typename T, int N, bool VEC>
__device__ __forceinline__ void load_run(const T *__restrict__ p, int k0, int d, bool valid, Run<T, N> &run) {
    T(&v)[N] = run.v;
    if (valid && k0 + N <= d) {
        if constexpr (VEC) {
            constexpr int PER = 16 / (int)sizeof(T);
            static_assert(N % PER == 0, "run must be whole 16-byte vectors");
            const uint4 *src = reinterpret_cast<const uint4 *>(p + k0);
            uint4 *dst = reinterpret_cast<uint4 *>(&v[0]);
#pragma unroll
            for (int i = 0; i < N / PER; ++i) dst[i] = __ldg(src + i);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = p[k0 + i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = (valid && k0 + i < d) ? p[k0 + i] : zero_of<T>();
    }
}

template <int ARITH>
__device__ __forceinline__ void tile_mac(const double *__restrict__ rsb, const double *__restrict__ qsb, int kmax,
                                         double (&acc)[RT][QT], int tx, int ty) {
    const double *rp0 = rsb + 2 * ty;
    const double *qp0 = qsb + 2 * tx;
#pragma unroll 2
    for (int k = 0; k < kmax; ++k) {
        double r[RT], q[QT];
#pragma unroll
        for (int i = 0; i < RT / 2; ++i) {
            const double2 t = *reinterpret_cast<const double2 *>(rp0 + k * TR + 32 * i);
            r[2 * i] = t.x;
            r[2 * i + 1] = t.y;
        }
#pragma unroll
        for (int j = 0; j < QT / 2; ++j) {
            const double2 t = *reinterpret_cast<const double2 *>(qp0 + k * TQ + 16 * j);
            q[2 * j] = t.x;
            q[2 * j + 1] = t.y;
        }
#pragma unroll
        for (int i = 0; i < RT; ++i)
#pragma unroll
            for (int j = 0; j < QT; ++j) acc[i][j] = mac_seq<ARITH>(q[j], r[i], acc[i][j]);
    }
}

// Runs the full k loop for one (row tile, query tile): rows rowp (this thread's loader row) and
// queries qp (this thread's loader query), results in acc.
template <typename TRow, typename TQry, int ARITH, bool VEC>
__device__ __forceinline__ void run_tile(const TRow *__restrict__ rowp, bool row_valid, const TQry *__restrict__ qp,
                                         bool q_valid, int d, double *smem, double (&acc)[RT][QT]) {
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int lrow = tid;                                  // loader: one row, the whole 16-element run
    const int lq = tid & (TQ - 1), qk = (tid >> 6) * 8;    // loader: query, k offset (8 elements)
    double *rs = smem;                 // [2][KC][TR]
    double *qs = smem + 2 * KC * TR;   // [2][KC][TQ]
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < QT; ++j) acc[i][j] = 0.0;

    Run<TRow, 16> rv;
    Run<TQry, 8> qv;
    const int nchunks = (d + KC - 1) / KC;
    load_run<TRow, 16, VEC>(rowp, 0, d, row_valid, rv);
    load_run<TQry, 8, VEC>(qp, qk, d, q_valid, qv);
    for (int c = 0; c < nchunks; ++c) {
        double *rsb = rs + (c & 1) * KC * TR;
        double *qsb = qs + (c & 1) * KC * TQ;
#pragma unroll
        for (int i = 0; i < 16; ++i) rsb[i * TR + lrow] = to_f64(rv[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) qsb[(qk + i) * TQ + lq] = to_f64(qv[i]);
        __syncthreads();
        if (c + 1 < nchunks) {
            load_run<TRow, 16, VEC>(rowp, (c + 1) * KC, d, row_valid, rv);
            load_run<TQry, 8, VEC>(qp, (c + 1) * KC + qk, d, q_valid, qv);
        }
        const int kmax = min(KC, d - c * KC);
        if (kmax == KC) tile_mac<ARITH>(rsb, qsb, KC, acc, tx, ty);
        else tile_mac<ARITH>(rsb, qsb, kmax, acc, tx, ty);
        // the next iteration writes the other buffer; its barrier orders this iteration's reads of (c&1)
        // against the writes two iterations later
    }
    __syncthreads();
}

template <typename TRow, typename TQry, int ARITH, bool VEC>
__global__ void __launch_bounds__(NT, 2) pairscan_kernel(const ScanParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    __shared__ int s_tile[4];
    __shared__ long long s_row0;
    __shared__ int s_qidx[TQ];
    __shared__ long long s_qout[TQ];
    __shared__ double s_qn[TQ];
    __shared__ double s_rn[TR];

    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int64_t total = P.tile_prefix[P.nlist];
    const TRow *rows = static_cast<const TRow *>(P.rows);
    const TQry *queries = static_cast<const TQry *>(P.queries);

    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        if (tid == 0) {
            int lo = 0, hi = P.nlist;  // last l with tile_prefix[l] <= t
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (P.tile_prefix[mid] <= t) lo = mid;
                else hi = mid;
            }
            const int l = lo;
            const int64_t local = t - P.tile_prefix[l];
            const int64_t len = P.list_off[l + 1] - P.list_off[l];
            const int64_t nrt = (len + TR - 1) / TR;
            const int64_t qt = local / nrt, rt = local - qt * nrt;
            s_tile[0] = l;
            s_tile[1] = (int)rt;
            s_tile[2] = (int)min((int64_t)TR, len - rt * TR);                                      // rows in tile
            s_tile[3] = (int)min((int64_t)TQ, P.lq_off[l + 1] - P.lq_off[l] - qt * TQ);            // queries
            s_row0 = P.list_off[l] + rt * TR;
            s_qout[0] = P.lq_off[l] + qt * TQ;  // temp: first selection (re-written below)
        }
        __syncthreads();
        const int rt = s_tile[1], nrows = s_tile[2], nqt = s_tile[3];
        const int64_t row0 = s_row0;
        const int64_t sel0 = s_qout[0];
        __syncthreads();
        if (tid < TQ) {
            int qi = -1;
            long long ob = 0;
            double qn = 0.0;
            if (tid < nqt) {
                const int64_t p = P.qsel ? (int64_t)P.qsel[sel0 + tid] : sel0 + tid;
                qi = P.pair_query ? P.pair_query[p] : (P.pair_div > 0 ? (int)(p / P.pair_div) : (int)p);
                ob = P.pair_out ? P.pair_out[p] : p * P.out_stride;
                qn = P.q_norm ? P.q_norm[qi] : 0.0;
            }
            s_qidx[tid] = qi;
            s_qout[tid] = ob;
            s_qn[tid] = qn;
        }
        s_rn[tid] = (P.row_norm && tid < nrows) ? P.row_norm[row0 + tid] : 0.0;
        __syncthreads();

        const int lrow = tid, lq = tid & (TQ - 1);
        const bool row_valid = lrow < nrows;
        const int my_q = s_qidx[lq];
        const TRow *rowp = rows + (row0 + (row_valid ? lrow : 0)) * (int64_t)P.d;
        const TQry *qp = queries + (int64_t)(my_q < 0 ? 0 : my_q) * P.d;
        double acc[RT][QT];
        run_tile<TRow, TQry, ARITH, VEC>(rowp, row_valid, qp, my_q >= 0, P.d, smem, acc);

#pragma unroll
        for (int j = 0; j < QT; ++j) {
            const int q = qry_of(tx, j);
            if (q < nqt) {
                const double qn = s_qn[q];
                double *o = P.out + s_qout[q] + (int64_t)rt * TR;
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    const int r = row_of(ty, i);
                    if (r < nrows) o[r] = apply_epi(P.epi, acc[i][j], qn, s_rn[r]);
                }
            }
        }
        __syncthreads();
    }
}

// ---- small-batch scan: lists with at most NQ <= 8 selections (search-knn with one query per call, the reference's own
// calling pattern: src/hnsw/ann/partition/ivf_flat.clj:300-317, src/hnsw/bench.clj:72-84) -------------------------------
// The 128 x 64 tile of pairscan_kernel would spend 63/64 of its DFMAs on padding; this scan is bound by HBM instead: a CTA
// owns 128 consecutive rows of a list, THREAD = ROW streams its own row with 128-bit loads (two 128-byte chunks in flight
// per thread: every sector it touches is consumed whole, no shared-memory staging, no barrier in the loop) and advances
// one sequential fp64 sum per selection -- the same mac_seq chain as pairscan_kernel, so the results are bit-identical.
// The selections' queries sit in shared memory as fp64, [k][NQ], read by broadcast.  Same ScanParams / tile numbering
// as pairscan_kernel (a tile = 128 rows of a list x all of its <= NQ selections).
template <typename TRow, typename TQry, int ARITH, int NQ>
__global__ void __launch_bounds__(TR) smallscan_kernel(const ScanParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *qs = reinterpret_cast<double *>(smem_raw);  // [d][NQ]
    __shared__ int s_tile[3];
    __shared__ long long s_row0, s_sel0;
    __shared__ int s_qidx[NQ];
    __shared__ long long s_qout[NQ];
    __shared__ double s_qn[NQ];
    constexpr int CH = sizeof(TRow) == 2 ? 32 : 128 / (int)sizeof(TRow);  // elements per chunk: 128 bytes (bf16: 64)

    const int tid = threadIdx.x;
    const int d = P.d;
    const int64_t total = P.tile_prefix[P.nlist];
    const TRow *rows = static_cast<const TRow *>(P.rows);
    const TQry *queries = static_cast<const TQry *>(P.queries);
    int loaded_list = -1;  // block-uniform: the list whose queries are in shared memory

    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        __syncthreads();
        if (tid == 0) {
            int lo = 0, hi = P.nlist;  // last l with tile_prefix[l] <= t
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (P.tile_prefix[mid] <= t) lo = mid;
                else hi = mid;
            }
            const int64_t rt = t - P.tile_prefix[lo];
            const int64_t len = P.list_off[lo + 1] - P.list_off[lo];
            s_tile[0] = lo;
            s_tile[1] = (int)min((int64_t)TR, len - rt * TR);
            s_tile[2] = (int)min((int64_t)NQ, P.lq_off[lo + 1] - P.lq_off[lo]);
            s_row0 = P.list_off[lo] + rt * TR;
            s_sel0 = P.lq_off[lo];
        }
        __syncthreads();
        const int l = s_tile[0], nrows = s_tile[1], nqt = s_tile[2];
        const int64_t row0 = s_row0, sel0 = s_sel0;
        const int64_t rt_off = row0 - P.list_off[l];
        if (l != loaded_list) {
            if (tid < NQ) {
                int qi = -1;
                long long ob = 0;
                double qn = 0.0;
                if (tid < nqt) {
                    const int64_t p = P.qsel ? (int64_t)P.qsel[sel0 + tid] : sel0 + tid;
                    qi = P.pair_query ? P.pair_query[p] : (P.pair_div > 0 ? (int)(p / P.pair_div) : (int)p);
                    ob = P.pair_out ? P.pair_out[p] : p * P.out_stride;
                    qn = P.q_norm ? P.q_norm[qi] : 0.0;
                }
                s_qidx[tid] = qi;
                s_qout[tid] = ob;
                s_qn[tid] = qn;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int qi = s_qidx[j];
                const TQry *qp = queries + (int64_t)max(qi, 0) * d;
                for (int k = tid; k < d; k += TR) qs[(int64_t)k * NQ + j] = qi >= 0 ? to_f64(qp[k]) : 0.0;
            }
            loaded_list = l;
            __syncthreads();
        }
        if (tid < nrows) {
            const TRow *rp = rows + (row0 + tid) * (int64_t)d;
            double acc[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) acc[j] = 0.0;
            auto consume = [&](const Run<TRow, CH> &v, int c) {
                const int kbase = c * CH;
                const double *qk = qs + (int64_t)kbase * NQ;
                if (kbase + CH <= d) {
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const double x = to_f64(v[i]);
#pragma unroll
                        for (int j = 0; j < NQ; ++j) acc[j] = mac_seq<ARITH>(qk[i * NQ + j], x, acc[j]);
                    }
                } else {
                    const int kmax = d - kbase;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        if (i < kmax) {
                            const double x = to_f64(v[i]);
#pragma unroll
                            for (int j = 0; j < NQ; ++j) acc[j] = mac_seq<ARITH>(qk[i * NQ + j], x, acc[j]);
                        }
                    }
                }
            };
            const int nch = (d + CH - 1) / CH;
            Run<TRow, CH> a, b;
            load_run<TRow, CH, true>(rp, 0, d, true, a);
            for (int c = 0; c < nch; c += 2) {
                if (c + 1 < nch) load_run<TRow, CH, true>(rp, (c + 1) * CH, d, true, b);
                consume(a, c);
                if (c + 2 < nch) load_run<TRow, CH, true>(rp, (c + 2) * CH, d, true, a);
                if (c + 1 < nch) consume(b, c + 1);
            }
            const double rn = P.row_norm ? P.row_norm[row0 + tid] : 0.0;
#pragma unroll
            for (int j = 0; j < NQ; ++j)
                if (j < nqt) P.out[s_qout[j] + rt_off + tid] = apply_epi(P.epi, acc[j], s_qn[j], rn);
        }
    }
}

// sqrt(sum v^2) for a handful of rows (the queries of a small batch): one warp per row stages it in shared memory with
// coalesced loads, lane 0 walks the sequential sum (one thread per row would wait on d dependent global loads).
template <typename T>
__global__ void __launch_bounds__(32) row_norms_warp_kernel(const T *__restrict__ rows, int d, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *v = reinterpret_cast<double *>(smem_raw);
    const T *p = rows + (int64_t)blockIdx.x * d;
    for (int k = threadIdx.x; k < d; k += 32) v[k] = to_f64(p[k]);
    __syncwarp();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < d; ++k) s = mac_seq<is_f32_repr<T>::value ? ARITH_FMA : ARITH_MULADD>(v[k], v[k], s);
        out[blockIdx.x] = __dsqrt_rn(s);
    }
}

// assign-to-nearest-centroid, src/hnsw/ann/partition/ivf_flat.clj:79-90: centroids scanned in index
// order from Double/MAX_VALUE with strict <, so the lowest index wins ties and NaN never wins.
template <typename TRow, int ARITH, bool VEC>
__global__ void __launch_bounds__(NT, 1) assign_kernel(const AssignParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    __shared__ double s_rn[TR];
    __shared__ double s_cn[TQ];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const TRow *rows = static_cast<const TRow *>(P.rows);
    const int64_t ntiles = (P.n + TR - 1) / TR;
    const int nct = (P.nlist + TQ - 1) / TQ;

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t row0 = t * TR;
        const int nrows = (int)min((int64_t)TR, P.n - row0);
        __syncthreads();
        s_rn[tid] = (P.row_norm && tid < nrows) ? P.row_norm[row0 + tid] : 0.0;
        double best[RT];
        int bidx[RT];
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            best[i] = DBL_MAX;
            bidx[i] = INT_MAX;
        }
        const int lrow = tid, lq = tid & (TQ - 1);
        const bool row_valid = lrow < nrows;
        const TRow *rowp = rows + (row0 + (row_valid ? lrow : 0)) * (int64_t)P.d;
        for (int ct = 0; ct < nct; ++ct) {
            const int c0 = ct * TQ;
            const int ncq = min(TQ, P.nlist - c0);
            __syncthreads();
            if (tid < TQ) s_cn[tid] = (P.cent_norm && tid < ncq) ? P.cent_norm[c0 + tid] : 0.0;
            const bool q_valid = lq < ncq;
            const double *qp = P.cents + (int64_t)(c0 + (q_valid ? lq : 0)) * P.d;
            double acc[RT][QT];
            run_tile<TRow, double, ARITH, VEC>(rowp, row_valid, qp, q_valid, P.d, smem, acc);
            // a thread's queries are not in index order (qry_of); keep "lowest index wins" explicit
#pragma unroll
            for (int j = 0; j < QT; ++j) {
                const int q = qry_of(tx, j);
                if (q < ncq) {
                    const double cn = s_cn[q];
#pragma unroll
                    for (int i = 0; i < RT; ++i) {
                        // distance-fn(vector, centroid): n1 = row, n2 = centroid; products commute
                        const double dist = apply_epi(P.epi, acc[i][j], s_rn[row_of(ty, i)], cn);
                        if (dist < best[i] || (dist == best[i] && c0 + q < bidx[i])) {
                            best[i] = dist;
                            bidx[i] = c0 + q;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < RT; ++i) {
#pragma unroll
            for (int m = 4; m >= 1; m >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, best[i], m);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx[i], m);
                if (od < best[i] || (od == best[i] && oi < bidx[i])) {
                    best[i] = od;
                    bidx[i] = oi;
                }
            }
        }
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const int r = row_of(ty, i);
                if (r < nrows) {
                    P.out_assign[row0 + r] = bidx[i] == INT_MAX ? 0 : bidx[i];
                    if (P.out_best) P.out_best[row0 + r] = best[i];
                }
            }
        }
    }
}

// sqrt(sum v^2), one thread per row, index order (ivf_flat.clj:171-177; simd_optimized.clj:206-216).
// Squares of fp32-representable values are exact in fp64, so DFMA rounds as mul-then-add.
template <typename T, bool VEC>
__global__ void __launch_bounds__(128) row_norms_kernel(const T *__restrict__ rows, int64_t n, int d,
                                                        double *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const T *p = rows + r * (int64_t)d;
    double s = 0.0;
    constexpr int N = 16, B = 4;  // four runs of 16 elements in flight: a batch of queries is 79 CTAs, the loads are the time
    for (int k0 = 0; k0 < d; k0 += N * B) {
        Run<T, N> v[B];
#pragma unroll
        for (int b = 0; b < B; ++b)
            if (k0 + b * N < d) load_run<T, N, VEC>(p, k0 + b * N, d, true, v[b]);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int kmax = min(N, d - (k0 + b * N));  // <= 0 past the end
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (i < kmax) {
                    const double x = to_f64(v[b][i]);
                    s = mac_seq<is_f32_repr<T>::value ? ARITH_FMA : ARITH_MULADD>(x, x, s);
                }
        }
    }
    out[r] = __dsqrt_rn(s);
}

// One thread per (query, row) pair, sequential (ultra_fast.clj:192 batched).
template <typename TRow, typename TQry, int ARITH, bool VEC>
__global__ void __launch_bounds__(128) gather_score_kernel(const TRow *__restrict__ rows, const double *__restrict__ row_norm,
                                                           const TQry *__restrict__ queries,
                                                           const double *__restrict__ q_norm, int d,
                                                           const int32_t *__restrict__ pair_query,
                                                           const int32_t *__restrict__ pair_row, int64_t npairs, int epi,
                                                           double *__restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const int qi = pair_query[p], ri = pair_row[p];
    const TRow *rp = rows + (int64_t)ri * d;
    const TQry *qp = queries + (int64_t)qi * d;
    double s = 0.0;
    constexpr int N = 8;
    for (int k0 = 0; k0 < d; k0 += N) {
        Run<TRow, N> rv;
        Run<TQry, N> qv;
        load_run<TRow, N, VEC>(rp, k0, d, true, rv);
        load_run<TQry, N, VEC>(qp, k0, d, true, qv);
        const int kmax = min(N, d - k0);
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < kmax) s = mac_seq<ARITH>(to_f64(qv[i]), to_f64(rv[i]), s);
    }
    out[p] = apply_epi(epi, s, q_norm ? q_norm[qi] : 0.0, row_norm ? row_norm[ri] : 0.0);
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
bool vec_ok(const void *p, int d, size_t elem) { return aligned16(p) && ((size_t)d * elem) % 16 == 0; }

template <typename K>
int resident_grid(K kernel, int threads, size_t smem) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    return per_sm * g_num_sms;
}

template <typename TRow, typename TQry, int ARITH>
void pairscan_dispatch(const ScanParams &P, bool vec) {
    auto go = [&](auto kernel) {
        HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        const int grid = resident_grid(kernel, NT, kSmemBytes);
        kernel<<<grid, NT, kSmemBytes, g_stream>>>(P);
        HB_LAUNCH_CHECK();
    };
    if (vec) go(pairscan_kernel<TRow, TQry, ARITH, true>);
    else go(pairscan_kernel<TRow, TQry, ARITH, false>);
}

template <typename TRow, typename TQry, int ARITH>
void smallscan_dispatch(const ScanParams &P, int max_sel) {
    auto go = [&](auto kernel, int nq) {
        const size_t smem = (size_t)P.d * nq * sizeof(double);
        HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = resident_grid(kernel, TR, smem);
        kernel<<<grid, TR, smem, g_stream>>>(P);
        HB_LAUNCH_CHECK();
    };
    if (max_sel <= 1) go(smallscan_kernel<TRow, TQry, ARITH, 1>, 1);
    else if (max_sel <= 4) go(smallscan_kernel<TRow, TQry, ARITH, 4>, 4);
    else go(smallscan_kernel<TRow, TQry, ARITH, 8>, 8);
}

template <typename TRow, typename TQry>
void pairscan_arith(const ScanParams &P, bool l2, bool vec, int max_sel) {
    // small batches (every list has at most kSmallScanQ selections): the HBM-bound thread-per-row scan
    if (max_sel >= 1 && max_sel <= kSmallScanQ && vec && (size_t)P.d * 8 * sizeof(double) <= 96 * 1024) {
        if (l2) smallscan_dispatch<TRow, TQry, ARITH_L2>(P, max_sel);
        else if (is_f32_repr<TRow>::value && is_f32_repr<TQry>::value) smallscan_dispatch<TRow, TQry, ARITH_FMA>(P, max_sel);
        else smallscan_dispatch<TRow, TQry, ARITH_MULADD>(P, max_sel);
        return;
    }
    if (l2) pairscan_dispatch<TRow, TQry, ARITH_L2>(P, vec);
    else if (is_f32_repr<TRow>::value && is_f32_repr<TQry>::value) pairscan_dispatch<TRow, TQry, ARITH_FMA>(P, vec);
    else pairscan_dispatch<TRow, TQry, ARITH_MULADD>(P, vec);
}

template <typename TRow, int ARITH>
void assign_dispatch(const AssignParams &P, bool vec) {
    auto go = [&](auto kernel) {
        HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        const int64_t ntiles = ceil_div(P.n, TR);
        const int grid = (int)std::min<int64_t>(ntiles, resident_grid(kernel, NT, kSmemBytes));
        kernel<<<grid, NT, kSmemBytes, g_stream>>>(P);
        HB_LAUNCH_CHECK();
    };
    if (vec) go(assign_kernel<TRow, ARITH, true>);
    else go(assign_kernel<TRow, ARITH, false>);
}

}  // namespace

int g_use_rowstream = 1;

void launch_pairscan(const ScanParams &P, int rdtype, int qdtype, bool l2, int max_sel, int batch_nq) {
    HB_REQUIRE(qdtype == HB_F32 || qdtype == HB_F64, "queries must be fp32 or fp64");
    const bool vec = vec_ok(P.rows, P.d, dtype_size(rdtype)) && vec_ok(P.queries, P.d, dtype_size(qdtype));
    if (g_use_rowstream && vec && P.nlist == 1 && max_sel >= 1 && max_sel <= kSmallScanQ &&
        launch_rowstream(P, rdtype, qdtype, l2, max_sel))
        return;
    if (g_use_rowstream && vec && P.nlist > 1 && batch_nq >= 1 && batch_nq <= kSmallScanQ &&
        launch_liststream(P, rdtype, qdtype, l2, batch_nq))
        return;
    if (rdtype == HB_F32 && qdtype == HB_F32) pairscan_arith<float, float>(P, l2, vec, max_sel);
    else if (rdtype == HB_F32 && qdtype == HB_F64) pairscan_arith<float, double>(P, l2, vec, max_sel);
    else if (rdtype == HB_BF16 && qdtype == HB_F32) pairscan_arith<__nv_bfloat16, float>(P, l2, vec, max_sel);
    else if (rdtype == HB_BF16 && qdtype == HB_F64) pairscan_arith<__nv_bfloat16, double>(P, l2, vec, max_sel);
    else if (rdtype == HB_F64 && qdtype == HB_F32) pairscan_arith<double, float>(P, l2, vec, max_sel);
    else if (rdtype == HB_F64 && qdtype == HB_F64) pairscan_arith<double, double>(P, l2, vec, max_sel);
    else throw Error(HB_ERR_INVALID, "unsupported row dtype");
}

void launch_assign(const AssignParams &P, int rdtype, bool l2) {
    if (P.n == 0) return;
    const bool vec = vec_ok(P.rows, P.d, dtype_size(rdtype)) && vec_ok(P.cents, P.d, 8);
    if (rdtype == HB_F32) {
        if (l2) assign_dispatch<float, ARITH_L2>(P, vec);
        else assign_dispatch<float, ARITH_MULADD>(P, vec);
    } else if (rdtype == HB_BF16) {
        if (l2) assign_dispatch<__nv_bfloat16, ARITH_L2>(P, vec);
        else assign_dispatch<__nv_bfloat16, ARITH_MULADD>(P, vec);
    } else if (rdtype == HB_F64) {
        if (l2) assign_dispatch<double, ARITH_L2>(P, vec);
        else assign_dispatch<double, ARITH_MULADD>(P, vec);
    } else throw Error(HB_ERR_INVALID, "unsupported row dtype");
}

void launch_row_norms(const void *rows, int dtype, int64_t n, int d, double *out) {
    if (n == 0) return;
    if (n <= 64 && (size_t)d * 8 <= 48 * 1024) {  // a small batch of queries: one warp per row (row_norms_warp_kernel)
        const size_t smem = (size_t)d * 8;
        if (dtype == HB_F32) row_norms_warp_kernel<float><<<(int)n, 32, smem, g_stream>>>((const float *)rows, d, out);
        else if (dtype == HB_BF16) row_norms_warp_kernel<__nv_bfloat16><<<(int)n, 32, smem, g_stream>>>((const __nv_bfloat16 *)rows, d, out);
        else if (dtype == HB_F64) row_norms_warp_kernel<double><<<(int)n, 32, smem, g_stream>>>((const double *)rows, d, out);
        else throw Error(HB_ERR_INVALID, "unsupported dtype");
        HB_LAUNCH_CHECK();
        return;
    }
    const int grid = (int)ceil_div(n, 128);
    const bool vec = vec_ok(rows, d, dtype_size(dtype));
#define HB_RN(T)                                                                                  \
    do {                                                                                          \
        if (vec) row_norms_kernel<T, true><<<grid, 128, 0, g_stream>>>((const T *)rows, n, d, out);  \
        else row_norms_kernel<T, false><<<grid, 128, 0, g_stream>>>((const T *)rows, n, d, out);     \
    } while (0)
    if (dtype == HB_F32) HB_RN(float);
    else if (dtype == HB_BF16) HB_RN(__nv_bfloat16);
    else if (dtype == HB_F64) HB_RN(double);
    else throw Error(HB_ERR_INVALID, "unsupported dtype");
#undef HB_RN
    HB_LAUNCH_CHECK();
}

namespace {
template <typename TRow, typename TQry>
void gather_arith(const void *rows, const double *row_norm, const void *queries, const double *q_norm, int d,
                  const int32_t *pq, const int32_t *pr, int64_t npairs, bool l2, int epi, double *out, bool vec) {
    const int grid = (int)ceil_div(npairs, 128);
    auto go = [&](auto kernel) {
        kernel<<<grid, 128, 0, g_stream>>>((const TRow *)rows, row_norm, (const TQry *)queries, q_norm, d, pq, pr,
                                           npairs, epi, out);
        HB_LAUNCH_CHECK();
    };
    constexpr bool exact = is_f32_repr<TRow>::value && is_f32_repr<TQry>::value;
    if (l2) {
        if (vec) go(gather_score_kernel<TRow, TQry, ARITH_L2, true>);
        else go(gather_score_kernel<TRow, TQry, ARITH_L2, false>);
    } else if (exact) {
        if (vec) go(gather_score_kernel<TRow, TQry, ARITH_FMA, true>);
        else go(gather_score_kernel<TRow, TQry, ARITH_FMA, false>);
    } else {
        if (vec) go(gather_score_kernel<TRow, TQry, ARITH_MULADD, true>);
        else go(gather_score_kernel<TRow, TQry, ARITH_MULADD, false>);
    }
}
}  // namespace

void launch_gather_score(const void *rows, int rdtype, const double *row_norm, const void *queries, int qdtype,
                         const double *q_norm, int d, const int32_t *pair_query, const int32_t *pair_row,
                         int64_t npairs, bool l2, int epi, double *out) {
    if (npairs == 0) return;
    HB_REQUIRE(qdtype == HB_F32 || qdtype == HB_F64, "queries must be fp32 or fp64");
    const bool vec = vec_ok(rows, d, dtype_size(rdtype)) && vec_ok(queries, d, dtype_size(qdtype));
#define HB_GS(TR_, TQ_) gather_arith<TR_, TQ_>(rows, row_norm, queries, q_norm, d, pair_query, pair_row, npairs, l2, epi, out, vec)
    if (rdtype == HB_F32 && qdtype == HB_F32) HB_GS(float, float);
    else if (rdtype == HB_F32 && qdtype == HB_F64) HB_GS(float, double);
    else if (rdtype == HB_BF16 && qdtype == HB_F32) HB_GS(__nv_bfloat16, float);
    else if (rdtype == HB_BF16 && qdtype == HB_F64) HB_GS(__nv_bfloat16, double);
    else if (rdtype == HB_F64 && qdtype == HB_F32) HB_GS(double, float);
    else if (rdtype == HB_F64 && qdtype == HB_F64) HB_GS(double, double);
    else throw Error(HB_ERR_INVALID, "unsupported row dtype");
#undef HB_GS
}

}  // namespace hb
