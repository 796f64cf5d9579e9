// hb_build.cu — IVF build (k-means++ / Lloyd update / list layout) and search bookkeeping kernels.
// Reference: src/hnsw/ann/partition/ivf_flat.clj:32-131 (build), :236-294 (search plan).
#include <cub/cub.cuh>

#include "hb_build.cuh"

namespace hb {
namespace {

template <typename T>
__global__ void init_centroids_kernel(const T *__restrict__ rows, int d, const int64_t *__restrict__ seed_rows,
                                      int nlist, double *__restrict__ cents) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nlist * d) return;
    const int c = (int)(i / d), j = (int)(i % d);
    cents[i] = to_f64(rows[seed_rows[c] * d + j]);
}

// One thread per (cluster, dim): the sum over members must be sequential in row order to reproduce
// compute-centroid (ivf_flat.clj:66-77) bit for bit; consecutive threads read consecutive dims, so each
// member row is one coalesced read.
template <typename T>
__global__ void __launch_bounds__(128) update_centroids_kernel(const T *__restrict__ rows, int d,
                                                               const int64_t *__restrict__ list_off,
                                                               const int64_t *__restrict__ list_rows, int nlist,
                                                               double *__restrict__ cents, double *__restrict__ out_sums,
                                                               int64_t *__restrict__ out_counts) {
    const int c = blockIdx.x;  // clusters on grid.x: nlist can exceed the 65535 limit of grid.y
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= d) return;
    const int64_t b = list_off[c], e = list_off[c + 1];
    double s = 0.0;
    int64_t t = b;
    for (; t + 4 <= e; t += 4) {  // independent loads first, then the ordered adds
        const double v0 = to_f64(rows[list_rows[t] * d + j]);
        const double v1 = to_f64(rows[list_rows[t + 1] * d + j]);
        const double v2 = to_f64(rows[list_rows[t + 2] * d + j]);
        const double v3 = to_f64(rows[list_rows[t + 3] * d + j]);
        s = __dadd_rn(s, v0);
        s = __dadd_rn(s, v1);
        s = __dadd_rn(s, v2);
        s = __dadd_rn(s, v3);
    }
    for (; t < e; ++t) s = __dadd_rn(s, to_f64(rows[list_rows[t] * d + j]));
    if (out_sums) {
        out_sums[(int64_t)c * d + j] = s;
        if (j == 0 && out_counts) out_counts[c] = e - b;
    } else if (e > b) {
        cents[(int64_t)c * d + j] = __ddiv_rn(s, (double)(e - b));
    }
}

// distance-fn(x_i, c) per row, one thread per row, sequential; mind = min(mind, dist)
template <typename T, bool L2>
__global__ void __launch_bounds__(128) kpp_update_kernel(const T *__restrict__ rows, int64_t n, int d,
                                                         const double *__restrict__ row_norm,
                                                         const int64_t *__restrict__ pick, double *__restrict__ mind) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = *pick;
    const T *x = rows + i * d;
    const T *cv = rows + c * d;
    double s = 0.0;
    for (int k = 0; k < d; ++k) {
        const double a = to_f64(x[k]), b = to_f64(cv[k]);
        if (L2) s = mac_seq<ARITH_L2>(b, a, s);
        else s = mac_seq<is_f32_repr<T>::value ? ARITH_FMA : ARITH_MULADD>(a, b, s);
    }
    const double dist = L2 ? __dsqrt_rn(s) : apply_epi(EPI_COS_GUARD, s, row_norm[i], row_norm[c]);
    if (dist < mind[i]) mind[i] = dist;
}

// S = sum_i mind_i^2 in row order (ivf_flat.clj:51-52): an ordered fp64 sum has no parallel form that
// is bit-identical, so one thread walks it while the CTA streams the data through shared memory.
// SQ = false: the weights are d_i themselves (Lightning's seeding, lightning.clj:100-106).
constexpr int PFX_NT = 256, PFX_CH = 4096;
template <bool SQ>
__global__ void __launch_bounds__(PFX_NT) kpp_prefix_kernel(const double *__restrict__ mind, int64_t n,
                                                            double *__restrict__ cum, double *__restrict__ total) {
    __shared__ double sq[PFX_CH];
    double run = 0.0;
    for (int64_t base = 0; base < n; base += PFX_CH) {
        const int m = (int)min((int64_t)PFX_CH, n - base);
        for (int j = threadIdx.x; j < m; j += PFX_NT) {
            const double v = mind[base + j];
            sq[j] = SQ ? __dmul_rn(v, v) : v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll 8
            for (int j = 0; j < m; ++j) {
                const double v = sq[j];
                sq[j] = run;  // cum before element j
                run = __dadd_rn(run, v);
            }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += PFX_NT) cum[base + j] = sq[j];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = run;
}

// first i with cum_i + mind_i^2 >= r (ivf_flat.clj:54-58); the running sum is monotone, so the
// predicate flips exactly once.
template <bool SQ>
__global__ void kpp_pick_kernel(const double *__restrict__ mind, const double *__restrict__ cum, int64_t n,
                                const double *__restrict__ total, const double *__restrict__ u,
                                int64_t *__restrict__ pick) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = __dmul_rn(*u, *total);
    const double v = mind[i];
    const bool here = __dadd_rn(cum[i], SQ ? __dmul_rn(v, v) : v) >= r;
    bool prev = false;
    if (i > 0) {
        const double w = mind[i - 1];
        prev = __dadd_rn(cum[i - 1], SQ ? __dmul_rn(w, w) : w) >= r;
    }
    if (here && !prev) *pick = i;
}
__global__ void kpp_prepick_kernel(int64_t n, int64_t *pick) { *pick = n - 1; }
__global__ void kpp_record_kernel(const int64_t *pick, int64_t *out_seed) { *out_seed = *pick; }

template <typename T>
__global__ void gather_rows_kernel(const T *__restrict__ rows, int d, const int64_t *__restrict__ list_rows, int64_t n,
                                   T *__restrict__ slab, const double *__restrict__ norm,
                                   double *__restrict__ slab_norm) {
    const int64_t j = blockIdx.x;
    if (j >= n) return;
    const int64_t src = list_rows[j];
    for (int k = threadIdx.x; k < d; k += blockDim.x) slab[j * d + k] = rows[src * d + k];
    if (threadIdx.x == 0 && norm) slab_norm[j] = norm[src];
}

__global__ void hist_kernel(const int32_t *__restrict__ assign, int64_t n, int nlist, unsigned long long *cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int a = assign[i];
        if (a >= 0 && a < nlist) atomicAdd(&cnt[a], 1ull);
    }
}
__global__ void iota64_kernel(int64_t *p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// ---- search plan ---------------------------------------------------------------------------
__global__ void plan_pairs_kernel(const int64_t *__restrict__ probe_pos, int64_t np, int nlist,
                                  const int64_t *__restrict__ list_off, int32_t *__restrict__ probes,
                                  int64_t *__restrict__ pair_len, int32_t *__restrict__ keys, int32_t *__restrict__ vals) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int64_t l = probe_pos[p];
    const bool ok = l >= 0 && l < nlist;
    probes[p] = ok ? (int32_t)l : -1;
    pair_len[p] = ok ? list_off[l + 1] - list_off[l] : 0;
    keys[p] = ok ? (int32_t)l : nlist;
    vals[p] = (int32_t)p;
}
__global__ void plan_lists_kernel(const int32_t *__restrict__ sorted_keys, int64_t np, int nlist,
                                  const int64_t *__restrict__ list_off, int64_t *__restrict__ lq_off,
                                  int64_t *__restrict__ tiles, int tile_rows, int tile_q) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > nlist) return;
    // lower_bound(sorted_keys, l)
    int64_t lo = 0, hi = np;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < l) lo = mid + 1;
        else hi = mid;
    }
    lq_off[l] = lo;
    if (l < nlist) {
        int64_t lo2 = lo, hi2 = np;
        while (lo2 < hi2) {
            const int64_t mid = (lo2 + hi2) >> 1;
            if (sorted_keys[mid] < l + 1) lo2 = mid + 1;
            else hi2 = mid;
        }
        const int64_t nq = lo2 - lo, len = list_off[l + 1] - list_off[l];
        tiles[l] = ((nq + tile_q - 1) / tile_q) * ((len + tile_rows - 1) / tile_rows);
    } else {
        tiles[l] = 0;
    }
}

__global__ void ivf_resolve_kernel(const int64_t *__restrict__ pos, int64_t nq, int k, int nprobe,
                                   const int32_t *__restrict__ probes, const int64_t *__restrict__ pair_out,
                                   const int64_t *__restrict__ list_off, const int64_t *__restrict__ list_rows,
                                   int64_t *__restrict__ out_ids, int32_t *__restrict__ ok, const int32_t *__restrict__ ok_other) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * k) return;
    const int64_t q = t / k;
    if (ok != nullptr && t == q * k) ok[q] &= ok_other[q];  // (the query's first thread) both stages must have proven their part
    const int64_t ps = pos[t];
    if (ps < 0) {
        out_ids[t] = -1;
        return;
    }
    const int64_t *po = pair_out + q * nprobe;
    const int64_t abs = po[0] + ps;
    int lo = 0, hi = nprobe;  // last r with po[r] <= abs
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (po[mid] <= abs) lo = mid;
        else hi = mid;
    }
    // skip empty lists that share the same offset: the owning probe is the last with po[r] <= abs
    const int l = probes[q * nprobe + lo];
    out_ids[t] = list_rows[list_off[l] + (abs - po[lo])];
}
__global__ void offset_ids_kernel(int64_t *ids, int64_t nq, int64_t run, int64_t stride, int64_t base) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * run) return;
    const int64_t q = t / run, j = t - q * run;
    int64_t *p = ids + q * stride + j;
    if (*p >= 0) *p += base;
}
__global__ void parts_to_query_major_kernel(const double *__restrict__ dist, const int64_t *__restrict__ ids, int nparts,
                                            int64_t nq, int k, double *__restrict__ out_dist,
                                            int64_t *__restrict__ out_ids) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nparts * nq * k) return;
    const int64_t j = t % k, q = (t / k) % nq, p = t / (k * nq);
    const int64_t o = (q * nparts + p) * k + j;
    out_dist[o] = dist[t];
    out_ids[o] = ids[t];
}
__global__ void fill_f64_kernel(double *p, int64_t n, double v) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = v;
}
__global__ void lookup_ids_kernel(const int64_t *__restrict__ pos, int64_t nq, int k, const int64_t *__restrict__ ids_in,
                                  int64_t stride, int64_t *__restrict__ out_ids) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * k) return;
    const int64_t q = t / k;
    out_ids[t] = pos[t] < 0 ? -1 : ids_in[q * stride + pos[t]];
}
__global__ void pos_to_i32_kernel(const int64_t *pos, int64_t n, int32_t *out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (int32_t)pos[t];
}

inline int blocks_for(int64_t n, int bs) { return (int)ceil_div(n > 0 ? n : 1, bs); }

}  // namespace

void launch_init_centroids(const void *rows, int dtype, int d, const int64_t *seed_rows, int nlist, double *cents) {
    const int64_t tot = (int64_t)nlist * d;
    if (tot == 0) return;
    if (dtype == HB_F32) init_centroids_kernel<float><<<blocks_for(tot, 256), 256, 0, g_stream>>>((const float *)rows, d, seed_rows, nlist, cents);
    else if (dtype == HB_BF16) init_centroids_kernel<__nv_bfloat16><<<blocks_for(tot, 256), 256, 0, g_stream>>>((const __nv_bfloat16 *)rows, d, seed_rows, nlist, cents);
    else init_centroids_kernel<double><<<blocks_for(tot, 256), 256, 0, g_stream>>>((const double *)rows, d, seed_rows, nlist, cents);
    HB_LAUNCH_CHECK();
}

void launch_update_centroids(const void *rows, int dtype, int d, const int64_t *list_off, const int64_t *list_rows,
                             int nlist, double *cents, double *out_sums, int64_t *out_counts) {
    if (nlist == 0 || d == 0) return;
    dim3 grid((unsigned)nlist, (unsigned)ceil_div(d, 128));
    if (dtype == HB_F32) update_centroids_kernel<float><<<grid, 128, 0, g_stream>>>((const float *)rows, d, list_off, list_rows, nlist, cents, out_sums, out_counts);
    else if (dtype == HB_BF16) update_centroids_kernel<__nv_bfloat16><<<grid, 128, 0, g_stream>>>((const __nv_bfloat16 *)rows, d, list_off, list_rows, nlist, cents, out_sums, out_counts);
    else update_centroids_kernel<double><<<grid, 128, 0, g_stream>>>((const double *)rows, d, list_off, list_rows, nlist, cents, out_sums, out_counts);
    HB_LAUNCH_CHECK();
}

void launch_kpp_step(const KppParams &P) {
    const int grid = blocks_for(P.n, 128);
#define HB_KPP(T)                                                                                                   \
    do {                                                                                                            \
        if (P.l2) kpp_update_kernel<T, true><<<grid, 128, 0, g_stream>>>((const T *)P.rows, P.n, P.d, P.row_norm, P.pick, P.mind); \
        else kpp_update_kernel<T, false><<<grid, 128, 0, g_stream>>>((const T *)P.rows, P.n, P.d, P.row_norm, P.pick, P.mind);     \
    } while (0)
    if (P.dtype == HB_F32) HB_KPP(float);
    else if (P.dtype == HB_BF16) HB_KPP(__nv_bfloat16);
    else HB_KPP(double);
#undef HB_KPP
    HB_LAUNCH_CHECK();
    if (P.linear) kpp_prefix_kernel<false><<<1, PFX_NT, 0, g_stream>>>(P.mind, P.n, P.cum, P.total);
    else kpp_prefix_kernel<true><<<1, PFX_NT, 0, g_stream>>>(P.mind, P.n, P.cum, P.total);
    HB_LAUNCH_CHECK();
    kpp_prepick_kernel<<<1, 1, 0, g_stream>>>(P.n, P.pick);
    HB_LAUNCH_CHECK();
    if (P.linear) kpp_pick_kernel<false><<<blocks_for(P.n, 256), 256, 0, g_stream>>>(P.mind, P.cum, P.n, P.total, P.u, P.pick);
    else kpp_pick_kernel<true><<<blocks_for(P.n, 256), 256, 0, g_stream>>>(P.mind, P.cum, P.n, P.total, P.u, P.pick);
    HB_LAUNCH_CHECK();
    kpp_record_kernel<<<1, 1, 0, g_stream>>>(P.pick, P.out_seed);
    HB_LAUNCH_CHECK();
}

void launch_gather_rows(const void *rows, int dtype, int d, const int64_t *list_rows, int64_t n, void *slab,
                        const double *norm, double *slab_norm) {
    if (n == 0) return;
    HB_REQUIRE(n < (1ll << 31), "too many rows for one gather launch");
    if (dtype == HB_F32) gather_rows_kernel<float><<<(unsigned)n, 128, 0, g_stream>>>((const float *)rows, d, list_rows, n, (float *)slab, norm, slab_norm);
    else if (dtype == HB_BF16) gather_rows_kernel<__nv_bfloat16><<<(unsigned)n, 128, 0, g_stream>>>((const __nv_bfloat16 *)rows, d, list_rows, n, (__nv_bfloat16 *)slab, norm, slab_norm);
    else gather_rows_kernel<double><<<(unsigned)n, 128, 0, g_stream>>>((const double *)rows, d, list_rows, n, (double *)slab, norm, slab_norm);
    HB_LAUNCH_CHECK();
}

void build_lists(const int32_t *assign, int64_t n, int nlist, int64_t *list_off, int64_t *list_rows, DevBuf &tmp) {
    HB_REQUIRE(n < (1ll << 31), "n must be < 2^31 for one device sort");
    // counts -> exclusive scan -> list_off; stable radix sort of row ids by assignment -> list_rows
    size_t scan_bytes = 0, sort_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int64_t *)nullptr, (int64_t *)nullptr, nlist + 1);
    int end_bit = 1;
    while ((1ll << end_bit) < nlist) ++end_bit;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                    (const int64_t *)nullptr, (int64_t *)nullptr, (int)n, 0, end_bit, g_stream);
    const size_t a_cnt = (size_t)(nlist + 1) * 8, a_keys = (size_t)n * 4, a_vals = (size_t)n * 8;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t total = al(a_cnt) + al(a_keys) + al(a_vals) + al(std::max(scan_bytes, sort_bytes)) + 256;
    char *base = (char *)tmp.get(total);
    int64_t *cnt = (int64_t *)base;
    int32_t *keys_out = (int32_t *)(base + al(a_cnt));
    int64_t *vals_in = (int64_t *)(base + al(a_cnt) + al(a_keys));
    void *cub_tmp = base + al(a_cnt) + al(a_keys) + al(a_vals);
    HB_CUDA(cudaMemsetAsync(cnt, 0, a_cnt, g_stream));
    if (n > 0) {
        hist_kernel<<<blocks_for(n, 256), 256, 0, g_stream>>>(assign, n, nlist, (unsigned long long *)cnt);
        HB_LAUNCH_CHECK();
    }
    cub::DeviceScan::ExclusiveSum(cub_tmp, scan_bytes, cnt, list_off, nlist + 1, g_stream);
    HB_LAUNCH_CHECK();
    if (n > 0) {
        iota64_kernel<<<blocks_for(n, 256), 256, 0, g_stream>>>(vals_in, n);
        HB_LAUNCH_CHECK();
        cub::DeviceRadixSort::SortPairs(cub_tmp, sort_bytes, assign, keys_out, vals_in, list_rows, (int)n, 0, end_bit,
                                        g_stream);
        HB_LAUNCH_CHECK();
    }
}

void ivf_plan(const int64_t *probe_pos, int64_t np, int nlist, const int64_t *list_off, int32_t *probes,
              int64_t *pair_out, int32_t *qsel, int64_t *lq_off, int64_t *tile_prefix, int tile_rows, int tile_q,
              DevBuf &tmp) {
    HB_REQUIRE(np < (1ll << 31), "too many (query, probe) pairs for one plan");
    size_t scan_bytes = 0, scan2_bytes = 0, sort_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int64_t *)nullptr, (int64_t *)nullptr, (int)np + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, scan2_bytes, (int64_t *)nullptr, (int64_t *)nullptr, nlist + 1);
    int end_bit = 1;
    while ((1ll << end_bit) < nlist + 1) ++end_bit;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, (int)np, 0, end_bit, g_stream);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t a_len = (size_t)(np + 1) * 8, a_k = (size_t)np * 4, a_tiles = (size_t)(nlist + 1) * 8;
    const size_t cubb = std::max(std::max(scan_bytes, scan2_bytes), sort_bytes);
    char *base = (char *)tmp.get(al(a_len) + 3 * al(a_k) + al(a_tiles) + al(cubb) + 256);
    int64_t *pair_len = (int64_t *)base;
    int32_t *keys = (int32_t *)(base + al(a_len));
    int32_t *vals = (int32_t *)(base + al(a_len) + al(a_k));
    int32_t *keys_sorted = (int32_t *)(base + al(a_len) + 2 * al(a_k));
    int64_t *tiles = (int64_t *)(base + al(a_len) + 3 * al(a_k));
    void *cub_tmp = base + al(a_len) + 3 * al(a_k) + al(a_tiles);

    HB_CUDA(cudaMemsetAsync(pair_len + np, 0, 8, g_stream));
    if (np > 0) {
        plan_pairs_kernel<<<blocks_for(np, 256), 256, 0, g_stream>>>(probe_pos, np, nlist, list_off, probes, pair_len,
                                                                     keys, vals);
        HB_LAUNCH_CHECK();
    }
    cub::DeviceScan::ExclusiveSum(cub_tmp, scan_bytes, pair_len, pair_out, (int)np + 1, g_stream);
    HB_LAUNCH_CHECK();
    if (np > 0) {
        cub::DeviceRadixSort::SortPairs(cub_tmp, sort_bytes, keys, keys_sorted, vals, qsel, (int)np, 0, end_bit,
                                        g_stream);
        HB_LAUNCH_CHECK();
    }
    plan_lists_kernel<<<blocks_for(nlist + 1, 128), 128, 0, g_stream>>>(keys_sorted, np, nlist, list_off, lq_off, tiles,
                                                                         tile_rows, tile_q);
    HB_LAUNCH_CHECK();
    cub::DeviceScan::ExclusiveSum(cub_tmp, scan2_bytes, tiles, tile_prefix, nlist + 1, g_stream);
    HB_LAUNCH_CHECK();
}

// ---- the same bookkeeping for the FAST list scan: three small kernels instead of two scans and a radix sort ----------------
// The candidate pass needs the pairs grouped by list, not sorted: a counting pass (atomics), one scan over the lists and a
// scatter.  pair_out holds the offsets of a query's probed lists inside THAT query's concatenation (its first entry is 0):
// every consumer takes differences within a query (unit_slots_kernel, ivf_resolve_kernel).  The order of the pairs inside a
// list is whatever the atomics give; no result depends on it (candidates are ranked by score, then position in the
// query's concatenation).
__global__ void planf_count_kernel(const int64_t *__restrict__ probe_pos, int pos_stride, int64_t nq, int npq, int nlist,
                                   const int64_t *__restrict__ list_off, int32_t *__restrict__ probes, int64_t *__restrict__ pair_out,
                                   int32_t *__restrict__ cnt) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int64_t run = 0;
    for (int j = 0; j < npq; ++j) {
        const int64_t p = q * npq + j;
        const int64_t l = probe_pos[q * pos_stride + j];
        const bool ok = l >= 0 && l < nlist;
        probes[p] = ok ? (int32_t)l : -1;
        if (pair_out) pair_out[p] = run;
        if (ok) {
            run += list_off[l + 1] - list_off[l];
            atomicAdd(&cnt[l], 1);
        }
    }
}
// lq_off = exclusive prefix of the pairs per list, unit_prefix = exclusive prefix of ceil(pairs / tile_q) for the lists that
// hold rows; cursor[l] = lq_off[l] (the scatter's write positions).  One block; both counts ride in one 64-bit shuffle scan.
__global__ void __launch_bounds__(1024) planf_scan_kernel(const int32_t *__restrict__ cnt, int nlist, const int64_t *__restrict__ list_off,
                                                          int tile_q, int64_t *__restrict__ lq_off, int64_t *__restrict__ unit_prefix,
                                                          int32_t *__restrict__ cursor) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_run;
    constexpr int PT = 4;  // lists per thread and round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base <= nlist; base += 1024 * PT) {
        const int l0 = base + threadIdx.x * PT;
        unsigned long long v[PT], x = 0;
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            const int l = l0 + j;
            v[j] = 0;
            if (l < nlist) {
                const unsigned long long c = (unsigned long long)(uint32_t)cnt[l];
                const unsigned long long units = list_off[l + 1] > list_off[l] ? (c + tile_q - 1) / tile_q : 0ull;
                v[j] = (units << 32) | c;
            }
            x += v[j];
        }
        const unsigned long long mine = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += t;
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long y = s_w[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, y, off);
                if (lane >= off) y += t;
            }
            s_w[lane] = y;
        }
        __syncthreads();
        unsigned long long pre = s_run + (warp > 0 ? s_w[warp - 1] : 0ull) + x - mine;
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            const int l = l0 + j;
            if (l <= nlist) {
                lq_off[l] = (int64_t)(pre & 0xffffffffull);
                unit_prefix[l] = (int64_t)(pre >> 32);
                if (l < nlist) cursor[l] = (int32_t)(pre & 0xffffffffull);
            }
            pre += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += s_w[31];
        __syncthreads();
    }
}
__global__ void planf_scatter_kernel(const int32_t *__restrict__ probes, int64_t np, int32_t *__restrict__ cursor,
                                     int32_t *__restrict__ qsel) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int l = probes[p];
    if (l >= 0) qsel[atomicAdd(&cursor[l], 1)] = (int32_t)p;
}
void ivf_plan_fast(const int64_t *probe_pos, int pos_stride, int64_t nq, int npq, int nlist, const int64_t *list_off, int32_t *probes,
                   int64_t *pair_out, int32_t *qsel, int64_t *lq_off, int64_t *unit_prefix, int tile_q, DevBuf &tmp) {
    const int64_t np = nq * npq;
    HB_REQUIRE(np < (1ll << 31), "too many (query, probe) pairs for one plan");
    int32_t *cnt = (int32_t *)tmp.get((size_t)(nlist + 1) * 8);
    int32_t *cursor = cnt + nlist + 1;
    HB_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(nlist + 1) * 4, g_stream));
    if (nq > 0) {
        planf_count_kernel<<<blocks_for(nq, 128), 128, 0, g_stream>>>(probe_pos, pos_stride, nq, npq, nlist, list_off, probes, pair_out,
                                                                      cnt);
        HB_LAUNCH_CHECK();
    }
    planf_scan_kernel<<<1, 1024, 0, g_stream>>>(cnt, nlist, list_off, tile_q, lq_off, unit_prefix, cursor);
    HB_LAUNCH_CHECK();
    if (np > 0) {
        planf_scatter_kernel<<<blocks_for(np, 256), 256, 0, g_stream>>>(probes, np, cursor, qsel);
        HB_LAUNCH_CHECK();
    }
}

void launch_ivf_resolve(const int64_t *pos, int64_t nq, int k, int nprobe, const int32_t *probes,
                        const int64_t *pair_out, const int64_t *list_off, const int64_t *list_rows, int64_t *out_ids, int32_t *ok,
                        const int32_t *ok_other) {
    if (nq * k == 0) return;
    ivf_resolve_kernel<<<blocks_for(nq * k, 256), 256, 0, g_stream>>>(pos, nq, k, nprobe, probes, pair_out, list_off,
                                                                      list_rows, out_ids, ok, ok_other);
    HB_LAUNCH_CHECK();
}
void launch_offset_ids(int64_t *ids, int64_t nq, int64_t run, int64_t stride, int64_t base) {
    if (nq * run == 0 || base == 0) return;
    offset_ids_kernel<<<blocks_for(nq * run, 256), 256, 0, g_stream>>>(ids, nq, run, stride, base);
    HB_LAUNCH_CHECK();
}
void launch_parts_to_query_major(const double *dist, const int64_t *ids, int nparts, int64_t nq, int k, double *out_dist,
                                 int64_t *out_ids) {
    const int64_t tot = (int64_t)nparts * nq * k;
    if (tot == 0) return;
    parts_to_query_major_kernel<<<blocks_for(tot, 256), 256, 0, g_stream>>>(dist, ids, nparts, nq, k, out_dist, out_ids);
    HB_LAUNCH_CHECK();
}
void launch_fill_f64(double *p, int64_t n, double v) {
    if (n == 0) return;
    fill_f64_kernel<<<blocks_for(n, 256), 256, 0, g_stream>>>(p, n, v);
    HB_LAUNCH_CHECK();
}
void launch_lookup_ids(const int64_t *pos, int64_t nq, int k, const int64_t *ids_in, int64_t stride, int64_t *out_ids) {
    if (nq * k == 0) return;
    lookup_ids_kernel<<<blocks_for(nq * k, 256), 256, 0, g_stream>>>(pos, nq, k, ids_in, stride, out_ids);
    HB_LAUNCH_CHECK();
}
void launch_pos_to_i32(const int64_t *pos, int64_t n, int32_t *out) {
    if (n == 0) return;
    pos_to_i32_kernel<<<blocks_for(n, 256), 256, 0, g_stream>>>(pos, n, out);
    HB_LAUNCH_CHECK();
}

}  // namespace hb
