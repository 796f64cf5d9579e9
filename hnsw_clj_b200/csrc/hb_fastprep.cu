// hb_fastprep.cu — preparation and finishing kernels of HB_MODE_FAST (see hb_fast.cuh for the scheme).
//
// Quantisation (both sides): u = max|x| / qmax, m_i = rint(x_i / u) computed in fp64, so |x_i - u*m_i| <= kFastRound*u.
// m is split into NS balanced signed 8-bit digits, most significant first.  For a pair (q, r):
//      | q.r - u_q*u_r*sum(mq_i*mr_i) |  <=  kFastRound * (u_q*||r||_1 + u_r*||q||_1) + kFastRound^2 * d * u_q*u_r
// which, divided by the norms (cosine), is the eps_q of launch_query_bounds once the per-row factors are
// replaced by their maxima over the index (stats[0], stats[1]).
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <float.h>
#include <math.h>

#include "hb_fast.cuh"

namespace cg = cooperative_groups;

namespace hb {
namespace {

inline int blocks_for(int64_t n, int bs) { return (int)ceil_div(n > 0 ? n : 1, bs); }

__device__ __forceinline__ int64_t img_offset(int r, int c) {  // byte (r, c) of a 128 x 128 B swizzled image
    return (int64_t)(r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 4) ^ (r & 7)) << 4) | (c & 15));
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// digits of m (|m| <= qmax), most significant first
template <int NS>
__device__ __forceinline__ void split_digits(int m, int8_t (&dg)[NS]) {
    if (NS == 2) {
        const int lo = (int)(int8_t)(m & 255);
        dg[1] = (int8_t)lo;
        dg[0] = (int8_t)((m - lo) >> 8);
    } else {
        const int l0 = (int)(int8_t)(m & 255);
        const int m1 = (m - l0) >> 8;
        const int l1 = (int)(int8_t)(m1 & 255);
        dg[NS - 1] = (int8_t)l0;
        dg[NS == 2 ? 0 : 1] = (int8_t)l1;
        dg[0] = (int8_t)((m1 - l1) >> 8);
    }
}

__global__ void tile_offsets_kernel(const int64_t *__restrict__ list_off, int nlist, int64_t *__restrict__ tile_off) {
    // single thread block; nlist <= a few 100k: a serial walk by one thread is microseconds and runs once per index
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int l = 0; l < nlist; ++l) {
            tile_off[l] = run;
            run += (list_off[l + 1] - list_off[l] + kFastTile - 1) / kFastTile;
        }
        tile_off[nlist] = run;
    }
}

__global__ void fill_f32_kernel(float *p, int64_t n, float v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// one warp per slab row
template <typename T, int NS>
__global__ void __launch_bounds__(256) quant_rows_kernel(const T *__restrict__ rows, int64_t n, int d, int kbn, int nlist,
                                                         const int64_t *__restrict__ list_off,
                                                         const int64_t *__restrict__ tile_off,
                                                         const double *__restrict__ norm, int8_t *__restrict__ img,
                                                         float *__restrict__ rs, float *__restrict__ ro,
                                                         float *__restrict__ stats) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    int lo = 0, hi = nlist;  // last l with list_off[l] <= r
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (list_off[mid] <= r) lo = mid;
        else hi = mid;
    }
    const int64_t in_list = r - list_off[lo];
    const int64_t tile = tile_off[lo] + in_list / kFastTile;
    const int tr = (int)(in_list % kFastTile);
    const T *x = rows + r * (int64_t)d;
    double amax = 0.0, l1 = 0.0;
    for (int i = lane; i < d; i += 32) {
        const double v = fabs(to_f64(x[i]));
        amax = fmax(amax, v);
        l1 += v;
    }
    amax = warp_max(amax);
    l1 = warp_sum(l1);
    const double u = amax > 0.0 ? amax / fast_qmax(NS) : 1.0;
    const double inv = 1.0 / u;
    int8_t *base = img + tile * kbn * NS * (int64_t)kFastImg;
    const int dpad = kbn * kFastKB;
    for (int i0 = lane * 4; i0 < dpad; i0 += 128) {  // 4 consecutive dims per lane: one 32-bit store per digit
        uint32_t w[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) w[s] = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            int m = 0;
            if (i < d) m = (int)rint(to_f64(x[i]) * inv);
            int8_t dg[NS];
            split_digits<NS>(m, dg);
#pragma unroll
            for (int s = 0; s < NS; ++s) w[s] |= (uint32_t)(uint8_t)dg[s] << (8 * e);
        }
        const int kb = i0 / kFastKB, c = i0 % kFastKB;
#pragma unroll
        for (int s = 0; s < NS; ++s)
            *reinterpret_cast<uint32_t *>(base + ((int64_t)kb * NS + s) * kFastImg + img_offset(tr, c)) = w[s];
    }
    if (lane == 0) {
        const double nr = norm ? norm[r] : 1.0;
        const bool usable = nr > 0.0 && isfinite(nr);
        const double w = usable ? 1.0 / nr : 0.0;
        const double f = NS == 2 ? 1.0 : 65536.0;
        rs[tile * kFastTile + tr] = (float)(u * w * f);
        ro[tile * kFastTile + tr] = usable ? 0.0f : -INFINITY;
        atomicMax(reinterpret_cast<int *>(&stats[0]), __float_as_int(__double2float_ru(u * w)));
        atomicMax(reinterpret_cast<int *>(&stats[1]), __float_as_int(__double2float_ru(l1 * w * (1.0 + 1e-9))));
        if (!usable) atomicMax(reinterpret_cast<int *>(&stats[2]), __float_as_int(1.0f));
    }
}

// one warp per query: plain digit rows [q][s][dpad]
template <typename T, int NS>
__global__ void __launch_bounds__(256) quant_queries_kernel(const T *__restrict__ queries, int64_t nq, int d, int kbn,
                                                            int8_t *__restrict__ dig, double *__restrict__ qu,
                                                            double *__restrict__ ql1) {
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const T *x = queries + q * (int64_t)d;
    double amax = 0.0, l1 = 0.0;
    for (int i = lane; i < d; i += 32) {
        const double v = fabs(to_f64(x[i]));
        amax = fmax(amax, v);
        l1 += v;
    }
    amax = warp_max(amax);
    l1 = warp_sum(l1);
    const double u = amax > 0.0 ? amax / fast_qmax(NS) : 1.0;
    const double inv = 1.0 / u;
    const int dpad = kbn * kFastKB;
    int8_t *base = dig + q * (int64_t)NS * dpad;
    for (int i0 = lane * 4; i0 < dpad; i0 += 128) {
        uint32_t w[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) w[s] = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            int m = 0;
            if (i < d) m = (int)rint(to_f64(x[i]) * inv);
            int8_t dg[NS];
            split_digits<NS>(m, dg);
#pragma unroll
            for (int s = 0; s < NS; ++s) w[s] |= (uint32_t)(uint8_t)dg[s] << (8 * e);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) *reinterpret_cast<uint32_t *>(base + (int64_t)s * dpad + i0) = w[s];
    }
    if (lane == 0) {
        qu[q] = u;
        ql1[q] = l1 * (1.0 + 1e-9);
    }
}

__global__ void unit_plan_kernel(int nlist, const int64_t *__restrict__ lq_off, const int64_t *__restrict__ unit_prefix,
                                 const int64_t *__restrict__ tile_off, int nunits, int nunits_real, int interleave,
                                 int tile_limit, int tile_div, int tile_start, UnitPlan U, int skip_tiles,
                                 const int32_t *__restrict__ qsel, int pair_div, unsigned long long *__restrict__ stats) {
    const int slot_u = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot_u > nunits) return;
    if (slot_u == nunits) {
        U.unit_ntile[slot_u] = 0;
        return;
    }
    // nunits_real < 0: `nunits` is the host's upper bound, the real count is the plan's unit_prefix[nlist] (device) --
    // a search then needs no host round trip between planning and the candidate pass
    if (nunits_real < 0) nunits_real = (int)unit_prefix[nlist];
    // interleave > 0: unit slots are laid out so that the contiguous item ranges the CTAs take visit the real units
    // round-robin (slot c*rows + r <- real unit r*interleave + c): the units of one list then run on neighbouring
    // CTAs at the same time and share their row tiles through L2.
    int u = slot_u;
    if (interleave > 0) {
        const int rows = nunits / interleave;
        u = (slot_u % rows) * interleave + slot_u / rows;
    }
    if (u >= nunits_real) {
        U.unit_list[slot_u] = 0;
        U.unit_sel0[slot_u] = 0;
        U.unit_nsel[slot_u] = 0;
        U.unit_ntile[slot_u] = 0;
        if (U.unit_tile0) U.unit_tile0[slot_u] = 0;
        return;
    }
    int lo = 0, hi = nlist;  // last l with unit_prefix[l] <= u (lists without selections share a prefix value:
    while (hi - lo > 1) {    //  take the LAST such list, the one that actually owns unit u)
        const int mid = (lo + hi) >> 1;
        if (unit_prefix[mid] <= u) lo = mid;
        else hi = mid;
    }
    const int l = lo;
    const int j = u - (int)unit_prefix[l];
    const int64_t sel0 = lq_off[l] + (int64_t)j * kFastTile;
    const int nsel = (int)min((int64_t)kFastTile, lq_off[l + 1] - sel0);
    int nt = (int)(tile_off[l + 1] - tile_off[l]);
    if (tile_div > 1) nt = (nt + tile_div - 1) / tile_div;  // tiles 0, tile_div, 2*tile_div, ...
    int t0 = tile_start;
    if (skip_tiles > 0 && pair_div > 0) {
        // every selection of the unit is its query's nearest list (probe rank 0): the sample pass has scored the first tiles
        bool all0 = true;
        for (int s2 = 0; s2 < nsel && all0; ++s2) {
            const int64_t pr = qsel ? (int64_t)qsel[sel0 + s2] : sel0 + s2;
            all0 = pr % pair_div == 0;
        }
        if (all0) t0 = skip_tiles;
    }
    nt = max(nt - t0, 0);                                   // tiles t0 .. (tile_div == 1 only)
    if (tile_limit > 0) nt = min(nt, tile_limit);
    if (U.unit_tile0) U.unit_tile0[slot_u] = t0;
    if (stats && nsel > 0) {
        // what the pass covers: [0] units, [1] items (unit x row tile), [2] distinct row tiles read (a list with several units:
        // at least its tiles beyond the skipped ones), [3] units of <= 64 selections, [4] narrow units, [5] their items,
        // [6] the query slots their images hold
        // (one atomic per counter and warp: thousands of units adding to the same six words serialise in L2)
        unsigned long long c[7] = {1ull, (unsigned long long)nt, 0ull, nsel <= 64 ? 1ull : 0ull, 0ull, 0ull, 0ull};
        if (j == 0) {
            const int units_l = (int)(unit_prefix[l + 1] - unit_prefix[l]);
            const int all_t = (int)(tile_off[l + 1] - tile_off[l]);
            c[2] = (unsigned long long)(units_l == 1 ? nt : max(all_t - skip_tiles, 0));
        }
        if (nsel <= kNarrowSlots) {
            c[4] = 1ull;
            c[5] = (unsigned long long)nt;
            c[6] = (unsigned long long)((nsel + 7) & ~7);  // query slots of the unit's image that are packed and read
        }
        const cg::coalesced_group g = cg::coalesced_threads();
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            const unsigned long long sum = cg::reduce(g, c[m], cg::plus<unsigned long long>());
            if (g.thread_rank() == 0 && sum != 0) atomicAdd(stats + m, sum);
        }
    }
    U.unit_list[slot_u] = l;
    U.unit_sel0[slot_u] = (int32_t)sel0;
    U.unit_nsel[slot_u] = nsel;
    U.unit_ntile[slot_u] = nt;
}

// exclusive prefix of ntile -> item0 (single block).  With item0n: the units of at most kNarrowSlots selections are counted in
// item0n (tc_narrow_kernel's items), the others in item0 (tc_pass_kernel's).  Both counts ride in one 64-bit value through
// one shuffle scan per 1024 entries (the plan of a 65,536-list shard has 68 k unit slots: the earlier shared-memory
// Hillis-Steele scan, two passes, took 0.3 ms there).
__global__ void __launch_bounds__(1024) unit_scan_kernel(const int32_t *__restrict__ ntile, int count, int32_t *__restrict__ item0_all,
                                                         const int32_t *__restrict__ nsel, int32_t *__restrict__ item0n) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_run;
    constexpr int PT = 4;  // entries per thread and round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < count; base += 1024 * PT) {
        const int i0 = base + threadIdx.x * PT;
        unsigned long long v[PT], x = 0;
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            const int i = i0 + j;
            v[j] = 0;
            if (i < count) {
                const unsigned long long nt = (unsigned long long)(uint32_t)ntile[i];
                const bool narrow = item0n != nullptr && i < count - 1 && nsel[i] <= kNarrowSlots;
                v[j] = narrow ? nt << 32 : nt;
            }
            x += v[j];
        }
        const unsigned long long mine = x;  // this thread's total
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
            const int i = i0 + j;
            if (i < count) {
                item0_all[i] = (int32_t)(uint32_t)(pre & 0xffffffffull);
                if (item0n) item0n[i] = (int32_t)(uint32_t)(pre >> 32);
            }
            pre += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += s_w[31];
        __syncthreads();
    }
}

__global__ void unit_slots_kernel(int nunits, const int32_t *__restrict__ unit_sel0, const int32_t *__restrict__ unit_nsel,
                                  const int32_t *__restrict__ qsel, const int64_t *__restrict__ pair_out, int pair_div,
                                  const int32_t *__restrict__ pair_query, int32_t *__restrict__ slot_query,
                                  int32_t *__restrict__ slot_rel0, bool narrow) {
    // a warp per unit slot.  The unit slots of an IVF plan are an upper bound and mostly empty: an empty one gets its first slot
    // marked (-1, what pack_units_kernel tests), a narrow unit its first kNarrowSlots (all tc_narrow_kernel reads), the others 128
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nunits) return;
    const int u = (int)w, ns = unit_nsel[u];
    const int need = ns == 0 ? 1 : ((narrow && ns <= kNarrowSlots) ? kNarrowSlots : kFastTile);
    for (int s = lane; s < need; s += 32) {
        int q = -1, rel0 = 0;
        if (s < ns) {
            const int64_t sel = (int64_t)unit_sel0[u] + s;
            const int64_t p = qsel ? (int64_t)qsel[sel] : sel;
            q = pair_query ? pair_query[p] : (pair_div > 0 ? (int)(p / pair_div) : (int)p);
            if (pair_out) rel0 = (int32_t)(pair_out[p] - pair_out[(int64_t)q * (pair_div > 0 ? pair_div : 1)]);
        }
        slot_query[(int64_t)u * kFastTile + s] = q;
        slot_rel0[(int64_t)u * kFastTile + s] = rel0;
    }
}

// copies the 16-byte chunks of the units' query digits into swizzled images; a block walks (unit, k-block) pairs with a
// grid stride: the unit slots of an IVF plan are an upper bound (ceil(pairs / 128) + nlist) and mostly empty — 68 k slots
// for ~9 k units on a 65,536-list shard — and one block per slot spent 0.3 ms per plan on blocks that only returned
template <int NS>
__global__ void __launch_bounds__(256) pack_units_kernel(const int8_t *__restrict__ dig, int kbn, int nunits,
                                                         const int32_t *__restrict__ slot_query, int8_t *__restrict__ aimg,
                                                         const int32_t *__restrict__ unit_nsel, bool narrow) {
    const int dpad = kbn * kFastKB;
    for (int64_t w = blockIdx.x; w < (int64_t)nunits * kbn; w += gridDim.x) {
    const int u = (int)(w / kbn), kb = (int)(w % kbn);
    if (slot_query[(int64_t)u * kFastTile] < 0) continue;  // slots fill from 0: an empty unit (padding / bound) has no items
    // M = 64 units: the first 8 row groups of each image; narrow units: the first 4
    const int ns_u = unit_nsel != nullptr ? unit_nsel[u] : kFastTile;
    // narrow units: only the 8-slot groups in use (tc_narrow_kernel copies no more than those)
    const int nslots = (narrow && ns_u <= kNarrowSlots) ? ((ns_u + 7) & ~7) : (ns_u <= 64 ? 64 : kFastTile);
    int8_t *dst = aimg + ((int64_t)u * kbn + kb) * NS * kFastImg;
    for (int i = threadIdx.x; i < NS * nslots * 8; i += blockDim.x) {
        const int ch = i & 7, slot = (i >> 3) % nslots, s = i / (8 * nslots);
        const int q = slot_query[(int64_t)u * kFastTile + slot];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q >= 0) v = *reinterpret_cast<const uint4 *>(dig + ((int64_t)q * NS + s) * dpad + kb * kFastKB + ch * 16);
        *reinterpret_cast<uint4 *>(dst + (int64_t)s * kFastImg + img_offset(slot, ch * 16)) = v;
    }
    }
}

__global__ void query_bounds_kernel(const double *__restrict__ qu, const double *__restrict__ ql1,
                                    const double *__restrict__ qnorm, int64_t nq, int ns, int d, int metric,
                                    const float *__restrict__ stats, double *__restrict__ q_scale,
                                    double *__restrict__ q_eps, float *__restrict__ q_margin, float *__restrict__ thr_init,
                                    int32_t *__restrict__ cnt_init) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    if (thr_init) thr_init[q] = -INFINITY;  // the start of a job: no threshold, no candidates
    if (cnt_init) cnt_init[q] = 0;
    const double w = metric == HB_COSINE ? 1.0 / qnorm[q] : 1.0;
    const double umax_r = (double)stats[0], l1max_r = (double)stats[1];
    const double us = qu[q] * w;
    double eps = kFastRound * (us * l1max_r + umax_r * ql1[q] * w) + kFastRound * kFastRound * d * us * umax_r;
    // the epilogue drops the low 8 bits of the lowest-weight accumulator and rounds one int32 -> fp32 conversion:
    // < 2 * 256 units of the integer score (in units of the accumulator weight 2^0 (ns = 2) or 2^16 (ns = 3))
    eps += 1024.0 * (ns == 3 ? 65536.0 : 1.0) * us * umax_r;
    if (ns == 3) eps += us * umax_r * (double)d * 16384.0 * 513.0;  // dropped digit products (1,2), (2,1), (2,2)
    eps *= 1.0 + 1e-6;
    const bool bad = !(w > 0.0) || !isfinite(w) || stats[2] > 0.0f;
    q_scale[q] = bad ? 0.0 : us;
    q_eps[q] = bad ? INFINITY : eps;
    // in score units (similarity = score * q_scale): two candidate scores further apart than this are ordered like their
    // exact similarities, with room for the final test's own slack
    q_margin[q] = (bad || !(us > 0.0)) ? INFINITY : __double2float_ru(2.2 * eps / us);
}

// unused candidate slots hold +inf (-score): the selection may return them when a query has fewer than kk candidates
// Only the candidates that can still reach the exact top-k are re-scored: the k best by approximate score and every
// further one within `margin` (> 2 eps_q) of the k-th.  One warp per query writes, per selected slot, the row (or -1:
// not re-scored, counted as a rejected row in fast_final_kernel's proof; its exact distance reads +inf) and appends the
// wanted (query, row, slot) triples to a dense pair list, so that the re-score kernel runs full warps.
__global__ void __launch_bounds__(256) rescore_pairs_kernel(const int64_t *__restrict__ sel_pos, const double *__restrict__ sel_negv,
                                                            const int32_t *__restrict__ cand_pos, int64_t nq, int kk, int cap, int k,
                                                            const float *__restrict__ margin, int32_t *__restrict__ slot_row,
                                                            double *__restrict__ exact, int32_t *__restrict__ total,
                                                            int32_t *__restrict__ pair_query, int32_t *__restrict__ pair_row,
                                                            int32_t *__restrict__ pair_slot, int set_only) {
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const double nk = k <= kk ? sel_negv[q * kk + k - 1] : INFINITY;  // the list is ascending in -score
    const double lim = nk + (double)margin[q] + 4e-6 * fabs(nk);
    // set_only (the caller wants the top-k SET, not its order or distances): a candidate more than `margin` (> 2 eps_q)
    // above the k-th approximate score is in the exact top-k whatever its exact distance is -- fewer than k rows can
    // reach its lower bound -- so only the candidates within `margin` of the k-th on either side are re-scored
    const double sure = set_only && nk < INFINITY ? nk - (double)margin[q] - 4e-6 * fabs(nk) : -INFINITY;
    int row[4] = {-1, -1, -1, -1};
    int nwant = 0;
    unsigned bal[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int j = h * 32 + lane;
        bool want = false;
        if (j < kk) {
            const int64_t p = sel_pos[q * kk + j];
            const double nv = sel_negv[q * kk + j];
            want = p >= 0 && nv < INFINITY && (j < k || !(nk < INFINITY) || nv <= lim);
            const bool certain = want && nv < sure;
            if (certain) want = false;
            if (want) row[h] = cand_pos[q * cap + p];
            slot_row[q * kk + j] = certain ? -2 : row[h];
            if (!want) exact[q * kk + j] = certain ? -INFINITY : INFINITY;
        }
        bal[h] = __ballot_sync(0xffffffffu, want);
        nwant += __popc(bal[h]);
    }
    int base = 0;
    if (lane == 0 && nwant > 0) base = atomicAdd(total, nwant);
    base = __shfl_sync(0xffffffffu, base, 0);
    int before = 0;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        if (row[h] >= 0) {
            const int o = base + before + __popc(bal[h] & ((1u << lane) - 1u));
            pair_query[o] = (int32_t)q;
            pair_row[o] = row[h];
            pair_slot[o] = (int32_t)(q * kk + h * 32 + lane);
        }
        before += __popc(bal[h]);
    }
}

// One CTA (128 threads) per query.  Entry j < kk: exact distance + rel; rank by (distance key, rel).
// One WARP per query (four queries per block): the kernel is a chain of dependent loads per query — selection -> pair -> exact
// distance / candidate record, then the proof's scalars — and a block per query kept 16 of those chains in flight per SM;
// this keeps 64.  Lane l owns the selections l, l + 32, ... (kk <= 128: at most four).
__global__ void __launch_bounds__(128) fast_final_kernel(const FinalParams P) {
    __shared__ uint64_t s_key_all[4][128];
    __shared__ int32_t s_rel_all[4][128];
    __shared__ uint64_t s_okey_all[4][129];  // the k + 1 best in rank order: ties across lists (tie_list_off)
    __shared__ int32_t s_orow_all[4][129];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + warp;
    if (q >= P.nq) return;
    uint64_t *s_key = s_key_all[warp], *s_okey = s_okey_all[warp];
    int32_t *s_rel = s_rel_all[warp], *s_orow = s_orow_all[warp];
    const int kk = P.kk, k = P.k;
    const int cnt = P.cnt[q];
    constexpr int U = 4;
    uint64_t key[U];
    int32_t rel[U], prow[U];
    double dist[U];
    int skipped = INT_MAX;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int j = u * 32 + lane;
        key[u] = kKeyEmpty;
        rel[u] = INT_MAX;
        dist[u] = INFINITY;
        prow[u] = -1;
        if (j < kk) {
            const int64_t p = P.sel_pos[q * kk + j];
            if (p >= 0 && P.sel_negv[q * kk + j] < INFINITY) {
                prow[u] = P.pair_row[q * kk + j];
                if (prow[u] >= 0 || prow[u] == -2) {  // re-scored, or (set_only) certainly in the top-k: exact reads -inf
                    dist[u] = P.exact[q * kk + j];
                    // a certain candidate keeps its approximate rank j (the selection is best-first), ahead of every re-scored one:
                    // column 0 stays the (approximately) nearest row, which seeds the IVF scan's thresholds
                    key[u] = prow[u] == -2 ? (uint64_t)j : dist_key(dist[u]);
                    rel[u] = P.cand_rel[q * P.cap + p];
                } else {
                    skipped = min(skipped, j);  // selected but not re-scored: the best of them bounds the others
                }
            }
            s_key[j] = key[u];
            s_rel[j] = rel[u];
        }
    }
    skipped = __reduce_min_sync(0xffffffffu, skipped);
    __syncwarp();
    int nvalid = 0;
    int rank[U] = {0, 0, 0, 0};
    for (int i = 0; i < kk; ++i) {
        const uint64_t ki = s_key[i];
        const int32_t ri = s_rel[i];
        nvalid += ki != kKeyEmpty;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u * 32 >= kk) break;  // (uniform: kk = 64 uses two of the four)
            const int j = u * 32 + lane;
            if (i != j && (ki < key[u] || (ki == key[u] && (ri < rel[u] || (ri == rel[u] && i < j))))) ++rank[u];
        }
    }
    double kth = INFINITY;  // exact k-th best distance among the selected (the lane that holds it)
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int j = u * 32 + lane;
        if (j < kk && key[u] != kKeyEmpty && rank[u] < k) {
            P.out_rel[q * k + rank[u]] = rel[u];
            P.out_dist[q * k + rank[u]] = dist[u];
            if (rank[u] == min(k, nvalid) - 1) kth = dist[u];
            if (P.out_simub) {  // upper bound of the winner's similarity: exact if re-scored, approximate score + eps_q if certain
                const double sim = prow[u] == -2 ? -P.sel_negv[q * kk + j] * P.q_scale[q] + P.q_eps[q]
                                                 : (P.metric == HB_COSINE ? 1.0 - dist[u] : -dist[u]);
                P.out_simub[q * k + rank[u]] = sim + 1e-9 * (1.0 + fabs(sim));
            }
        }
    }
    {   // one lane at most holds a finite kth (ranks are distinct); +inf otherwise, as before
        const unsigned has = __ballot_sync(0xffffffffu, kth != INFINITY);
        kth = __shfl_sync(0xffffffffu, kth, has ? __ffs(has) - 1 : 0);
    }
    int tie = 0;
    if (P.tie_list_off) {
        // The caller's `rel` order across lists is not the reference's (approximate probe order, see set_only): equal
        // distances inside a list still fall in row order, but a tie between rows of different lists among the k + 1 best
        // would be broken by probe rank in the reference (ivf_flat.clj:281-294) -- such a query goes to the exact path.
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = u * 32 + lane;
            if (j < kk && key[u] != kKeyEmpty && rank[u] <= k) s_okey[rank[u]] = key[u], s_orow[rank[u]] = prow[u];
        }
        __syncwarp();
        for (int j = lane; j < k; j += 32) {
            if (j + 1 < nvalid && s_okey[j] == s_okey[j + 1]) {
                auto list_of = [&](int32_t row) {
                    int lo = 0, hi = P.tie_nlist;  // last l with list_off[l] <= row
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (P.tie_list_off[mid] <= row) lo = mid; else hi = mid;
                    }
                    return lo;
                };
                if (list_of(s_orow[j]) != list_of(s_orow[j + 1])) tie = 1;
            }
        }
        tie = __any_sync(0xffffffffu, tie);
    }
    for (int r = nvalid + lane; r < k; r += 32) {
        P.out_rel[q * k + r] = -1;
        P.out_dist[q * k + r] = INFINITY;
        if (P.out_simub) P.out_simub[q * k + r] = INFINITY;
    }
    if (lane == 0) {
        // rejected rows: selected but not re-scored (score <= the best of them), emitted but not selected (score <= worst
        // selected) or never emitted (score < thr)
        bool ok = cnt <= P.cap && !tie;
        const float thr = P.thr[q];
        const bool none_rejected = cnt <= kk && thr == -INFINITY && skipped == INT_MAX;
        if (ok && !none_rejected) {
            if (nvalid < k) ok = false;  // rejected rows would be needed to fill k
            else {
                double t_v = (double)thr;
                if (cnt > kk) t_v = fmax(t_v, -P.sel_negv[q * kk + kk - 1]);
                if (skipped != INT_MAX) t_v = fmax(t_v, -P.sel_negv[q * kk + skipped]);
                const double t_sim = t_v * P.q_scale[q];
                const double t_up = t_sim + P.q_eps[q] + 1e-6 * fabs(t_sim);
                const double sim_k = P.metric == HB_COSINE ? 1.0 - kth : -kth;
                ok = (sim_k - t_up) > 1e-12 && isfinite(t_up);
            }
        }
        P.out_ok[q] = ok ? 1 : 0;
    }
}


// ---- candidate selection: the kk best (lowest -score) of the cnt[q] candidates a query collected -------------------
// One CTA per query; only the min(cnt, cap) live entries are read.  Keys = (order-preserving bits of the fp32 score,
// slot) packed in 64 bits, bitonic-sorted in shared memory at the next power of two >= count.
__device__ __forceinline__ uint32_t f32_desc_key(float v) {  // larger score -> smaller key
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // ascending order-preserving
    return ~b;
}
struct CompactParams {  // non-null cand_rel: keep only the kk best candidates, in place, and raise the query's threshold
    double *cand_negv = nullptr;
    int32_t *cand_rel = nullptr, *cand_pos = nullptr;
    int32_t *cnt_rw = nullptr;
    float *thr = nullptr;
    const float *margin = nullptr;
    int k = 0;
};
__device__ __forceinline__ void cand_select_block(uint64_t *s_keys, int64_t q, const double *__restrict__ cand_negv,
                                                  const int32_t *__restrict__ cnt, int kk, int cap, double *__restrict__ sel_negv,
                                                  int64_t *__restrict__ sel_pos, const CompactParams &CP, int skip_upto) {
    const int craw = cnt[q];
    const int n = min(craw, cap);
    if (craw <= cap && n <= skip_upto) return;  // cand_select_warp_kernel has served this query
    int m = kk > 64 ? 128 : 64;
    while (m < n) m <<= 1;
    const double *v = cand_negv + q * cap;
    for (int i = threadIdx.x; i < m; i += blockDim.x)
        s_keys[i] = i < n ? ((uint64_t)f32_desc_key((float)(-v[i])) << 32) | (uint32_t)i : ~0ull;
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < m / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const uint64_t a = s_keys[lo], b = s_keys[hi];
                if ((a > b) == asc) {
                    s_keys[lo] = b;
                    s_keys[hi] = a;
                }
            }
        }
    }
    __syncthreads();
    if (CP.cand_rel == nullptr) {
        for (int j = threadIdx.x; j < kk; j += blockDim.x) {
            const bool ok = j < n;
            const int slot = ok ? (int)(uint32_t)s_keys[j] : -1;
            sel_pos[q * kk + j] = slot;
            sel_negv[q * kk + j] = ok ? v[slot] : INFINITY;
        }
        return;
    }
    // compaction (kk <= blockDim.x): the kk best move to the front of the query's candidate list, best first; the
    // discarded ones score at most the kk-th best, which the threshold now records for the proof
    const int j = threadIdx.x;
    const bool ok = j < kk && j < n;
    const int slot = ok ? (int)(uint32_t)s_keys[j] : -1;
    double nv = INFINITY;
    int32_t rel = 0, pos = 0;
    if (ok) {
        nv = v[slot];
        rel = CP.cand_rel[q * cap + slot];
        pos = CP.cand_pos[q * cap + slot];
    }
    __syncthreads();
    if (j < kk) {
        if (ok && craw <= cap) {
            CP.cand_negv[q * cap + j] = nv;
            CP.cand_rel[q * cap + j] = rel;
            CP.cand_pos[q * cap + j] = pos;
        }
        sel_pos[q * kk + j] = ok ? j : -1;
        sel_negv[q * kk + j] = nv;
    }
    if (craw > cap) return;  // overflowed: the count stays above cap and the query goes to the exact path
    if (j == 0) CP.cnt_rw[q] = min(n, kk);
    if (n >= kk && j == kk - 1) CP.thr[q] = fmaxf(CP.thr[q], (float)(-nv));
    __syncthreads();
    if (CP.k <= kk && n >= CP.k && j == CP.k - 1) {
        const float sk = (float)(-nv);
        const float t = sk - CP.margin[q] - 4e-6f * fabsf(sk);
        if (t > CP.thr[q]) CP.thr[q] = t;
    }
}
// the queries the warp kernel left (more than kWarpSelMax candidates, or overflowed): a small grid walks all queries — nearly
// all of them return on the first test, and a block per query was 10,000 blocks launched to do so
__global__ void __launch_bounds__(256) cand_select_kernel(const double *__restrict__ cand_negv, const int32_t *__restrict__ cnt,
                                                          int64_t nq, int kk, int cap, double *__restrict__ sel_negv,
                                                          int64_t *__restrict__ sel_pos, const CompactParams CP, int skip_upto) {
    extern __shared__ uint64_t s_keys[];
    for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
        __syncthreads();  // the previous query's keys are consumed
        cand_select_block(s_keys, q, cand_negv, cnt, kk, cap, sel_negv, sel_pos, CP, skip_upto);
    }
}

// The same selection for the queries with at most kWarpSelMax candidates — nearly all of them once thresholds are in place —
// by one WARP per query (four queries per block, __syncwarp instead of __syncthreads, 4 KB of shared memory per query): the
// block-per-query kernel above holds eight queries per SM in flight and spends most of its time in load latency and barriers
// (0.11 ms per call for 10,000 queries whatever their length); this one holds 64.  Keys are unique (the slot index is the low
// word), so both kernels produce the same order.  Longer lists are left to cand_select_kernel(skip_upto = kWarpSelMax).
constexpr int kWarpSelMax = 512;
// Bitonic sort of 32 * EPL unique 64-bit keys held in registers, lane-blocked (position i = lane * EPL + e): strides below EPL
// are compare-exchanges between a lane's own registers, the others one __shfl_xor per key.  The same network in shared memory
// is bound by the shared-memory pipe (64-bit keys, ~12 wavefronts per compare-exchange step of a warp): 80 us for 10,000
// lists of 256 candidates.
template <int EPL>
__device__ __forceinline__ void warp_sort_regs(uint64_t (&k)[EPL], int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * EPL; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride < EPL) {
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    if ((e & stride) == 0) {
                        const bool asc = (((lane * EPL + e) & size) == 0);
                        const uint64_t a = k[e], b = k[e | stride];
                        if ((a > b) == asc) {
                            k[e] = b;
                            k[e | stride] = a;
                        }
                    }
                }
            } else {
                const int ld = stride / EPL;  // partner lane distance
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, k[e], ld);
                    const int i = lane * EPL + e;
                    const bool keep_min = (((i & size) == 0) == ((i & stride) == 0));
                    k[e] = keep_min ? (o < k[e] ? o : k[e]) : (o > k[e] ? o : k[e]);
                }
            }
        }
    }
}
// loads the query's n candidate keys (candidate c = e * 32 + lane: coalesced), sorts them, writes them to keys[] in order
template <int EPL>
__device__ __forceinline__ void warp_sort_candidates(const double *__restrict__ v, int n, int lane, uint64_t *keys) {
    uint64_t k[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int c = e * 32 + lane;
        k[e] = c < n ? ((uint64_t)f32_desc_key((float)(-v[c])) << 32) | (uint32_t)c : ~0ull;
    }
    warp_sort_regs<EPL>(k, lane);
#pragma unroll
    for (int e = 0; e < EPL; ++e) keys[lane * EPL + e] = k[e];
}
__global__ void __launch_bounds__(128) cand_select_warp_kernel(const double *__restrict__ cand_negv, const int32_t *__restrict__ cnt,
                                                               int64_t nq, int kk, int cap, double *__restrict__ sel_negv,
                                                               int64_t *__restrict__ sel_pos, const CompactParams CP) {
    __shared__ uint64_t s_all[4][kWarpSelMax];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + warp;
    if (q >= nq) return;
    const int craw = cnt[q];
    const int n = min(craw, cap);
    if (craw > cap || n > kWarpSelMax) return;  // overflowed or long: the block kernel
    uint64_t *keys = s_all[warp];
    int m = kk > 64 ? 128 : 64;
    while (m < n) m <<= 1;
    const double *v = cand_negv + q * cap;
    if (m == 64) warp_sort_candidates<2>(v, n, lane, keys);
    else if (m == 128) warp_sort_candidates<4>(v, n, lane, keys);
    else if (m == 256) warp_sort_candidates<8>(v, n, lane, keys);
    else warp_sort_candidates<16>(v, n, lane, keys);
    __syncwarp();
    if (CP.cand_rel == nullptr) {
        for (int j = lane; j < kk; j += 32) {
            const bool ok = j < n;
            const int slot = ok ? (int)(uint32_t)keys[j] : -1;
            sel_pos[q * kk + j] = slot;
            sel_negv[q * kk + j] = ok ? v[slot] : INFINITY;
        }
        return;
    }
    // compaction: read the kk best (kk <= 128: four per lane), then write them to the front of the list, best first
    double nv[4];
    int32_t rel[4], pos[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = u * 32 + lane;
        const bool ok = j < kk && j < n;
        const int slot = ok ? (int)(uint32_t)keys[j] : 0;
        nv[u] = ok ? v[slot] : INFINITY;
        rel[u] = ok ? CP.cand_rel[q * cap + slot] : 0;
        pos[u] = ok ? CP.cand_pos[q * cap + slot] : 0;
    }
    __syncwarp();
    double *s_nv = reinterpret_cast<double *>(keys);  // the sorted values, for the two threshold reads below
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = u * 32 + lane;
        if (j >= kk) continue;
        const bool ok = j < n;
        if (ok) {  // (n <= kWarpSelMax <= cap: no overflow here)
            CP.cand_negv[q * cap + j] = nv[u];
            CP.cand_rel[q * cap + j] = rel[u];
            CP.cand_pos[q * cap + j] = pos[u];
        }
        sel_pos[q * kk + j] = ok ? j : -1;
        sel_negv[q * kk + j] = nv[u];
        s_nv[j] = nv[u];
    }
    __syncwarp();
    if (lane == 0) {
        CP.cnt_rw[q] = min(n, kk);
        float t = CP.thr[q];
        if (n >= kk) t = fmaxf(t, (float)(-s_nv[kk - 1]));
        if (CP.k <= kk && n >= CP.k) {
            const float sk = (float)(-s_nv[CP.k - 1]);
            const float t2 = sk - CP.margin[q] - 4e-6f * fabsf(sk);
            if (t2 > t) t = t2;
        }
        CP.thr[q] = t;
    }
}

// ---- dense selection: short flat scans (<= 2048 rows: IVF coarse routing, k-means assignment) --------------------------
// The candidate pass writes every score of the short list (hb_tc.cu, FAST_DUMP); one warp per query then finds its k-th
// best score exactly (bisection on the order-preserving bits, counting across the warp), keeps the rows within
// `margin` of it as the candidate list (sorted, best first) and records the best rejected score as the threshold the
// proof needs.  Replaces the threshold / emit / sort machinery where the whole score row fits a warp's registers.
__device__ __forceinline__ uint32_t f32_asc_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_from_asc_key(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
template <int NPL>
__global__ void __launch_bounds__(128) dense_select_kernel(const float *__restrict__ dump, int ntiles, int64_t nq, int nrows, int k,
                                                           int kk, int cap, const float *__restrict__ margin,
                                                           double *__restrict__ cand_negv, int32_t *__restrict__ cand_rel,
                                                           int32_t *__restrict__ cand_pos, int32_t *__restrict__ cnt,
                                                           float *__restrict__ thr, double *__restrict__ sel_negv,
                                                           int64_t *__restrict__ sel_pos) {
    __shared__ uint32_t s_key[4][128];
    __shared__ int32_t s_row[4][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + warp;
    if (q >= nq) return;
    const float *base = dump + ((q >> 7) * ntiles) * (int64_t)(kFastTile * kFastTile) + (q & 127) * kFastTile;
    uint32_t key[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int r = i * 32 + lane;
        const float v = r < nrows ? base[(int64_t)(r >> 7) * (kFastTile * kFastTile) + (r & 127)] : -INFINITY;
        key[i] = f32_asc_key(v);
    }
    const uint32_t key_ninf = f32_asc_key(-INFINITY);
    // k-th largest key: the largest x with count(key >= x) >= k
    uint32_t x = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t c = x | (1u << bit);
        int n = 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) n += key[i] >= c;
        n = __reduce_add_sync(0xffffffffu, n);
        if (n >= k) x = c;
    }
    // cut: rows scoring at least (k-th best - margin) are candidates; fewer than k real rows: all of them
    float cutf = -FLT_MAX;
    if (x > key_ninf) {
        const float sk = f32_from_asc_key(x);
        cutf = fmaxf(sk - margin[q] - 4e-6f * fabsf(sk), -FLT_MAX);
    }
    const uint32_t cut = f32_asc_key(cutf);
    int m = 0;
    uint32_t best_rej = 0;  // largest key below the cut
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const bool in = key[i] >= cut;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (in) {
            const int o = m + __popc(bal & ((1u << lane) - 1u));
            if (o < 128) {
                s_key[warp][o] = key[i];
                s_row[warp][o] = i * 32 + lane;
            }
        } else {
            best_rej = max(best_rej, key[i]);
        }
        m += __popc(bal);
    }
    best_rej = __reduce_max_sync(0xffffffffu, best_rej);
    __syncwarp();
    if (m > kk) {  // too many rows within the margin: the exact path answers this query
        if (lane == 0) cnt[q] = cap + 1;
        for (int j = lane; j < kk; j += 32) {
            sel_pos[q * kk + j] = -1;
            sel_negv[q * kk + j] = INFINITY;
        }
        return;
    }
    // rank by (score descending, row ascending) and write the candidate list, best first
    for (int c = lane; c < m; c += 32) {
        const uint32_t kc = s_key[warp][c];
        const int rc = s_row[warp][c];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
            const uint32_t kj = s_key[warp][j];
            rank += (kj > kc) || (kj == kc && s_row[warp][j] < rc);
        }
        const double nv = -(double)f32_from_asc_key(kc);
        cand_negv[q * cap + rank] = nv;
        cand_rel[q * cap + rank] = rc;
        cand_pos[q * cap + rank] = rc;
        sel_negv[q * kk + rank] = nv;
        sel_pos[q * kk + rank] = rank;
    }
    for (int j = m + lane; j < kk; j += 32) {
        sel_pos[q * kk + j] = -1;
        sel_negv[q * kk + j] = INFINITY;
    }
    if (lane == 0) {
        cnt[q] = m;
        thr[q] = best_rej > key_ninf ? f32_from_asc_key(best_rej) : -INFINITY;
    }
}

// ---- exact re-score of (query, row) pairs: one thread per pair walks the reference's sequential fp64 sum
// (src/hnsw/ultra_fast.clj:53-95, ivf_flat.clj:224-226); a warp stages its 32 rows chunk by chunk through shared
// memory with full-line loads (the rows are scattered, 3-6 KB each); the queries were widened to fp64 once per call
// (one query per group of kk pairs: broadcast loads).
template <typename T>
struct Widen;
template <>
struct Widen<float> {
    static constexpr int PER = 4;  // elements per 16 bytes
    __device__ static __forceinline__ double at(const uint4 &v, int e) {
        return (double)__uint_as_float(e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w);
    }
};
template <>
struct Widen<__nv_bfloat16> {
    static constexpr int PER = 8;
    __device__ static __forceinline__ double at(const uint4 &v, int e) {
        const uint32_t w = (e >> 1) == 0 ? v.x : (e >> 1) == 1 ? v.y : (e >> 1) == 2 ? v.z : v.w;
        return (double)__uint_as_float((e & 1) ? (w & 0xFFFF0000u) : (w << 16));
    }
};
template <>
struct Widen<double> {
    static constexpr int PER = 2;
    __device__ static __forceinline__ double at(const uint4 &v, int e) {
        return e == 0 ? __hiloint2double((int)v.y, (int)v.x) : __hiloint2double((int)v.w, (int)v.z);
    }
};

constexpr int RS_QSLOTS = 4;  // distinct queries per warp whose chunks are staged in shared memory (others: direct loads)
// QLANE (fp64 rows only): every lane's query chunk is staged like its row chunk (a second 128-byte piece per lane and chunk).
// The run slots above serve pair lists with ~35 pairs per query; the coarse stage that proves a set re-scores ~2 centroids per
// query, a warp then holds ~18 queries, 14 of them fell to per-element global loads inside the dependent fp64 chain, and the
// launch took 0.16 ms for 18,000 pairs.
template <typename TRow, int ARITH, bool QLANE = false>
__global__ void __launch_bounds__(128) rescore_kernel(const TRow *__restrict__ rows, const double *__restrict__ row_norm,
                                                      const double *__restrict__ queries, const double *__restrict__ q_norm, int d,
                                                      const int32_t *__restrict__ pair_query, const int32_t *__restrict__ pair_row,
                                                      const int32_t *__restrict__ pair_slot, const int32_t *__restrict__ total,
                                                      int epi, double *__restrict__ out) {
    constexpr int CH = 128 / (int)sizeof(TRow);  // elements per 128-byte chunk
    constexpr int EPL = 16 / (int)sizeof(TRow);  // elements per lane and round (16 B)
    constexpr int PER = Widen<TRow>::PER;
    constexpr int QPL = (CH + 31) / 32;          // query elements a lane stages per chunk and slot
    static_assert(!QLANE || sizeof(TRow) == 8, "per-lane query staging: fp64 rows (query chunk = row chunk = 128 bytes)");
    __shared__ __align__(16) unsigned char s_raw[4][32][128 + 16];
    __shared__ __align__(16) double s_q[4][QLANE ? 1 : RS_QSLOTS][QLANE ? 2 : CH];
    __shared__ __align__(16) unsigned char s_rawq[QLANE ? 4 : 1][QLANE ? 32 : 1][128 + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npairs = *total;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p - lane >= npairs) return;  // whole warp past the end of the pair list
    const bool live = p < npairs;
    const int qi = live ? pair_query[p] : -1, ri = live ? pair_row[p] : -1;
    const double *qp = queries + (int64_t)max(qi, 0) * d;
    const int nch = (d + CH - 1) / CH;
    const bool vec = (((size_t)d * sizeof(TRow)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0);
    // the pairs of a query are consecutive: number the runs of equal queries in this warp; the chunks of the first
    // RS_QSLOTS runs are staged in shared memory with one coalesced load each, later runs read their query directly
    const int qprev = __shfl_up_sync(0xffffffffu, qi, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || qi != qprev);
    const int qslot = __popc(heads & ((2u << lane) - 1u)) - 1;
    const double *qsrc[RS_QSLOTS];
#pragma unroll
    for (int sidx = 0; sidx < RS_QSLOTS; ++sidx) {
        // lane of the sidx-th head
        unsigned h = heads;
        for (int t = 0; t < sidx; ++t) h &= h - 1;
        const int hl = h ? __ffs(h) - 1 : -1;
        const int qh = __shfl_sync(0xffffffffu, qi, max(hl, 0));
        qsrc[sidx] = (hl >= 0 && qh >= 0) ? queries + (int64_t)qh * d : nullptr;
    }
    // cooperative loads: in round r, lanes 8*(j%4) .. +7 fetch the 128-byte chunk of the warp's row j = 4*r + lane/8
    const int part = lane & 7;
    const TRow *src[8];
    const double *srcq[QLANE ? 8 : 1];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int rj = __shfl_sync(0xffffffffu, ri, r * 4 + (lane >> 3));
        src[r] = rj >= 0 ? rows + (int64_t)rj * d : nullptr;
        if (QLANE) {
            const int qj = __shfl_sync(0xffffffffu, qi, r * 4 + (lane >> 3));
            srcq[r] = qj >= 0 ? queries + (int64_t)qj * d : nullptr;
        }
    }
    const bool vecq = QLANE && (((size_t)d * 8) % 16 == 0) && ((reinterpret_cast<uintptr_t>(queries) & 15) == 0);
    uint4 pre[8];
    uint4 preq2[QLANE ? 8 : 1];
    double preq[QLANE ? 1 : RS_QSLOTS][QPL];
    auto fetch = [&](int c) {
        const int e0 = c * CH + part * EPL;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint4 val = make_uint4(0, 0, 0, 0);
            if (src[r] == nullptr) {
            } else if (vec && e0 + EPL <= d) val = __ldg(reinterpret_cast<const uint4 *>(src[r] + e0));
            else {
                alignas(16) TRow tmp[EPL];
#pragma unroll
                for (int e = 0; e < EPL; ++e) tmp[e] = (e0 + e < d) ? src[r][e0 + e] : TRow(0.0f);
                val = *reinterpret_cast<uint4 *>(tmp);
            }
            pre[r] = val;
        }
        if (QLANE) {
            const int q0e = c * CH + part * 2;  // two doubles per 16-byte piece
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                uint4 val = make_uint4(0, 0, 0, 0);
                if (srcq[r] == nullptr) {
                } else if (vecq && q0e + 2 <= d) val = __ldg(reinterpret_cast<const uint4 *>(srcq[r] + q0e));
                else {
                    alignas(16) double tmp[2];
                    tmp[0] = q0e < d ? srcq[r][q0e] : 0.0;
                    tmp[1] = q0e + 1 < d ? srcq[r][q0e + 1] : 0.0;
                    val = *reinterpret_cast<uint4 *>(tmp);
                }
                preq2[r] = val;
            }
            return;
        }
#pragma unroll
        for (int sidx = 0; sidx < RS_QSLOTS; ++sidx) {
#pragma unroll
            for (int t = 0; t < QPL; ++t) {
                const int i = t * 32 + lane;
                preq[sidx][t] = (qsrc[sidx] != nullptr && i < CH && c * CH + i < d) ? __ldg(qsrc[sidx] + c * CH + i) : 0.0;
            }
        }
    };
    fetch(0);
    double s = 0.0;
    for (int c = 0; c < nch; ++c) {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r) *reinterpret_cast<uint4 *>(&s_raw[warp][r * 4 + (lane >> 3)][part * 16]) = pre[r];
        if (QLANE) {
#pragma unroll
            for (int r = 0; r < 8; ++r) *reinterpret_cast<uint4 *>(&s_rawq[warp][r * 4 + (lane >> 3)][part * 16]) = preq2[r];
        } else {
#pragma unroll
            for (int sidx = 0; sidx < RS_QSLOTS; ++sidx) {
#pragma unroll
                for (int t = 0; t < QPL; ++t) {
                    const int i = t * 32 + lane;
                    if (i < CH) s_q[warp][sidx][i] = preq[sidx][t];
                }
            }
        }
        __syncwarp();
        if (c + 1 < nch) fetch(c + 1);  // in flight while this chunk is accumulated
        // own row chunk back as 128-bit reads (row stride 144 B: conflict-free per quarter warp)
        const uint4 *mine = reinterpret_cast<const uint4 *>(&s_raw[warp][lane][0]);
        const int kmax = min(CH, d - c * CH);
        if (live) {
            const bool staged = QLANE || qslot < RS_QSLOTS;
            const double *qs = QLANE ? reinterpret_cast<const double *>(&s_rawq[warp][lane][0])
                                     : (staged ? &s_q[warp][qslot][0] : qp + c * CH);
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const uint4 bits = mine[w];
#pragma unroll
                for (int e = 0; e < PER; e += 2) {
                    const int i0 = w * PER + e;
                    if (i0 < kmax) {
                        double q0, q1 = 0.0;
                        if (staged) {
                            const double2 qq = *reinterpret_cast<const double2 *>(qs + i0);  // zero-padded past d
                            q0 = qq.x;
                            q1 = qq.y;
                        } else {
                            q0 = qs[i0];
                            if (i0 + 1 < kmax) q1 = qs[i0 + 1];
                        }
                        s = mac_seq<ARITH>(q0, Widen<TRow>::at(bits, e), s);
                        if (i0 + 1 < kmax) s = mac_seq<ARITH>(q1, Widen<TRow>::at(bits, e + 1), s);
                    }
                }
            }
        }
    }
    if (live) out[pair_slot[p]] = apply_epi(epi, s, q_norm ? q_norm[qi] : 0.0, row_norm ? row_norm[ri] : 0.0);
}

template <typename T>
__global__ void widen_queries_kernel(const T *__restrict__ in, int64_t count, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = to_f64(in[i]);
}

__global__ void gather_bytes_kernel(const uint32_t *__restrict__ src, const int32_t *__restrict__ idx, int64_t n, int64_t row_words,
                                    uint32_t *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * row_words) return;
    const int64_t r = i / row_words, w = i % row_words;
    dst[i] = src[(int64_t)idx[r] * row_words + w];
}
__global__ void scatter_rows64_kernel(const uint64_t *__restrict__ src, const int32_t *__restrict__ idx, int64_t n, int k,
                                      uint64_t *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * k) return;
    const int64_t r = i / k, w = i % k;
    dst[(int64_t)idx[r] * k + w] = src[i];
}
__global__ void first_column_kernel(const int64_t *__restrict__ pos, int64_t nq, int stride, int64_t *__restrict__ first) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) first[q] = pos[q * stride];
}
// seed thresholds from the sample pass: the kk-th best candidate of the sample is a row, so >= kk rows score at least that
// ... and its k-th best candidate is a row too: a row more than `margin` below it can never enter the exact top-k
__global__ void thr_from_sample_kernel(const double *__restrict__ sel_negv, const int32_t *__restrict__ cnt, int64_t nq, int kk,
                                       int cap, int k, const float *__restrict__ margin, float *__restrict__ thr) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int c = cnt[q];
    if (c > cap) return;
    float t = thr[q];
    // -sel_negv is a float score widened to double: the conversion back is exact
    if (c >= kk) t = fmaxf(t, (float)(-sel_negv[q * kk + kk - 1]));
    if (k <= kk && c >= k) {
        const float sk = (float)(-sel_negv[q * kk + k - 1]);
        t = fmaxf(t, sk - margin[q] - 4e-6f * fabsf(sk));
    }
    thr[q] = t;
}
// ---- exact pruning of probed lists by the triangle inequality on angles ------------------------------------------------
// radius[l] = max over the rows r of list l of angle(r, centroid_l) (+ a rounding margin).  For any query q,
// angle(q, r) >= angle(q, c) - angle(r, c), so every row of list l has cosine similarity <= cos(angle(q, c_l) - radius[l])
// when the query is farther from the centroid than the list's radius.  One warp per slab row.
template <typename T>
__global__ void __launch_bounds__(256) list_radius_kernel(const T *__restrict__ slab, const double *__restrict__ slab_norm,
                                                          const double *__restrict__ cents, const double *__restrict__ cent_norm,
                                                          const int64_t *__restrict__ list_off, int nlist, int64_t n, int d,
                                                          unsigned long long *__restrict__ radius_bits) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    int lo = 0, hi = nlist;  // last l with list_off[l] <= r
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (list_off[mid] <= r) lo = mid; else hi = mid;
    }
    const T *x = slab + r * d;
    const double *c = cents + (int64_t)lo * d;
    double s = 0.0;
    for (int i = lane; i < d; i += 32) s += to_f64(x[i]) * c[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const double den = slab_norm[r] * cent_norm[lo];
        double ang = 3.141592653589794;  // zero / non-finite norms: the list is never pruned
        if (den > 0.0 && isfinite(den) && isfinite(s)) ang = acos(fmin(1.0, fmax(-1.0, s / den))) + 1e-7;
        atomicMax(&radius_bits[lo], (unsigned long long)__double_as_longlong(ang));  // positive doubles order like their bits
    }
}

// probe_pos[q][p] (list id, p >= 1) becomes -1 when no row of the list can reach the query's threshold:
//   cos(angle_lb(q, c) - radius) < thr[q] * q_scale[q] - q_eps[q],
// i.e. even the approximate score of every such row would lie below the threshold the candidate pass applies (and the
// proof of fast_final_kernel accounts for: rows "never emitted").  sim_ub = upper bound of cos(q, c) from the coarse stage.
__global__ void prune_probes_kernel(int64_t *__restrict__ probe_pos, const double *__restrict__ sim_ub,
                                    const double *__restrict__ radius, const float *__restrict__ thr,
                                    const double *__restrict__ q_scale, const double *__restrict__ q_eps, int64_t nq, int np,
                                    const int64_t *__restrict__ list_off, unsigned long long *__restrict__ pruned) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool cut = false;
    unsigned rows_cut = 0;
    if (i < nq * np && (i % np) != 0) {
        const int64_t q = i / np;
        const int64_t l = probe_pos[i];
        const float t = thr[q];
        if (l >= 0 && t > -INFINITY) {
            const double ub = fmin(1.0, fmax(-1.0, sim_ub[i]));
            const double a = acos(ub) - 1e-7;  // lower bound of angle(q, c)
            const double rad = radius[l];
            if (a > rad) {
                const double best = cos(a - rad) + 1e-9;  // no row of the list is more similar to q than this
                if (best < (double)t * q_scale[q] - q_eps[q]) {
                    probe_pos[i] = -1;
                    cut = true;
                    rows_cut = (unsigned)(list_off[l + 1] - list_off[l]);
                }
            }
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, cut);
    const unsigned rows = __reduce_add_sync(0xffffffffu, rows_cut);
    if ((threadIdx.x & 31) == 0 && b) {  // pruned[0] = (query, list) pairs, pruned[1] = (query, row) pairs
        atomicAdd(pruned, (unsigned long long)__popc(b));
        atomicAdd(pruned + 1, (unsigned long long)rows);
    }
}

// acc[0..n) += src[0..n)  (profiling counters kept on the device: no host round trip inside a search)
__global__ void accumulate_u64_kernel(unsigned long long *__restrict__ acc, const unsigned long long *__restrict__ src, int n) {
    if ((int)threadIdx.x < n) acc[threadIdx.x] += src[threadIdx.x];
}
// what a candidate pass covers: acc[0] += units, acc[1] += items (unit x row tile), acc[2] += row tiles of lists with units
// acc[3] += units whose selections fit 64 slots (run with M = 64: half of the query image is staged), counted if lq_off
__global__ void tc_cover_kernel(const int64_t *__restrict__ unit_prefix, const int64_t *__restrict__ tile_off,
                                const int64_t *__restrict__ lq_off, int nlist, unsigned long long *__restrict__ acc) {
    unsigned long long u = 0, it = 0, t = 0, h = 0, nw = 0, nwi = 0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < nlist; l += gridDim.x * blockDim.x) {
        const unsigned long long units = (unsigned long long)(unit_prefix[l + 1] - unit_prefix[l]);
        const unsigned long long tiles = (unsigned long long)(tile_off[l + 1] - tile_off[l]);
        u += units;
        it += units * tiles;
        if (units > 0) t += tiles;
        if (lq_off != nullptr && units > 0) {
            const int64_t last = (lq_off[l + 1] - lq_off[l]) - (int64_t)(units - 1) * kFastTile;  // selections of the list's last unit
            if (last <= 64) h += 1;
            if (last <= kNarrowSlots) nw += 1, nwi += tiles;  // acc[4], acc[5]: narrow units and their items
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        u += __shfl_xor_sync(0xffffffffu, u, o);
        it += __shfl_xor_sync(0xffffffffu, it, o);
        t += __shfl_xor_sync(0xffffffffu, t, o);
        h += __shfl_xor_sync(0xffffffffu, h, o);
        nw += __shfl_xor_sync(0xffffffffu, nw, o);
        nwi += __shfl_xor_sync(0xffffffffu, nwi, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc, u);
        atomicAdd(acc + 1, it);
        atomicAdd(acc + 2, t);
        atomicAdd(acc + 3, h);
        atomicAdd(acc + 4, nw);
        atomicAdd(acc + 5, nwi);
    }
}
// upper bound of the cosine similarity from an exact cosine distance (probe pruning needs sim(q, centroid) from above)
__global__ void sim_from_dist_kernel(const double *__restrict__ dist, int64_t n, double *__restrict__ sim) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sim[i] = (1.0 - dist[i]) + 1e-12;
}
__global__ void set_i64x4_kernel(int64_t *p, int64_t a, int64_t b, int64_t c, int64_t d) {
    p[0] = a, p[1] = b, p[2] = c, p[3] = d;
}
__global__ void and_flags_kernel(int32_t *__restrict__ ok, const int32_t *__restrict__ other, int64_t nq) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) ok[q] = ok[q] & other[q];
}

}  // namespace

void launch_cand_select(const double *cand_negv, const int32_t *cnt, int64_t nq, int kk, int cap, double *sel_negv,
                        int64_t *sel_pos) {
    if (nq == 0) return;
    HB_REQUIRE(cap <= 4096 && kk <= 128, "candidate select: cap <= 4096, kk <= 128");
    cand_select_warp_kernel<<<blocks_for(nq, 4), 128, 0, g_stream>>>(cand_negv, cnt, nq, kk, cap, sel_negv, sel_pos, CompactParams{});
    HB_LAUNCH_CHECK();
    cand_select_kernel<<<(unsigned)std::min<int64_t>(nq, (int64_t)g_num_sms * 8), 256, (size_t)cap * 8, g_stream>>>(
        cand_negv, cnt, nq, kk, cap, sel_negv, sel_pos, CompactParams{}, kWarpSelMax);
    HB_LAUNCH_CHECK();
}
void launch_dense_select(const float *dump, int ntiles, int64_t nq, int nrows, int k, int kk, int cap, const float *margin,
                         double *cand_negv, int32_t *cand_rel, int32_t *cand_pos, int32_t *cnt, float *thr, double *sel_negv,
                         int64_t *sel_pos) {
    if (nq == 0) return;
    HB_REQUIRE(nrows <= 2048 && kk <= 128 && k >= 1, "dense select: at most 2048 rows, kk <= 128");
    const int grid = blocks_for(nq, 4);
#define HB_DS(NPL_) dense_select_kernel<NPL_><<<grid, 128, 0, g_stream>>>(dump, ntiles, nq, nrows, k, kk, cap, margin, cand_negv, cand_rel, cand_pos, cnt, thr, sel_negv, sel_pos)
    if (nrows <= 512) HB_DS(16);
    else if (nrows <= 1024) HB_DS(32);
    else HB_DS(64);
#undef HB_DS
    HB_LAUNCH_CHECK();
}
void launch_cand_compact(double *cand_negv, int32_t *cand_rel, int32_t *cand_pos, int32_t *cnt, int64_t nq, int kk, int cap, int k,
                         const float *margin, float *thr, double *sel_negv, int64_t *sel_pos) {
    if (nq == 0) return;
    HB_REQUIRE(cap <= 4096 && kk <= 128, "candidate compaction: cap <= 4096, kk <= 128");
    CompactParams CP;
    CP.cand_negv = cand_negv;
    CP.cand_rel = cand_rel;
    CP.cand_pos = cand_pos;
    CP.cnt_rw = cnt;
    CP.thr = thr;
    CP.margin = margin;
    CP.k = k;
    // the warp kernel first: it lowers cnt to <= kk for the queries it compacts, so the block kernel skips them
    cand_select_warp_kernel<<<blocks_for(nq, 4), 128, 0, g_stream>>>(cand_negv, cnt, nq, kk, cap, sel_negv, sel_pos, CP);
    HB_LAUNCH_CHECK();
    cand_select_kernel<<<(unsigned)std::min<int64_t>(nq, (int64_t)g_num_sms * 8), 256, (size_t)cap * 8, g_stream>>>(
        cand_negv, cnt, nq, kk, cap, sel_negv, sel_pos, CP, kWarpSelMax);
    HB_LAUNCH_CHECK();
}

namespace {
template <typename TRow>
void rescore_arith(const void *rows, const double *row_norm, const double *queries, bool q_f32_repr, const double *q_norm, int d,
                   const int32_t *pq, const int32_t *pr, const int32_t *ps, const int32_t *total, int64_t max_pairs, int epi,
                   double *out, bool sparse) {
    const int grid = blocks_for(max_pairs, 128);
    if constexpr (sizeof(TRow) == 8) {
        if (sparse) {  // (fp64 rows are never fp32-representable products: MULADD)
            rescore_kernel<TRow, ARITH_MULADD, true><<<grid, 128, 0, g_stream>>>((const TRow *)rows, row_norm, queries, q_norm, d, pq, pr,
                                                                                  ps, total, epi, out);
            HB_LAUNCH_CHECK();
            return;
        }
    }
    // both factors fp32-representable: the fp64 product is exact, DFMA rounds like multiply-then-add
    if (is_f32_repr<TRow>::value && q_f32_repr)
        rescore_kernel<TRow, ARITH_FMA><<<grid, 128, 0, g_stream>>>((const TRow *)rows, row_norm, queries, q_norm, d, pq, pr, ps, total,
                                                                     epi, out);
    else
        rescore_kernel<TRow, ARITH_MULADD><<<grid, 128, 0, g_stream>>>((const TRow *)rows, row_norm, queries, q_norm, d, pq, pr, ps,
                                                                        total, epi, out);
    HB_LAUNCH_CHECK();
}
}  // namespace

// queries64: the queries widened to fp64 (launch_widen_queries); q_f32_repr: they hold fp32-representable values.
// The dense pair list (pair_query / pair_row / pair_slot, *total entries, at most max_pairs) comes from
// launch_rescore_pairs; out[pair_slot] receives the distance.
void launch_rescore(const void *rows, int rdtype, const double *row_norm, const double *queries64, bool q_f32_repr,
                    const double *q_norm, int d, const int32_t *pair_query, const int32_t *pair_row, const int32_t *pair_slot,
                    const int32_t *total, int64_t max_pairs, int epi, double *out, bool sparse) {
    if (max_pairs == 0) return;
#define HB_RS(T_) rescore_arith<T_>(rows, row_norm, queries64, q_f32_repr, q_norm, d, pair_query, pair_row, pair_slot, total, max_pairs, epi, out, sparse)
    if (rdtype == HB_F32) HB_RS(float);
    else if (rdtype == HB_BF16) HB_RS(__nv_bfloat16);
    else if (rdtype == HB_F64) HB_RS(double);
    else throw Error(HB_ERR_INVALID, "unsupported dtype for the exact re-score");
#undef HB_RS
}

// fp64 copy of the queries for the re-score (returns `queries` itself when they already are fp64)
const double *launch_widen_queries(const void *queries, int qdtype, int64_t count, double *buf) {
    if (qdtype == HB_F64) return (const double *)queries;
    if (count == 0) return buf;
    if (qdtype == HB_F32) widen_queries_kernel<float><<<blocks_for(count, 256), 256, 0, g_stream>>>((const float *)queries, count, buf);
    else throw Error(HB_ERR_INVALID, "queries must be fp32 or fp64");
    HB_LAUNCH_CHECK();
    return buf;
}

namespace {
__global__ void i32_to_i64_kernel(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
}  // namespace
void launch_i32_to_i64(const int32_t *in, int64_t n, int64_t *out) {
    if (n == 0) return;
    i32_to_i64_kernel<<<blocks_for(n, 256), 256, 0, g_stream>>>(in, n, out);
    HB_LAUNCH_CHECK();
}

void launch_gather_bytes(const void *src, const int32_t *idx, int64_t n, int64_t row_bytes, void *dst) {
    if (n == 0) return;
    HB_REQUIRE(row_bytes % 4 == 0, "gather rows must be whole 32-bit words");
    gather_bytes_kernel<<<blocks_for(n * (row_bytes / 4), 256), 256, 0, g_stream>>>((const uint32_t *)src, idx, n, row_bytes / 4,
                                                                                     (uint32_t *)dst);
    HB_LAUNCH_CHECK();
}
void launch_scatter_rows64(const void *src, const int32_t *idx, int64_t n, int k, void *dst) {
    if (n * k == 0) return;
    scatter_rows64_kernel<<<blocks_for(n * k, 256), 256, 0, g_stream>>>((const uint64_t *)src, idx, n, k, (uint64_t *)dst);
    HB_LAUNCH_CHECK();
}
void launch_first_column(const int64_t *pos, int64_t nq, int stride, int64_t *first) {
    if (nq == 0) return;
    first_column_kernel<<<blocks_for(nq, 256), 256, 0, g_stream>>>(pos, nq, stride, first);
    HB_LAUNCH_CHECK();
}
void launch_thr_from_sample(const double *sel_negv, const int32_t *cnt, int64_t nq, int kk, int cap, int k, const float *margin,
                            float *thr) {
    if (nq == 0) return;
    thr_from_sample_kernel<<<blocks_for(nq, 256), 256, 0, g_stream>>>(sel_negv, cnt, nq, kk, cap, k, margin, thr);
    HB_LAUNCH_CHECK();
}
void launch_list_radius(const void *slab, int dtype, const double *slab_norm, const double *cents, const double *cent_norm,
                        const int64_t *list_off, int nlist, int64_t n, int d, double *radius) {
    HB_CUDA(cudaMemsetAsync(radius, 0, (size_t)nlist * 8, g_stream));
    if (n == 0) return;
    const int grid = blocks_for(n * 32, 256);
    unsigned long long *rb = reinterpret_cast<unsigned long long *>(radius);
    if (dtype == HB_F32) list_radius_kernel<float><<<grid, 256, 0, g_stream>>>((const float *)slab, slab_norm, cents, cent_norm, list_off, nlist, n, d, rb);
    else if (dtype == HB_BF16) list_radius_kernel<__nv_bfloat16><<<grid, 256, 0, g_stream>>>((const __nv_bfloat16 *)slab, slab_norm, cents, cent_norm, list_off, nlist, n, d, rb);
    else list_radius_kernel<double><<<grid, 256, 0, g_stream>>>((const double *)slab, slab_norm, cents, cent_norm, list_off, nlist, n, d, rb);
    HB_LAUNCH_CHECK();
}
void launch_prune_probes(int64_t *probe_pos, const double *sim_ub, const double *radius, const float *thr, const double *q_scale,
                         const double *q_eps, int64_t nq, int np, const int64_t *list_off, unsigned long long *pruned) {
    if (nq * np == 0) return;
    prune_probes_kernel<<<blocks_for(nq * np, 256), 256, 0, g_stream>>>(probe_pos, sim_ub, radius, thr, q_scale, q_eps, nq, np, list_off,
                                                                        pruned);
    HB_LAUNCH_CHECK();
}
void launch_accumulate_u64(unsigned long long *acc, const unsigned long long *src, int n) {
    accumulate_u64_kernel<<<1, 32, 0, g_stream>>>(acc, src, n);
    HB_LAUNCH_CHECK();
}
void launch_tc_cover(const int64_t *unit_prefix, const int64_t *tile_off, const int64_t *lq_off, int nlist,
                     unsigned long long *acc) {
    if (nlist == 0) return;
    tc_cover_kernel<<<blocks_for(nlist, 256), 256, 0, g_stream>>>(unit_prefix, tile_off, lq_off, nlist, acc);
    HB_LAUNCH_CHECK();
}
// gathered[g] = [per x np positions (int64) | per x np distances (fp64)] of rank g's query block -> ppos / simub of all nq queries
__global__ void unpack_probe_blocks_kernel(const char *__restrict__ gathered, int64_t per, int64_t nq, int np,
                                           int64_t *__restrict__ ppos, double *__restrict__ simub) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * np) return;
    const int64_t q = i / np, g = q / per, o = (q - g * per) * np + i % np;
    const char *blk = gathered + (size_t)g * (size_t)per * np * 16;
    ppos[i] = reinterpret_cast<const int64_t *>(blk)[o];
    simub[i] = (1.0 - reinterpret_cast<const double *>(blk + (size_t)per * np * 8)[o]) + 1e-12;
}
void launch_unpack_probe_blocks(const void *gathered, int nranks, int64_t per, int64_t nq, int np, int64_t *ppos, double *simub) {
    if (nq * np == 0) return;
    (void)nranks;
    unpack_probe_blocks_kernel<<<blocks_for(nq * np, 256), 256, 0, g_stream>>>((const char *)gathered, per, nq, np, ppos, simub);
    HB_LAUNCH_CHECK();
}
// idx[0 .. count) = the queries with ok[q] == 0, ascending; idx[nq] = count (one block: ordered, and nq is a few thousand)
__global__ void __launch_bounds__(1024) compact_failed_kernel(const int32_t *__restrict__ ok, int64_t nq, int32_t *__restrict__ idx) {
    __shared__ int s_w[32];
    __shared__ int s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int64_t base = 0; base < nq; base += 1024) {
        const int64_t q = base + threadIdx.x;
        const bool bad = q < nq && ok[q] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, bad);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = s_run;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (bad) idx[before + __popc(m & ((1u << lane) - 1u))] = (int32_t)q;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += s_w[w];
            s_run += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) idx[nq] = s_run;
}
void launch_compact_failed(const int32_t *ok, int64_t nq, int32_t *idx) {
    compact_failed_kernel<<<1, 1024, 0, g_stream>>>(ok, nq, idx);
    HB_LAUNCH_CHECK();
}
void launch_sim_from_dist(const double *dist, int64_t n, double *sim) {
    if (n == 0) return;
    sim_from_dist_kernel<<<blocks_for(n, 256), 256, 0, g_stream>>>(dist, n, sim);
    HB_LAUNCH_CHECK();
}
void launch_set_i64x4(int64_t *p, int64_t a, int64_t b, int64_t c, int64_t d) {
    set_i64x4_kernel<<<1, 1, 0, g_stream>>>(p, a, b, c, d);
    HB_LAUNCH_CHECK();
}
void launch_and_flags(int32_t *ok, const int32_t *other, int64_t nq) {
    if (nq == 0) return;
    and_flags_kernel<<<blocks_for(nq, 256), 256, 0, g_stream>>>(ok, other, nq);
    HB_LAUNCH_CHECK();
}

void launch_tile_offsets(const int64_t *list_off, int nlist, int64_t *tile_off) {
    tile_offsets_kernel<<<1, 32, 0, g_stream>>>(list_off, nlist, tile_off);
    HB_LAUNCH_CHECK();
}

void launch_quant_rows(const void *rows, int dtype, int d, int kbn, int ns, int nlist, const int64_t *list_off,
                       const int64_t *tile_off, const double *norm, int8_t *img, float *rs, float *ro, float *stats) {
    // caller has zeroed img/rs/stats and filled ro with -inf for the padded tile positions
    int64_t n = 0;
    HB_CUDA(cudaMemcpyAsync(&n, list_off + nlist, 8, cudaMemcpyDeviceToHost, g_stream));
    HB_CUDA(cudaStreamSynchronize(g_stream));
    if (n == 0) return;
    const int grid = blocks_for(n * 32, 256);
#define HB_QR(T, NS_) quant_rows_kernel<T, NS_><<<grid, 256, 0, g_stream>>>((const T *)rows, n, d, kbn, nlist, list_off, tile_off, norm, img, rs, ro, stats)
    if (ns == 2) {
        if (dtype == HB_F32) HB_QR(float, 2);
        else if (dtype == HB_BF16) HB_QR(__nv_bfloat16, 2);
        else HB_QR(double, 2);
    } else {
        if (dtype == HB_F32) HB_QR(float, 3);
        else if (dtype == HB_BF16) HB_QR(__nv_bfloat16, 3);
        else HB_QR(double, 3);
    }
#undef HB_QR
    HB_LAUNCH_CHECK();
}

void launch_fill_f32(float *p, int64_t n, float v) {
    if (n == 0) return;
    fill_f32_kernel<<<blocks_for(n, 256), 256, 0, g_stream>>>(p, n, v);
    HB_LAUNCH_CHECK();
}

void launch_quant_queries(const void *queries, int qdtype, int64_t nq, int d, int kbn, int ns, int8_t *dig, double *qu,
                          double *ql1) {
    if (nq == 0) return;
    const int grid = blocks_for(nq * 32, 256);
#define HB_QQ(T, NS_) quant_queries_kernel<T, NS_><<<grid, 256, 0, g_stream>>>((const T *)queries, nq, d, kbn, dig, qu, ql1)
    if (ns == 2) {
        if (qdtype == HB_F32) HB_QQ(float, 2);
        else HB_QQ(double, 2);
    } else {
        if (qdtype == HB_F32) HB_QQ(float, 3);
        else HB_QQ(double, 3);
    }
#undef HB_QQ
    HB_LAUNCH_CHECK();
}

void launch_unit_plan(int nlist, const int64_t *lq_off, const int64_t *unit_prefix, const int64_t *tile_off, int nunits,
                      int nunits_real, int interleave, int tile_limit, int tile_div, int tile_start, const int32_t *qsel,
                      const int64_t *pair_out, int pair_div, const int32_t *pair_query, UnitPlan U, int skip_tiles,
                      unsigned long long *stats) {
    if (nunits == 0) return;
    unit_plan_kernel<<<blocks_for(nunits + 1, 256), 256, 0, g_stream>>>(nlist, lq_off, unit_prefix, tile_off, nunits, nunits_real,
                                                                        interleave, tile_limit, tile_div, tile_start, U, skip_tiles, qsel,
                                                                        pair_div, stats);
    HB_LAUNCH_CHECK();
    unit_scan_kernel<<<1, 1024, 0, g_stream>>>(U.unit_ntile, nunits + 1, U.unit_item0, U.unit_nsel, U.unit_item0n);
    HB_LAUNCH_CHECK();
    unit_slots_kernel<<<blocks_for((int64_t)nunits * 32, 256), 256, 0, g_stream>>>(
        nunits, U.unit_sel0, U.unit_nsel, qsel, pair_out, pair_div, pair_query, U.slot_query, U.slot_rel0, U.unit_item0n != nullptr);
    HB_LAUNCH_CHECK();
}

void launch_pack_units(const int8_t *dig, int kbn, int ns, int nunits, const int32_t *slot_query, int8_t *aimg,
                       const int32_t *unit_nsel, bool narrow) {
    if (nunits == 0) return;
    const int grid = (int)std::min<int64_t>((int64_t)nunits * kbn, (int64_t)g_num_sms * 32);
    if (ns == 2) pack_units_kernel<2><<<grid, 256, 0, g_stream>>>(dig, kbn, nunits, slot_query, aimg, unit_nsel, narrow);
    else pack_units_kernel<3><<<grid, 256, 0, g_stream>>>(dig, kbn, nunits, slot_query, aimg, unit_nsel, narrow);
    HB_LAUNCH_CHECK();
}

void launch_query_bounds(const double *qu, const double *ql1, const double *qnorm, int64_t nq, int ns, int d, int metric,
                         const float *stats, double *q_scale, double *q_eps, float *q_margin, float *thr_init, int32_t *cnt_init) {
    if (nq == 0) return;
    query_bounds_kernel<<<blocks_for(nq, 256), 256, 0, g_stream>>>(qu, ql1, qnorm, nq, ns, d, metric, stats, q_scale, q_eps,
                                                                   q_margin, thr_init, cnt_init);
    HB_LAUNCH_CHECK();
}

void launch_rescore_pairs(const int64_t *sel_pos, const double *sel_negv, const int32_t *cand_pos, int64_t nq, int kk, int cap,
                          int k, const float *margin, int32_t *slot_row, double *exact, int32_t *total, int32_t *pair_query,
                          int32_t *pair_row, int32_t *pair_slot, bool set_only) {
    if (nq * kk == 0) return;
    HB_REQUIRE(kk <= 128, "re-score pairs: kk <= 128");
    HB_CUDA(cudaMemsetAsync(total, 0, 4, g_stream));
    rescore_pairs_kernel<<<blocks_for(nq * 32, 256), 256, 0, g_stream>>>(sel_pos, sel_negv, cand_pos, nq, kk, cap, k, margin, slot_row,
                                                                         exact, total, pair_query, pair_row, pair_slot, set_only ? 1 : 0);
    HB_LAUNCH_CHECK();
}

void launch_fast_final(const FinalParams &P) {
    if (P.nq == 0) return;
    HB_REQUIRE(P.kk <= 128 && P.k <= P.kk, "fast final: k <= kk <= 128");
    fast_final_kernel<<<blocks_for(P.nq, 4), 128, 0, g_stream>>>(P);
    HB_LAUNCH_CHECK();
}

}  // namespace hb
