// hb_kpp.cu — kmeans-plus-plus-init (src/hnsw/ann/partition/ivf_flat.clj:32-60; Lightning's d_i-weighted variant,
// lightning.clj:86-109) at scale, bit for bit.
//
// One step of the reference: d_i = min(d_i, distance-fn(x_i, newest seed)) for every row (:45-49); S = sum d_i^2 in row
// order (:51-52, a sequential fp64 sum); r = nextDouble * S; the pick is the first i whose running sum reaches r (:54-58).
// Both halves look inherently serial / exhaustive.  They are not:
//
// (1) The ordered sum.  While the running sum c stays inside one binade [2^e, 2^(e+1)), c is a multiple of U = ulp = 2^(e-52)
//     and fl(c + x) = c + U * rint_even(x / U) — an INTEGER addition unless x / U falls exactly half-way (a tie, decided by
//     the parity of c / U).  Rows are cut into chunks of 512.  An approximate prefix (any order; the ordered sum differs from
//     it by < 2^-24 relative for n < 2^28 non-negative terms) tells for each chunk whether all its partial sums provably lie in
//     one binade; if so, and no element is a tie, the chunk maps c -> c + U * Q with Q = sum rint(x_i / U) computed in
//     parallel with exact integer adds.  The ~log2(n / 512) chunks that straddle a power of two, hold a tie, or start from a
//     sum outside the assumed binade are walked with real fp64 adds by one thread.  The composition over chunks is
//     sequential but one integer add per chunk.  The result is the reference's S and running sums exactly, not approximately
//     (tests/test_gpu_kpp.py: adversarial ties, cancellation-free ranges, 2^k boundaries vs a sequential loop).
// (2) The exhaustive distance pass.  Angles obey the triangle inequality: with theta_i the angle between row i and its nearest
//     seed s(i), a new seed c can only be nearer if angle(c, s(i)) < 2 theta_i.  One warp per earlier seed computes
//     angle(c, seed_j); a row whose bound fails by a margin (1e-4 rad, far above the 1e-13 of the arithmetic) is skipped
//     without reading its 3 KB — its d_i provably keeps its bits.  Euclidean: dist(c, s(i)) >= 2 d_i.  On clustered data
//     almost every row is skipped once every cluster holds a seed; on structureless data nothing is, and the pass is the
//     HBM-bound scan it was.
#include <float.h>
#include <limits.h>
#include <math.h>

#include "hb_build.cuh"

namespace hb {
namespace {

constexpr int KC = kKppChunk;  // rows per chunk
constexpr double kDelta = 1.0 / 16777216.0;  // 2^-24: bound of |ordered sum - approximate sum| / sum, n < 2^28

// ---- (2) bounds between the newest seed and every earlier one: one warp per earlier seed ------------------------------------
template <typename T, bool L2>
__global__ void __launch_bounds__(128) kpp_seed_bounds_kernel(const T *__restrict__ rows, int d, const double *__restrict__ row_norm,
                                                              const int64_t *__restrict__ seeds, int t, double *__restrict__ bound) {
    const int j = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= t) return;
    const int64_t c = seeds[t - 1], s = seeds[j];
    const T *x = rows + c * (int64_t)d, *y = rows + s * (int64_t)d;
    double acc = 0.0;
    for (int k = lane; k < d; k += 32) {
        const double a = to_f64(x[k]), b = to_f64(y[k]);
        acc += L2 ? (a - b) * (a - b) : a * b;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane != 0) return;
    double out;
    if (L2) {
        out = sqrt(acc) * (1.0 - 1e-9);  // a lower bound of dist(c, seed_j)
    } else {
        const double nc = row_norm[c], ns = row_norm[s];
        // a zero-norm seed has distance 1.0 to everything (the guard of cosine-distance-ultra): no angle, no pruning
        if (!(nc > 0.0) || !(ns > 0.0)) out = -INFINITY;
        else out = acos(fmin(1.0, fmax(-1.0, acc / (nc * ns)))) - 1e-6;  // a lower bound of angle(c, seed_j)
    }
    bound[j] = j == t - 1 ? 0.0 : out;
}

// d_i = min(d_i, distance-fn(x_i, newest seed)): one thread per row, the reference's sequential sum — for the rows the
// triangle inequality cannot exclude.  near / theta: the seed (by pick order) that gave d_i and the row's bound towards it.
// The rows of a block travel through shared memory in chunks of 32 elements: a warp reads 32 consecutive elements of one row
// per load (whole 128-byte lines) and the owner thread walks its row's chunk in order from a transposed, conflict-free tile;
// the loads of the next chunk are in flight (32 registers per lane) while the current one is consumed.  Measured on 262,144
// x 768 fp32 rows, 70 % of them scored: one thread per row reading rows[i * d + k] (lanes 3 KB apart) 0.8 TB/s; staged
// without the prefetch 2.3 TB/s; 512-byte chunks (three blocks per SM) 1.2 TB/s.  On structureless data (nothing to prune
// until every cluster holds a seed) this pass IS the seeding time.
constexpr int KU_CH = 32;           // elements per staged chunk
constexpr int KU_LD = 129;          // tile row stride (elements): (k * 129 + r) mod 32 distinct over r and over k
template <typename T, bool L2>
__global__ void __launch_bounds__(128, 5) kpp_update_pruned_kernel(const T *__restrict__ rows, int64_t n, int d,
                                                                const double *__restrict__ row_norm, const int64_t *__restrict__ seeds,
                                                                int t, const double *__restrict__ bound, double *__restrict__ mind,
                                                                int32_t *__restrict__ near, float *__restrict__ theta,
                                                                unsigned long long *__restrict__ n_scored) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    double *s_c = reinterpret_cast<double *>(s_raw);                       // the newest seed, widened: [d]
    T *s_x = reinterpret_cast<T *>(s_raw + (size_t)d * sizeof(double));   // [KU_CH][KU_LD]
    __shared__ unsigned char s_act[128];
    const int64_t i0 = (int64_t)blockIdx.x * 128, i = i0 + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    bool active = false;
    if (i < n) {
        const float th = theta[i];
        const double b = bound[near[i]];
        active = !(L2 ? (b >= 2.0 * (double)th) : (b >= 2.0 * (double)th + 1e-4));  // else: the newest seed cannot be nearer
    }
    s_act[threadIdx.x] = active ? 1 : 0;
    const int nact = __syncthreads_count(active);
    if (nact == 0) return;
    if (n_scored && threadIdx.x == 0) atomicAdd(n_scored, (unsigned long long)nact);
    const int64_t c = seeds[t - 1];
    const T *cv = rows + c * d;
    for (int k = threadIdx.x; k < d; k += 128) s_c[k] = to_f64(cv[k]);
    double s = 0.0;
    const T *blk = rows + i0 * d;
    T v[32];  // this warp's rows warp, warp + 4, ...: element k0 + lane of each
    auto fetch = [&](int k0) {
        const int kn = min(KU_CH, d - k0);
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            const int r = m * 4 + warp;
            if (s_act[r] && lane < kn) v[m] = blk[(uint32_t)(r * d + k0 + lane)];  // 32-bit offsets: 128 rows of the block
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < d; k0 += KU_CH) {
        const int kn = min(KU_CH, d - k0);
        __syncthreads();  // the previous chunk is consumed (and s_c is complete)
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            const int r = m * 4 + warp;
            if (s_act[r] && lane < kn) s_x[lane * KU_LD + r] = v[m];
        }
        __syncthreads();
        if (k0 + KU_CH < d) fetch(k0 + KU_CH);
        if (active) {
#pragma unroll 8
            for (int kk = 0; kk < kn; ++kk) {
                const double a = to_f64(s_x[kk * KU_LD + threadIdx.x]), bb = s_c[k0 + kk];
                if (L2) s = mac_seq<ARITH_L2>(bb, a, s);
                else s = mac_seq<is_f32_repr<T>::value ? ARITH_FMA : ARITH_MULADD>(a, bb, s);
            }
        }
    }
    if (!active) return;
    const double ni = L2 ? 0.0 : row_norm[i], nc = L2 ? 0.0 : row_norm[c];
    const double dist = L2 ? __dsqrt_rn(s) : apply_epi(EPI_COS_GUARD, s, ni, nc);
    if (dist < mind[i]) {
        mind[i] = dist;
        near[i] = t - 1;
        float nt;
        if (L2) {
            nt = __double2float_ru(dist * (1.0 + 1e-6));
        } else if (!(ni > 0.0)) {
            nt = -INFINITY;  // a zero row is at 1.0 from every seed for ever: never needs another look
        } else if (!(nc > 0.0)) {
            nt = INFINITY;   // 1.0 by the guard, not by an angle: no bound
        } else {
            nt = __double2float_ru(acos(fmin(1.0, fmax(-1.0, 1.0 - dist))) + 1e-6);
        }
        theta[i] = nt;
    }
}

// ---- (1) the ordered sum -----------------------------------------------------------------------------------------------------
template <bool SQ>
__device__ __forceinline__ double weight_of(double v) { return SQ ? __dmul_rn(v, v) : v; }

// approximate sum of each chunk (fixed tree: deterministic)
template <bool SQ>
__global__ void __launch_bounds__(128) kpp_chunk_approx_kernel(const double *__restrict__ mind, int64_t n, double *__restrict__ approx) {
    __shared__ double s_w[4];
    const int64_t base = (int64_t)blockIdx.x * KC;
    double a = 0.0;
    for (int j = threadIdx.x; j < KC; j += 128) {
        const int64_t i = base + j;
        if (i < n) a += weight_of<SQ>(mind[i]);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) approx[blockIdx.x] = (s_w[0] + s_w[1]) + (s_w[2] + s_w[3]);
}

// exclusive prefix of the chunk sums (one block; any fixed order will do)
__global__ void __launch_bounds__(1024) kpp_chunk_scan_kernel(const double *__restrict__ approx, int nchunks, double *__restrict__ prefix) {
    __shared__ double s_part[1024];
    __shared__ double s_run;
    if (threadIdx.x == 0) s_run = 0.0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024) {
        const int i = base + threadIdx.x;
        const double v = i < nchunks ? approx[i] : 0.0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const double tt = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0.0;
            __syncthreads();
            s_part[threadIdx.x] += tt;
            __syncthreads();
        }
        if (i < nchunks) prefix[i] = s_run + s_part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_run += s_part[1023];
        __syncthreads();
    }
}

// per chunk: the binade all its partial sums provably lie in (INT_MIN: unknown) and Q = sum rint(x_i / ulp)
template <bool SQ>
__global__ void __launch_bounds__(128) kpp_chunk_exact_kernel(const double *__restrict__ mind, int64_t n, const double *__restrict__ approx,
                                                              const double *__restrict__ prefix, int *__restrict__ c_exp,
                                                              long long *__restrict__ c_q) {
    __shared__ long long s_q[4];
    __shared__ int s_tie[4];
    const int k = blockIdx.x;
    const double lo = prefix[k] * (1.0 - kDelta), hi = (prefix[k] + approx[k]) * (1.0 + kDelta);
    int e = INT_MIN;
    if (lo > 0.0 && hi < DBL_MAX && lo >= DBL_MIN * 4503599627370496.0) {  // away from the subnormal range: ulp = 2^(e-52) is a normal number
        const int elo = ilogb(lo), ehi = ilogb(hi);
        if (elo == ehi) e = elo;
    }
    long long q = 0;
    int tie = 0;
    if (e != INT_MIN) {
        const int64_t base = (int64_t)k * KC;
        for (int j = threadIdx.x; j < KC; j += 128) {
            const int64_t i = base + j;
            if (i >= n) continue;
            const double x = weight_of<SQ>(mind[i]);
            if (!(x >= 0.0) || !(x < DBL_MAX)) {  // NaN / inf / negative: not this machinery's business
                tie = 1;
                continue;
            }
            const double sc = scalbn(x, 52 - e);  // x / ulp, exact (x < 2^(e+1) keeps it below 2^53; tiny x may flush to a tie-free 0)
            const double fl = floor(sc);
            const double fr = sc - fl;            // exact
            if (sc >= 9007199254740992.0) tie = 1;  // cannot happen when the bounds hold; be safe
            q += (long long)fl + (fr > 0.5 ? 1 : 0);
            if (fr == 0.5) tie = 1;
            // a subnormal-tiny x whose scaling underflowed inexactly could hide a tie: x / ulp < 2^-1000 is never 0.5
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        q += __shfl_xor_sync(0xffffffffu, q, m);
        tie |= __shfl_xor_sync(0xffffffffu, tie, m);
    }
    if ((threadIdx.x & 31) == 0) {
        s_q[threadIdx.x >> 5] = q;
        s_tie[threadIdx.x >> 5] = tie;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int any = s_tie[0] | s_tie[1] | s_tie[2] | s_tie[3];
        c_exp[k] = any ? INT_MIN : e;
        c_q[k] = s_q[0] + s_q[1] + s_q[2] + s_q[3];
    }
}

// Composition over the chunks, S, r = u * S and the pick: the first i with fl(cum_i + x_i) >= r (ivf_flat.clj:54-58), clamped
// to n - 1.  Thread 0 composes (one integer add per chunk; the chunk records are staged through shared memory by the block);
// a chunk that needs real fp64 adds is first copied into shared memory by the whole block (one coalesced read instead of 512
// dependent global loads), then walked by thread 0.
template <bool SQ>
__global__ void __launch_bounds__(256) kpp_compose_pick_kernel(const double *__restrict__ mind, int64_t n, int nchunks,
                                                               const int *__restrict__ c_exp, const long long *__restrict__ c_q,
                                                               double *__restrict__ c_start, const double *__restrict__ u,
                                                               double *__restrict__ total, int64_t *__restrict__ pick,
                                                               unsigned long long *__restrict__ n_walked) {
    constexpr int TILE = 2048;
    __shared__ int s_e[TILE];
    __shared__ long long s_qq[TILE];
    __shared__ double s_cs[TILE];
    __shared__ double s_w[KC];
    __shared__ double s_cum, s_r;
    __shared__ int s_k, s_next, s_need;
    __shared__ unsigned long long s_walk;
    if (threadIdx.x == 0) {
        s_cum = 0.0;
        s_walk = 0;
    }
    for (int base = 0; base < nchunks; base += TILE) {
        const int m = min(TILE, nchunks - base);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            s_e[j] = c_exp[base + j];
            s_qq[j] = c_q[base + j];
        }
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        while (true) {
            if (threadIdx.x == 0) {
                double cum = s_cum;
                int j = s_next;
                s_need = -1;
                for (; j < m; ++j) {
                    const int e = s_e[j];
                    s_cs[j] = cum;
                    // cum in the chunk's binade 2^e (a normal number: kpp_chunk_exact_kernel keeps e away from the subnormals):
                    // cum / ulp is its 53-bit significand, and adding Q ulps is an integer add on the bit pattern as long as the
                    // significand stays below 2^53 (the sum stays in the binade)
                    const long long bits = __double_as_longlong(cum);
                    if (e != INT_MIN && bits > 0 && (int)((bits >> 52) & 0x7ff) - 1023 == e) {
                        const long long sig = (bits & 0x000fffffffffffffll) | 0x0010000000000000ll;
                        const long long q = s_qq[j];
                        if (q >= 0 && sig + q < 9007199254740992ll) {
                            cum = __longlong_as_double(bits + q);
                            continue;
                        }
                    }
                    s_need = base + j;  // real fp64 adds in row order
                    break;
                }
                s_cum = cum;
                s_next = j + 1;
            }
            __syncthreads();
            const int need = s_need;
            if (need < 0) break;
            const int64_t b = (int64_t)need * KC;
            const int cnt = (int)min((int64_t)KC, n - b);
            for (int j = threadIdx.x; j < cnt; j += blockDim.x) s_w[j] = weight_of<SQ>(mind[b + j]);
            __syncthreads();
            if (threadIdx.x == 0) {
                double cum = s_cum;
#pragma unroll 8
                for (int j = 0; j < cnt; ++j) cum = __dadd_rn(cum, s_w[j]);
                s_cum = cum;
                ++s_walk;
            }
            __syncthreads();
        }
        for (int j = threadIdx.x; j < m; j += blockDim.x) c_start[base + j] = s_cs[j];  // the running sum at each chunk's start
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        c_start[nchunks] = s_cum;
        *total = s_cum;
        s_r = __dmul_rn(*u, s_cum);
        s_k = nchunks;  // none
    }
    __syncthreads();
    const double r = s_r;
    // first chunk whose END value reaches r (the running sums never decrease)
    for (int k = threadIdx.x; k < nchunks; k += blockDim.x)
        if (c_start[k + 1] >= r) {
            atomicMin(&s_k, k);
            break;
        }
    __syncthreads();
    const int kp = s_k;
    if (kp < nchunks) {
        const int64_t b = (int64_t)kp * KC;
        const int cnt = (int)min((int64_t)KC, n - b);
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) s_w[j] = weight_of<SQ>(mind[b + j]);
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t p = n - 1;
            double cum = c_start[kp];
            for (int j = 0; j < cnt; ++j) {
                cum = __dadd_rn(cum, s_w[j]);
                if (cum >= r) {
                    p = b + j;
                    break;
                }
            }
            *pick = p;
        }
    } else if (threadIdx.x == 0) {
        *pick = n - 1;
    }
    if (threadIdx.x == 0 && n_walked) atomicAdd(n_walked, s_walk);
}

__global__ void kpp_record_pick_kernel(const int64_t *pick, int64_t *out_seed) { *out_seed = *pick; }

__global__ void kpp_init_state_kernel(int64_t n, double *mind, int32_t *near, float *theta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mind[i] = DBL_MAX;  // (reduce min Double/MAX_VALUE ...), ivf_flat.clj:47
    near[i] = 0;
    theta[i] = INFINITY;
}

inline int blocks_for(int64_t n, int bs) { return (int)ceil_div(n > 0 ? n : 1, bs); }

}  // namespace

void launch_kpp_init_state(int64_t n, double *mind, int32_t *near, float *theta) {
    kpp_init_state_kernel<<<blocks_for(n, 256), 256, 0, g_stream>>>(n, mind, near, theta);
    HB_LAUNCH_CHECK();
}

template <bool SQ>
static void kpp_sum_pick(const KppScaleParams &P) {
    const int nchunks = (int)ceil_div(P.n, KC);
    kpp_chunk_approx_kernel<SQ><<<nchunks, 128, 0, g_stream>>>(P.mind, P.n, P.c_approx);
    HB_LAUNCH_CHECK();
    kpp_chunk_scan_kernel<<<1, 1024, 0, g_stream>>>(P.c_approx, nchunks, P.c_prefix);
    HB_LAUNCH_CHECK();
    kpp_chunk_exact_kernel<SQ><<<nchunks, 128, 0, g_stream>>>(P.mind, P.n, P.c_approx, P.c_prefix, P.c_exp, P.c_q);
    HB_LAUNCH_CHECK();
    kpp_compose_pick_kernel<SQ><<<1, 256, 0, g_stream>>>(P.mind, P.n, nchunks, P.c_exp, P.c_q, P.c_start, P.u, P.total, P.pick, P.n_walked);
    HB_LAUNCH_CHECK();
}

// r = u * S and the pick for given weights (already in `mind`; linear = the weights themselves, else their squares)
void launch_kpp_sum_pick(const KppScaleParams &P) {
    HB_REQUIRE(P.n >= 1 && P.n < (1ll << 28) * (int64_t)KC, "k-means++: too many rows for the chunked sum");
    if (P.linear) kpp_sum_pick<false>(P);
    else kpp_sum_pick<true>(P);
}

void launch_kpp_scale_step(const KppScaleParams &P) {
    HB_REQUIRE(P.t >= 1, "k-means++ step needs a seed");
    const int grid_b = blocks_for((int64_t)P.t * 32, 128), grid_u = blocks_for(P.n, 128);
#define HB_KPPS(T)                                                                                                               \
    do {                                                                                                                         \
        const size_t smem_u = (size_t)P.d * sizeof(double) + (size_t)KU_CH * KU_LD * sizeof(T);                                  \
        HB_REQUIRE(smem_u <= 200 * 1024, "k-means++: dimension too large for the update kernel");                                \
        if (smem_u > 48 * 1024) {                                                                                                \
            if (P.l2) HB_CUDA(cudaFuncSetAttribute(kpp_update_pruned_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));  \
            else HB_CUDA(cudaFuncSetAttribute(kpp_update_pruned_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));      \
        }                                                                                                                        \
        if (P.l2) {                                                                                                              \
            kpp_seed_bounds_kernel<T, true><<<grid_b, 128, 0, g_stream>>>((const T *)P.rows, P.d, P.row_norm, P.seeds, P.t, P.bound);   \
            HB_LAUNCH_CHECK();                                                                                                   \
            kpp_update_pruned_kernel<T, true><<<grid_u, 128, smem_u, g_stream>>>((const T *)P.rows, P.n, P.d, P.row_norm, P.seeds, P.t,      \
                                                                            P.bound, P.mind, P.near, P.theta, P.n_scored);        \
        } else {                                                                                                                 \
            kpp_seed_bounds_kernel<T, false><<<grid_b, 128, 0, g_stream>>>((const T *)P.rows, P.d, P.row_norm, P.seeds, P.t, P.bound);  \
            HB_LAUNCH_CHECK();                                                                                                   \
            kpp_update_pruned_kernel<T, false><<<grid_u, 128, smem_u, g_stream>>>((const T *)P.rows, P.n, P.d, P.row_norm, P.seeds, P.t,     \
                                                                             P.bound, P.mind, P.near, P.theta, P.n_scored);       \
        }                                                                                                                        \
        HB_LAUNCH_CHECK();                                                                                                       \
    } while (0)
    if (P.dtype == HB_F32) HB_KPPS(float);
    else if (P.dtype == HB_BF16) HB_KPPS(__nv_bfloat16);
    else HB_KPPS(double);
#undef HB_KPPS
    launch_kpp_sum_pick(P);
    kpp_record_pick_kernel<<<1, 1, 0, g_stream>>>(P.pick, P.seeds_rw + P.t);
    HB_LAUNCH_CHECK();
}

}  // namespace hb
