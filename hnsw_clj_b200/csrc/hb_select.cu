// hb_select.cu — segmented top-k selection by (distance, position).
//
// Replaces the reference's "compute every distance, stable full sort, take k"
// (src/hnsw/simd_optimized.clj:271-280; src/hnsw/bench.clj:72-84; Collections/sort + take at
// src/hnsw/ann/partition/ivf_flat.clj:229-234 and sort-by/take at :291-294).  A stable sort on distance
// followed by take-k equals the k smallest under the lexicographic order (distance, position in the
// sorted sequence), and every caller lays its candidates out in the reference's pre-sort order, so
// position is the tie-break.
//
// One CTA per segment streams the candidates once (coalesced 8-byte loads), keeps the current k-th
// best as a threshold, appends survivors to a shared-memory buffer and bitonic-sorts buffer + best
// whenever the buffer fills.  After warm-up almost nothing survives, so the kernel is a pure read
// stream: HBM-bound at 8 B per candidate.
#include "hb_kernels.cuh"

namespace hb {
namespace {

constexpr int SNT = 128;  // threads per CTA
constexpr int SU = 4;     // candidates per thread per step

__device__ __forceinline__ bool kp_less(uint64_t ka, uint32_t pa, uint64_t kb, uint32_t pb) {
    return ka < kb || (ka == kb && pa < pb);
}

template <int CAP>
__device__ void bitonic_sort(uint64_t *keys, uint32_t *pos, int tid) {
    for (int size = 2; size <= CAP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = tid; i < CAP / 2; i += SNT) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const uint64_t ka = keys[lo], kb = keys[hi];
                const uint32_t pa = pos[lo], pb = pos[hi];
                const bool swap = asc ? kp_less(kb, pb, ka, pa) : kp_less(ka, pa, kb, pb);
                if (swap) {
                    keys[lo] = kb;
                    keys[hi] = ka;
                    pos[lo] = pb;
                    pos[hi] = pa;
                }
            }
        }
    }
    __syncthreads();
}

template <int CAP>
__global__ void __launch_bounds__(SNT) select_kernel(const SelectParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *pos = reinterpret_cast<uint32_t *>(keys + CAP);
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    const int k = P.k;

    for (int64_t w = blockIdx.x; w < P.nseg * P.nsub; w += gridDim.x) {
        const int64_t s = w / P.nsub;
        const int64_t sub0 = (w - s * P.nsub) * P.sub_len;  // first candidate of this sub-range
        const int64_t sbegin = P.seg_off ? P.seg_off[s * P.seg_off_stride] : s * P.seg_stride;
        const int64_t slen = P.seg_off ? P.seg_off[(s + 1) * P.seg_off_stride] - sbegin : P.seg_len_const;
        const int64_t len = P.nsub > 1 ? max((int64_t)0, min(P.sub_len, slen - sub0)) : slen;
        const double *v = P.vals + sbegin + sub0;
        const int64_t oslot = s * P.out_seg_stride + P.out_slot_base + (w - s * P.nsub);
        __syncthreads();
        if (tid == 0) s_cnt = 0;
        int nbest = 0;  // block-uniform
        int ub = 0;     // block-uniform upper bound on s_cnt
        uint64_t thr = kKeyEmpty;
        __syncthreads();

        for (int64_t base = 0; base < len; base += SNT * SU) {
            if (ub + SNT * SU > CAP) {  // no room for a full step: fold the buffer into the best-k
                const int filled = s_cnt;  // stable: barrier at the end of the previous step
                __syncthreads();
                for (int i = filled + tid; i < CAP; i += SNT) {
                    keys[i] = kKeyEmpty;
                    pos[i] = 0xFFFFFFFFu;
                }
                bitonic_sort<CAP>(keys, pos, tid);
                nbest = min(k, filled);
                if (nbest == k) thr = keys[k - 1];
                __syncthreads();
                if (tid == 0) s_cnt = nbest;
                ub = nbest;
                __syncthreads();
            }
            uint64_t kk[SU];
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const int64_t i = base + u * SNT + tid;
                kk[u] = i < len ? dist_key(v[i]) : kKeyEmpty;
            }
            bool any = false;
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const int64_t i = base + u * SNT + tid;
                if (i < len && (nbest < k || kk[u] < thr)) {
                    const int slot = atomicAdd(&s_cnt, 1);
                    keys[slot] = kk[u];
                    pos[slot] = (uint32_t)i;
                    any = true;
                }
            }
            ub += SU * __syncthreads_count(any);
        }
        {
            const int filled = s_cnt;
            __syncthreads();
            for (int i = filled + tid; i < CAP; i += SNT) {
                keys[i] = kKeyEmpty;
                pos[i] = 0xFFFFFFFFu;
            }
            bitonic_sort<CAP>(keys, pos, tid);
            nbest = min(k, filled);
        }
        for (int j = tid; j < k; j += SNT) {
            const bool ok = j < nbest;
            P.out_val[oslot * k + j] = ok ? key_dist(keys[j]) : __longlong_as_double(0x7FF0000000000000ll);
            P.out_pos[oslot * k + j] = ok ? sub0 + (int64_t)pos[j] : -1;
        }
    }
}

// ---- short ranges, few of them (small batches: one or a few queries per call) ------------------------------------------
// select_kernel above is a streaming kernel: per 512 candidates two barriers and, while its buffer warms up, 1024-wide
// bitonic sorts -- fine when thousands of segments run side by side, 40-200 us when one query is all there is.  Here a CTA
// holds its whole (sub-)range of <= kSmallSelMax keys in shared memory, finds the k-th smallest key by bisection over
// the 64 key bits (block-wide counts, no sorting, no atomics on hot addresses), collects the keys below it plus the
// first ties in position order, and sorts just those k by (key, position).  Same result as select_kernel by construction:
// the k smallest under (key, position).
constexpr int kSmallSelMax = 4096;
constexpr int SSN = 256;

__device__ __forceinline__ int block_sum_256(int v, int *s_red) {  // every thread gets the sum; two barriers
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < SSN / 32; ++w) t += s_red[w];
    return t;
}

__global__ void __launch_bounds__(SSN) select_small_kernel(const SelectParams P) {
    __shared__ uint64_t s_key[kSmallSelMax];
    __shared__ uint64_t s_wkey[1024];
    __shared__ uint32_t s_wpos[1024];
    __shared__ int s_red[SSN / 32];
    __shared__ int s_n, s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = P.k;
    const int64_t w = blockIdx.x;
    const int64_t s = w / P.nsub;
    const int64_t sub0 = (w - s * P.nsub) * P.sub_len;
    const int64_t sbegin = P.seg_off ? P.seg_off[s * P.seg_off_stride] : s * P.seg_stride;
    const int64_t slen = P.seg_off ? P.seg_off[(s + 1) * P.seg_off_stride] - sbegin : P.seg_len_const;
    const int len = (int)(P.nsub > 1 ? max((int64_t)0, min(P.sub_len, slen - sub0)) : slen);
    const double *v = P.vals + sbegin + sub0;
    const int64_t oslot = s * P.out_seg_stride + P.out_slot_base + (w - s * P.nsub);
    for (int i = tid; i < len; i += SSN) s_key[i] = dist_key(v[i]);
    if (tid == 0) s_n = 0, s_run = 0;
    __syncthreads();
    const int nbest = min(k, len);
    uint64_t T = kKeyEmpty;  // len <= k: everything wins
    int need = 0;            // ties (key == T) to take, first positions first
    if (len > k) {
        uint64_t p = 0;
        for (int b = 63; b >= 0; --b) {  // the largest p with #{key < p} < k is the k-th smallest key
            const uint64_t cand = p | (1ull << b);
            int c = 0;
            for (int i = tid; i < len; i += SSN) c += s_key[i] < cand;
            if (block_sum_256(c, s_red) < k) p = cand;
        }
        T = p;
        int c = 0;
        for (int i = tid; i < len; i += SSN) c += s_key[i] < T;
        need = k - block_sum_256(c, s_red);
    }
    for (int i = tid; i < len; i += SSN) {
        const uint64_t key = s_key[i];
        if (key < T || len <= k) {
            const int slot = atomicAdd(&s_n, 1);
            s_wkey[slot] = key;
            s_wpos[slot] = (uint32_t)i;
        }
    }
    __syncthreads();
    for (int base = 0; base < len && need > 0; base += SSN) {  // block-uniform loop: ties in position order
        const int i = base + tid;
        const bool tie = i < len && s_key[i] == T;
        const unsigned m = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) s_red[warp] = __popc(m);
        __syncthreads();
        int before = s_run, total = 0;
#pragma unroll
        for (int x = 0; x < SSN / 32; ++x) {
            if (x < warp) before += s_red[x];
            total += s_red[x];
        }
        const int rank = before + __popc(m & ((1u << lane) - 1u));
        if (tie && rank < need) {
            const int slot = atomicAdd(&s_n, 1);
            s_wkey[slot] = T;
            s_wpos[slot] = (uint32_t)i;
        }
        __syncthreads();
        if (tid == 0) s_run += total;
        __syncthreads();
        if (s_run >= need) break;
    }
    __syncthreads();
    // sort the nbest winners by (key, position): bitonic over the next power of two
    int m2 = 1;
    while (m2 < nbest) m2 <<= 1;
    for (int i = nbest + tid; i < m2; i += SSN) {
        s_wkey[i] = kKeyEmpty;
        s_wpos[i] = 0xFFFFFFFFu;
    }
    for (int size = 2; size <= m2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = tid; i < m2 / 2; i += SSN) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const uint64_t ka = s_wkey[lo], kb = s_wkey[hi];
                const uint32_t pa = s_wpos[lo], pb = s_wpos[hi];
                const bool swap = asc ? kp_less(kb, pb, ka, pa) : kp_less(ka, pa, kb, pb);
                if (swap) {
                    s_wkey[lo] = kb;
                    s_wkey[hi] = ka;
                    s_wpos[lo] = pb;
                    s_wpos[hi] = pa;
                }
            }
        }
    }
    __syncthreads();
    for (int j = tid; j < k; j += SSN) {
        const bool ok = j < nbest;
        P.out_val[oslot * k + j] = ok ? key_dist(s_wkey[j]) : __longlong_as_double(0x7FF0000000000000ll);
        P.out_pos[oslot * k + j] = ok ? sub0 + (int64_t)s_wpos[j] : -1;
    }
}

}  // namespace

int select_small_max() { return kSmallSelMax; }

void launch_select(const SelectParams &P) {
    if (P.nseg == 0 || P.k == 0) return;
    HB_REQUIRE(P.k <= 1024, "k > 1024 is not supported by the device top-k");
    HB_REQUIRE(P.nsub >= 1, "nsub must be >= 1");
    {
        // few, short (sub-)ranges: the whole range in shared memory, k-th key by bisection (select_small_kernel)
        const int64_t max_len = P.nsub > 1 ? P.sub_len : (P.seg_off ? INT64_MAX : P.seg_len_const);
        if (max_len <= kSmallSelMax && P.nseg * P.nsub <= (int64_t)g_num_sms * 4) {
            select_small_kernel<<<(unsigned)(P.nseg * P.nsub), SSN, 0, g_stream>>>(P);
            HB_LAUNCH_CHECK();
            return;
        }
    }
    const int grid = (int)std::min<int64_t>(P.nseg * P.nsub, (int64_t)g_num_sms * 16);
    if (P.k <= 256) {
        constexpr int CAP = 1024;
        select_kernel<CAP><<<grid, SNT, CAP * 12, g_stream>>>(P);
    } else {
        constexpr int CAP = 4096;
        HB_CUDA(cudaFuncSetAttribute(select_kernel<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, CAP * 12));
        select_kernel<CAP><<<grid, SNT, CAP * 12, g_stream>>>(P);
    }
    HB_LAUNCH_CHECK();
}

}  // namespace hb
