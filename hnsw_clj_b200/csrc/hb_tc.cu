// hb_tc.cu — the tensor-core candidate pass of HB_MODE_FAST (sm_100a: tcgen05 + TMEM + bulk async copies).
//
// Replaces, for candidate selection only, the per-pair scoring loops of the reference
// (src/hnsw/ann/partition/ivf_flat.clj:217-234 list scan, :261-269 coarse routing, src/hnsw/bench.clj:72-84 flat
// scan): every (query slot, row) score of a 128 x 128 tile is an exact integer dot product
//      S = sum_i mq_i * mr_i      (mq, mr = block-fixed-point mantissas, NS signed 8-bit digits each)
// evaluated as NS*NS (NS = 2) or 6 (NS = 3, the two lowest-weight digit products dropped) int8 GEMMs that
// accumulate into three int32 TMEM accumulators, one per digit weight 2^16, 2^8, 2^0.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      producer: one lane streams the pre-swizzled digit images of A (query unit) and B (row tile)
//               into a ring of shared-memory stages with cp.async.bulk (UBLKCP), completion on mbarriers;
//   warp 1      MMA issuer: one lane issues tcgen05.mma.kind::i8 (M = N = 128, K = 32) from shared-memory
//               descriptors (K-major, SWIZZLE_128B), commits to the stage's "empty" barrier and, after the
//               last k-block, to the "accumulators full" barrier; also owns the TMEM allocation;
//   warps 2..5  epilogue: thread = query slot = TMEM lane; tcgen05.ld 16 columns (rows) at a time, combine
//               the three digit-weight accumulators, apply the per-row scale (1/norm folded in) and append
//               rows at or above the query's running threshold to its candidate list, while 32 two-deep
//               buckets keep raising that threshold (it always has >= 64 emitted rows at or above it).
//               The score matrix never goes to HBM.
// Work = items (unit, row tile); item ranges are split evenly across CTAs, consecutive items share the unit so
// the A images stay hot in L2.
#include <float.h>
#include <math.h>

#include "hb_fast.cuh"

namespace hb {
namespace {

constexpr int TC_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int NB = 32;  // threshold buckets (two best scores each: >= 64 rows at or above the bucket minimum)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, M = 128, N = 128, K = 32
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor: D = S32 (2 @bit4), A = B = signed int8 (1 @bit7, 1 @bit10), K-major both, N>>3 @bit17, M>>4 @bit24
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kFastTile >> 3) << 17) | ((uint32_t)(kFastTile >> 4) << 24);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

template <int NS>
struct Prod;
template <>
struct Prod<2> {  // (a digit, b digit): weight class = sum of digit indices
    static constexpr int N = 4;
    __device__ static constexpr int sa(int p) { return p == 0 ? 0 : p == 1 ? 0 : p == 2 ? 1 : 1; }
    __device__ static constexpr int sb(int p) { return p == 0 ? 0 : p == 1 ? 1 : p == 2 ? 0 : 1; }
    __device__ static constexpr bool first(int p) { return p == 0 || p == 1 || p == 3; }
};
template <>
struct Prod<3> {  // classes 0..2 of 0..4; (1,2), (2,1), (2,2) are dropped and bounded (hb_fastprep.cu)
    static constexpr int N = 6;
    __device__ static constexpr int sa(int p) { return p == 0 ? 0 : p == 1 ? 0 : p == 2 ? 1 : p == 3 ? 0 : p == 4 ? 1 : 2; }
    __device__ static constexpr int sb(int p) { return p == 0 ? 0 : p == 1 ? 1 : p == 2 ? 0 : p == 3 ? 2 : p == 4 ? 1 : 0; }
    __device__ static constexpr bool first(int p) { return p == 0 || p == 1 || p == 3; }
};

template <int NS>
struct TcCfg {
    static constexpr int STAGE_BYTES = 2 * NS * kFastImg;
    static constexpr int NSTAGE = NS == 2 ? 3 : 2;
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024;
};

__device__ __forceinline__ int find_unit(const int32_t *__restrict__ item0, int nunits, int item) {
    int lo = 0, hi = nunits;  // last u with item0[u] <= item
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (item0[mid] <= item) lo = mid;
        else hi = mid;
    }
    return lo;
}

template <int NS, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_pass_kernel(const TcParams P) {
    using Cfg = TcCfg<NS>;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t s_full[Cfg::NSTAGE];
    __shared__ __align__(8) uint64_t s_empty[Cfg::NSTAGE];
    __shared__ __align__(8) uint64_t s_tmem_full;
    __shared__ __align__(8) uint64_t s_tmem_empty;
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_rs[2][kFastTile];
    __shared__ float s_ro[2][kFastTile];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage0 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const int kbn = P.kbn;

    const int total_items = P.unit_item0[P.nunits];
    const int item_begin = (int)((int64_t)total_items * blockIdx.x / gridDim.x);
    const int item_end = (int)((int64_t)total_items * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::NSTAGE; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        mbar_init(smem_u32(&s_tmem_full), 1);
        mbar_init(smem_u32(&s_tmem_empty), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0 && item_begin < item_end) {
            int stage = 0;
            uint32_t phase = 0;
            int u = find_unit(P.unit_item0, P.nunits, item_begin);
            for (int item = item_begin; item < item_end; ++item) {
                while (item >= P.unit_item0[u + 1]) ++u;
                const int t = (item - P.unit_item0[u]) * P.tile_stride;
                const int64_t btile = P.tile_off[P.unit_list[u]] + t;
                const int8_t *asrc = P.aimg + (int64_t)u * kbn * NS * kFastImg;
                const int8_t *bsrc = P.bimg + btile * kbn * NS * kFastImg;
                for (int kb = 0; kb < kbn; ++kb) {
                    mbar_wait(smem_u32(&s_empty[stage]), phase ^ 1u);
                    const uint32_t full = smem_u32(&s_full[stage]);
                    const uint32_t dst = stage0 + (uint32_t)stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(full, Cfg::STAGE_BYTES);
                    bulk_g2s(dst, asrc + (int64_t)kb * NS * kFastImg, NS * kFastImg, full);
                    bulk_g2s(dst + NS * kFastImg, bsrc + (int64_t)kb * NS * kFastImg, NS * kFastImg, full);
                    if (++stage == Cfg::NSTAGE) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && item_begin < item_end) {
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int item = item_begin; item < item_end; ++item) {
                mbar_wait(smem_u32(&s_tmem_empty), tphase ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < kbn; ++kb) {
                    mbar_wait(smem_u32(&s_full[stage]), phase);
                    tc_fence_after();
                    const uint32_t abase = stage0 + (uint32_t)stage * Cfg::STAGE_BYTES;
                    const uint32_t bbase = abase + NS * kFastImg;
#pragma unroll
                    for (int k4 = 0; k4 < kFastKB / 32; ++k4) {
#pragma unroll
                        for (int p = 0; p < Prod<NS>::N; ++p) {
                            const int sa = Prod<NS>::sa(p), sb = Prod<NS>::sb(p);
                            const uint64_t ad = smem_desc(abase + sa * kFastImg + k4 * 32);
                            const uint64_t bd = smem_desc(bbase + sb * kFastImg + k4 * 32);
                            const uint32_t acc = (kb == 0 && k4 == 0 && Prod<NS>::first(p)) ? 0u : 1u;
                            umma_i8(tmem_base + (uint32_t)(sa + sb) * kFastTile, ad, bd, kIdescI8, acc);
                        }
                    }
                    umma_commit(smem_u32(&s_empty[stage]));
                    if (++stage == Cfg::NSTAGE) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(smem_u32(&s_tmem_full));
                tphase ^= 1u;
            }
        }
    } else {
        // ===== epilogue: thread = query slot = TMEM lane =====
        const int quarter = warp & 3;  // TMEM lanes this warp may touch: 32*quarter ..
        const int slot = quarter * 32 + lane;
        const int et = (warp - 2) * 32 + lane;  // 0..127, for cooperative loads
        uint32_t tphase = 0;
        int u = item_begin < item_end ? find_unit(P.unit_item0, P.nunits, item_begin) : 0;
        int qi = -1, rel0 = 0;
        float thr = INFINITY;
        // Running lower bound of the query's 64th best score: bucket b holds the two best scores among the rows
        // with (row % NB == b) seen in this unit, so min_b(b2) has at least 2*NB = 64 rows at or above it.
        float b1[NB], b2[NB];
        int it = 0;
        for (int item = item_begin; item < item_end; ++item, ++it) {
            while (item >= P.unit_item0[u + 1]) ++u;
            const int t = (item - P.unit_item0[u]) * P.tile_stride;
            const int ntu = P.unit_item0[u + 1] - P.unit_item0[u];
            const bool last_of_unit = (item - P.unit_item0[u]) == ntu - 1 || item == item_end - 1;
            if (item == item_begin || item == P.unit_item0[u]) {
                qi = P.slot_query[(int64_t)u * kFastTile + slot];
                rel0 = P.slot_rel0[(int64_t)u * kFastTile + slot];
                if (MODE == FAST_EMIT) thr = qi >= 0 ? P.thr[qi] : INFINITY;
                if (MODE != FAST_DUMP) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) b1[b] = b2[b] = -INFINITY;
                }
            }
            const int l = P.unit_list[u];
            const int64_t btile = P.tile_off[l] + t;
            const int64_t row0 = P.list_off[l] + (int64_t)t * kFastTile;
            const int par = it & 1;
            s_rs[par][et] = P.rs[btile * kFastTile + et];
            s_ro[par][et] = P.ro[btile * kFastTile + et];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(smem_u32(&s_tmem_full), tphase);
            tc_fence_after();
            tphase ^= 1u;
            const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
            // 4 x 32 columns; NOT fully unrolled: the body must stay inside the instruction cache (one warp per
            // scheduler cannot hide instruction-fetch misses).  Within a 32-column step the bucket of column
            // (h*16 + j) is a compile-time register.
#pragma unroll 1
            for (int c32 = 0; c32 < kFastTile / 32; ++c32) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col0 = c32 * 32 + h * 16;
                    uint32_t c0[16], c1[16], c2[16];
                    tmem_ld16(tlane + col0, c0);
                    tmem_ld16(tlane + kFastTile + col0, c1);
                    tmem_ld16(tlane + 2 * kFastTile + col0, c2);
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float s = fmaf((float)(int)c0[j], 65536.0f, fmaf((float)(int)c1[j], 256.0f, (float)(int)c2[j]));
                        v[j] = fmaf(s, s_rs[par][col0 + j], s_ro[par][col0 + j]);
                    }
                    if (MODE != FAST_DUMP) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int b = h * 16 + j;
                            const float lo = fminf(b1[b], v[j]);
                            b1[b] = fmaxf(b1[b], v[j]);
                            b2[b] = fmaxf(b2[b], lo);
                        }
                    }
                    if (MODE == FAST_EMIT) {
                        uint32_t mask = 0;
                        const float cut = fmaxf(thr, -FLT_MAX);  // padding rows score -inf: never candidates
#pragma unroll
                        for (int j = 0; j < 16; ++j) mask |= (v[j] >= cut ? 1u : 0u) << j;
                        if (mask) {
                            int base = atomicAdd(&P.cnt[qi], __popc(mask));
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if ((mask >> j) & 1u) {
                                    if (base < P.cap) {
                                        const int64_t o = (int64_t)qi * P.cap + base;
                                        P.cand_negv[o] = -(double)v[j];
                                        P.cand_rel[o] = rel0 + t * kFastTile + col0 + j;
                                        P.cand_pos[o] = (int32_t)(row0 + col0 + j);
                                    }
                                    ++base;
                                }
                            }
                        }
                    } else if (MODE == FAST_DUMP) {
                        float *o = P.dump + ((int64_t)item * kFastTile + slot) * kFastTile + col0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = v[j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_tmem_empty));
            if (MODE != FAST_DUMP && qi >= 0) {
                // every row counted in the buckets was emitted (EMIT) under a threshold <= the new one, so raising
                // thr keeps "at least 64 candidates at or above thr" true
                float m = b2[0];
#pragma unroll
                for (int b = 1; b < NB; ++b) m = fminf(m, b2[b]);
                thr = fmaxf(thr, m);
                if (last_of_unit) {
                    if (thr > -INFINITY) atomic_max_float(&P.thr[qi], thr);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

template <int NS, int MODE>
void tc_launch(const TcParams &P, int total_items) {
    auto kernel = tc_pass_kernel<NS, MODE>;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<NS>::SMEM_BYTES));
    const int grid = std::max(1, std::min(g_num_sms, total_items));
    kernel<<<grid, TC_THREADS, TcCfg<NS>::SMEM_BYTES, g_stream>>>(P);
    HB_LAUNCH_CHECK();
}

}  // namespace

// total_items is only used to size the grid; the kernel reads the exact count from unit_item0[nunits]
void launch_tc_pass(const TcParams &P, int ns, int mode) {
    if (P.nunits == 0) return;
    HB_REQUIRE(ns == 2 || ns == 3, "digit count must be 2 or 3");
    const int hint = P.nunits * 8;
#define HB_TC(NS_)                                                        \
    do {                                                                  \
        if (mode == FAST_EMIT) tc_launch<NS_, FAST_EMIT>(P, hint);        \
        else tc_launch<NS_, FAST_DUMP>(P, hint);                          \
    } while (0)
    if (ns == 2) HB_TC(2);
    else HB_TC(3);
#undef HB_TC
}

}  // namespace hb
