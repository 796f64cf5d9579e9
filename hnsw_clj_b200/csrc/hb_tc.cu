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
//   warps 2..9  epilogue: thread = (query slot = TMEM lane, half of the tile's columns); tcgen05.ld 16 columns
//               (rows) at a time, combine the three digit-weight accumulators, apply the per-row scale (1/norm
//               folded in) and stash rows at or above the query's running threshold in shared memory (appended
//               to the query's candidate list with one atomic per batch), while 2 x 16 two-deep buckets per
//               slot keep raising that threshold (it always has >= 64 emitted rows at or above it).
//               The score matrix never goes to HBM.
// Work = items (unit, row tile); item ranges are split evenly across CTAs, consecutive items share the unit so
// the A images stay hot in L2.
#include <float.h>
#include <math.h>

#include "hb_fast.cuh"

namespace hb {
namespace {

constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int TC_THREADS = 64 + EPI_THREADS;
constexpr int TMEM_COLS = 512;
// threshold buckets per thread (two best scores each; two threads per query slot): NB = 16 bounds the 64th best score
// (64 candidates re-scored), NB = 32 the 128th (128 re-scored, k > 48)
constexpr int STASH = 8;   // candidates a thread keeps in shared memory before it appends them to the query's list

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
// the same with M = 64: a unit with <= 64 query slots in use.  The accumulator rows then sit in TMEM lanes 0-15 of each
// 32-lane quadrant (row r -> lane 32 * (r / 16) + r % 16), the A operand is the first 8 KB of each 16 KB image.
constexpr uint32_t kIdescI8M64 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kFastTile >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);

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

// Appends a thread's stashed candidates (score, position inside the list) to its query's candidate list, at
// `base_in` if the slots were reserved beforehand, else behind one atomic reservation.
// Kept out of line: the epilogue must stay small.
struct CandOut {
    int32_t *cnt;
    double *negv;
    int32_t *rel, *pos;
    int cap;
};
__device__ __noinline__ void flush_stash(const CandOut C, int qi, int rel0, int64_t list_row0, int n, int base_in,
                                         const uint2 *stash) {
    int base = base_in >= 0 ? base_in : atomicAdd(&C.cnt[qi], n);
    for (int i = 0; i < n; ++i, ++base) {
        if (base < C.cap) {
            const uint2 e = stash[i * EPI_THREADS];
            const int64_t o = (int64_t)qi * C.cap + base;
            C.negv[o] = -(double)__uint_as_float(e.x);
            C.rel[o] = rel0 + (int)e.y;
            C.pos[o] = (int32_t)(list_row0 + (int)e.y);
        }
    }
}

template <int NS, int MODE, int NB>
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
    __shared__ float s_m[2][2][2][kFastTile];              // [item parity][second / first best][column half][slot]: bucket minima
    __shared__ __align__(8) uint2 s_stash[STASH][EPI_THREADS];  // per-thread pending candidates (score bits, list position)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage0 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const int kbn = P.kbn;

    const int total_items = P.unit_item0[P.nunits];
    if (total_items == 0) return;  // (every unit of the pass was narrow: nothing to set up, 24 us saved per launch)
    const int item_begin = (int)((int64_t)total_items * blockIdx.x / gridDim.x);
    const int item_end = (int)((int64_t)total_items * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::NSTAGE; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        mbar_init(smem_u32(&s_tmem_full), 1);
        mbar_init(smem_u32(&s_tmem_empty), EPI_WARPS);
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
                const int t = (P.unit_tile0 ? P.unit_tile0[u] : P.tile_start) + (item - P.unit_item0[u]) * P.tile_stride;
                const int64_t btile = P.tile_off[P.unit_list[u]] + t;
                const int8_t *asrc = P.aimg + (int64_t)u * kbn * NS * kFastImg;
                const int8_t *bsrc = P.bimg + btile * kbn * NS * kFastImg;
                const bool hm = MODE == FAST_EMIT && P.unit_nsel != nullptr && P.unit_nsel[u] <= 64;
                for (int kb = 0; kb < kbn; ++kb) {
                    mbar_wait(smem_u32(&s_empty[stage]), phase ^ 1u);
                    const uint32_t full = smem_u32(&s_full[stage]);
                    const uint32_t dst = stage0 + (uint32_t)stage * Cfg::STAGE_BYTES;
                    if (hm) {  // M = 64: the first 64 slots = the first 8 row groups (8 KB) of each digit image
                        mbar_expect_tx(full, NS * (kFastImg / 2) + NS * kFastImg);
#pragma unroll
                        for (int sl = 0; sl < NS; ++sl)
                            bulk_g2s(dst + sl * kFastImg, asrc + ((int64_t)kb * NS + sl) * kFastImg, kFastImg / 2, full);
                    } else {
                        mbar_expect_tx(full, Cfg::STAGE_BYTES);
                        bulk_g2s(dst, asrc + (int64_t)kb * NS * kFastImg, NS * kFastImg, full);
                    }
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
            long long t_tmem = 0, t_full = 0;
            const long long t_all0 = clock64();
            int u = find_unit(P.unit_item0, P.nunits, item_begin);
            for (int item = item_begin; item < item_end; ++item) {
                while (item >= P.unit_item0[u + 1]) ++u;
                const uint32_t idesc = (MODE == FAST_EMIT && P.unit_nsel != nullptr && P.unit_nsel[u] <= 64) ? kIdescI8M64 : kIdescI8;
                long long c0 = P.timing ? clock64() : 0;
                mbar_wait(smem_u32(&s_tmem_empty), tphase ^ 1u);
                if (P.timing) t_tmem += clock64() - c0;
                tc_fence_after();
                for (int kb = 0; kb < kbn; ++kb) {
                    c0 = P.timing ? clock64() : 0;
                    mbar_wait(smem_u32(&s_full[stage]), phase);
                    if (P.timing) t_full += clock64() - c0;
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
                            umma_i8(tmem_base + (uint32_t)(sa + sb) * kFastTile, ad, bd, idesc, acc);
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
            if (P.timing) {  // debug: where the issuing thread spent its clocks
                P.timing[blockIdx.x * 8 + 0] = clock64() - t_all0;
                P.timing[blockIdx.x * 8 + 1] = t_tmem;
                P.timing[blockIdx.x * 8 + 2] = t_full;
                P.timing[blockIdx.x * 8 + 3] = item_end - item_begin;
            }
        }
    } else {
        // ===== epilogue: 8 warps; thread = (query slot = TMEM lane, half of the tile's 128 columns) =====
        const int quarter = warp & 3;              // TMEM lanes this warp may touch: 32*quarter ..
        const int half = (warp - 2) >> 2;          // columns 64*half ..
        int slot = quarter * 32 + lane;  // query slot of this thread's TMEM lane (M = 64 units: see new_unit below)
        const int et = (warp - 2) * 32 + lane;     // 0..255
        const uint2 *my_stash = &s_stash[0][et];
        const CandOut C{P.cnt, P.cand_negv, P.cand_rel, P.cand_pos, P.cap};
        const int kq = P.k;
        uint32_t tphase = 0;
        int u = item_begin < item_end ? find_unit(P.unit_item0, P.nunits, item_begin) : 0;
        int qi = -1, rel0 = 0, nst = 0, list_len = 0, resv = 0, resv_base = 0;
        int64_t list_row0 = 0;
        float thr = INFINITY, margin = INFINITY, pend_m1 = -INFINITY, pend_m2 = -INFINITY;
        bool have_pend = false;
        // Running lower bounds on the query's best scores.  Bucket b of this thread holds the two best scores among the
        // rows with (column % NB == b) of its column half seen in this unit:
        //   min_b(b2) has >= 2*NB rows at or above it, the minimum over the two threads of a slot 4*NB = kk (all emitted:
        //   the kk re-scored candidates can be taken from them);
        //   min_b(b1) has >= NB rows at or above it (2*NB over both threads): for k <= NB (2*NB) the k-th best score is at
        //   least that, and a row more than `margin` (> 2 eps_q) below it can never enter the exact top-k.
        float b1[NB], b2[NB];
        int it = 0;
        for (int item = item_begin; item < item_end; ++item, ++it) {
            while (item >= P.unit_item0[u + 1]) ++u;
            const int t = (P.unit_tile0 ? P.unit_tile0[u] : P.tile_start) + (item - P.unit_item0[u]) * P.tile_stride;
            const bool new_unit = item == item_begin || item == P.unit_item0[u];
            const int l = P.unit_list[u];
            const int64_t btile = P.tile_off[l] + t;
            const int par = it & 1;
            if (et < kFastTile) s_rs[par][et] = P.rs[btile * kFastTile + et] * 256.0f;
            else s_ro[par][et - kFastTile] = P.ro[btile * kFastTile + et - kFastTile];
            float thr_g = -INFINITY;
            if (MODE == FAST_EMIT && !new_unit && qi >= 0) thr_g = P.thr[qi];  // raised meanwhile by other CTAs
            asm volatile("bar.sync 1, 256;" ::: "memory");
            bool flood = false;
            if (MODE == FAST_EMIT) {
                if (have_pend && qi >= 0) {
                    // every row counted in the buckets was emitted under a threshold <= the new one, so raising thr keeps
                    // "at least 64 candidates at or above thr" true
                    float c = fminf(pend_m2, s_m[par ^ 1][0][half ^ 1][slot]);
                    if (kq <= 2 * NB) {
                        const float c1 = fminf(pend_m1, s_m[par ^ 1][1][half ^ 1][slot]);
                        c = fmaxf(c, c1 - margin - 4e-6f * fabsf(c1));
                    }
                    if (c > thr) {
                        thr = c;
                        atomic_max_float(&P.thr[qi], thr);
                    }
                }
                if (new_unit) {
                    if (nst > 0) flush_stash(C, qi, rel0, list_row0, nst, -1, my_stash);
                    nst = 0;
                    // M = 64 unit: accumulator row r lives in lane 32 * (r / 16) + r % 16; lanes 16-31 of a quadrant hold
                    // nothing and are pointed at the unit's unused slots 64.. (slot_query = -1 there)
                    const bool hm = P.unit_nsel != nullptr && P.unit_nsel[u] <= 64;
                    slot = hm ? ((lane < 16 ? 0 : 64) + quarter * 16 + (lane & 15)) : quarter * 32 + lane;
                    qi = P.slot_query[(int64_t)u * kFastTile + slot];
                    rel0 = P.slot_rel0[(int64_t)u * kFastTile + slot];
                    list_row0 = P.list_off[l];
                    list_len = (int)(P.list_off[l + 1] - list_row0);
                    thr = qi >= 0 ? P.thr[qi] : INFINITY;
                    margin = qi >= 0 ? P.margin[qi] : INFINITY;
                    have_pend = false;
#pragma unroll
                    for (int b = 0; b < NB; ++b) b1[b] = b2[b] = -INFINITY;
                } else {
                    thr = fmaxf(thr, thr_g);
                }
                // no bound yet: every real row of this tile will be a candidate -- reserve their slots now, so that the
                // round trip of the atomic overlaps the wait for the accumulators
                flood = qi >= 0 && thr == -INFINITY && !(rel0 == 0 && t < P.skip_tiles);
                if (flood) {
                    if (nst > 0) flush_stash(C, qi, rel0, list_row0, nst, -1, my_stash);
                    nst = 0;
                    resv = min(max(list_len - t * kFastTile - half * 64, 0), 64);
                    resv_base = resv > 0 ? atomicAdd(&P.cnt[qi], resv) : 0;
                }
            }
            const long long e0 = (P.timing && et == 0) ? clock64() : 0;
            mbar_wait(smem_u32(&s_tmem_full), tphase);
            const long long e1 = (P.timing && et == 0) ? clock64() : 0;
            tc_fence_after();
            tphase ^= 1u;
            const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);
            // phase A: accumulators -> one fp32 score per column, in registers; then the accumulators are free again
            float v[64];
#pragma unroll
            for (int c16 = 0; c16 < 4; ++c16) {
                uint32_t c0[16], c1[16], c2[16];
                tmem_ld16(tlane + c16 * 16, c0);
                tmem_ld16(tlane + kFastTile + c16 * 16, c1);
                tmem_ld16(tlane + 2 * kFastTile + c16 * 16, c2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // S / 256 = c0 * 256 + c1 + c2 / 256; the dropped fraction is bounded in eps_q (query_bounds_kernel)
                    const int lo = (int)c1[j] + ((int)c2[j] >> 8);
                    const float s = fmaf((float)(int)c0[j], 256.0f, (float)lo);
                    const int col = half * 64 + c16 * 16 + j;
                    v[c16 * 16 + j] = fmaf(s, s_rs[par][col], s_ro[par][col]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_tmem_empty));
            if (P.timing && et == 0) {
                const long long e2 = clock64();
                P.timing[blockIdx.x * 8 + 4] += e1 - e0;  // waiting for the accumulators
                P.timing[blockIdx.x * 8 + 5] += e2 - e1;  // phase A
            }
            // phase B: thresholds and candidates
            if (MODE == FAST_EMIT) {
                // padding rows score -inf: never candidates; a nearest-list slot emits nothing for the tiles the sample pass took
                const float cut = (rel0 == 0 && t < P.skip_tiles) ? INFINITY : fmaxf(thr, -FLT_MAX);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint32_t mask = 0;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = g * 8 + jj, b = j % NB;
                        const float lo = fminf(b1[b], v[j]);
                        b1[b] = fmaxf(b1[b], v[j]);
                        b2[b] = fmaxf(b2[b], lo);
                        mask |= (v[j] >= cut ? 1u : 0u) << jj;
                    }
                    // the whole warp empties its stashes together: one round trip of the atomics instead of one per lane
                    if (__any_sync(0xffffffffu, nst + __popc(mask) > STASH)) {
                        if (nst > 0) {
                            flush_stash(C, qi, rel0, list_row0, nst, flood ? resv_base : -1, my_stash);
                            if (flood) {
                                resv_base += nst;
                                resv -= nst;
                            }
                            nst = 0;
                        }
                    }
                    if (mask) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            if ((mask >> jj) & 1u) {
                                s_stash[nst][et] =
                                    make_uint2(__float_as_uint(v[g * 8 + jj]), (uint32_t)(t * kFastTile + half * 64 + g * 8 + jj));
                                ++nst;
                            }
                        }
                    }
                }
                if (flood) {
                    if (nst > 0) flush_stash(C, qi, rel0, list_row0, nst, resv_base, my_stash);
                    // the reservation assumed that exactly the real rows pass; anything else sends the query to the exact path
                    if (resv != nst) atomicAdd(&P.cnt[qi], P.cap + 1);
                    nst = 0;
                    resv = 0;
                }
                float m1 = b1[0], m2 = b2[0];
#pragma unroll
                for (int b = 1; b < NB; ++b) {
                    m1 = fminf(m1, b1[b]);
                    m2 = fminf(m2, b2[b]);
                }
                s_m[par][0][half][slot] = m2;
                s_m[par][1][half][slot] = m1;
                pend_m1 = m1;
                pend_m2 = m2;
                have_pend = true;
                if (kq <= NB && qi >= 0) {
                    const float c = m1 - margin - 4e-6f * fabsf(m1);
                    if (c > thr) {
                        thr = c;
                        atomic_max_float(&P.thr[qi], thr);
                    }
                }
            } else {
                float *o = P.dump + ((int64_t)item * kFastTile + slot) * kFastTile + half * 64;
#pragma unroll
                for (int j = 0; j < 64; ++j) o[j] = v[j];
            }
        }
        if (MODE == FAST_EMIT) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (have_pend && qi >= 0) {
                const int pp = (it - 1) & 1;
                float c = fminf(pend_m2, s_m[pp][0][half ^ 1][slot]);
                if (kq <= 2 * NB) {
                    const float c1 = fminf(pend_m1, s_m[pp][1][half ^ 1][slot]);
                    c = fmaxf(c, c1 - margin - 4e-6f * fabsf(c1));
                }
                if (c > thr) atomic_max_float(&P.thr[qi], c);
            }
            if (nst > 0) flush_stash(C, qi, rel0, list_row0, nst, -1, my_stash);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Narrow units: rows on the M side.
//
// An IVF scan after probe pruning (and the threshold sample of every scan) is made of units that hold a handful of
// queries: a 128 x 128 item then costs its 96 MMAs at the N = 128 floor (64 clocks each, ~12 k clocks with the hand-off)
// while HBM needs 8.4 k clocks to deliver the 196 KB of row digits — the tensor pipe, 90 % of it padding, is the bound.
// tc_narrow_kernel swaps the operands: A = the row tile (M = 128 rows), B = the unit's first 32 query slots (N = 32), so
// an MMA takes 16 clocks (1.5 k per item), the accumulators are 3 x 32 TMEM columns (two sets: the epilogue of item i
// overlaps the MMAs of item i + 1) and the only cost left is the bulk copy of the row digits: the pass is HBM-bound.
// Epilogue: thread = row (TMEM lane), columns = queries.  Thresholds are read, not raised, here (they come from the
// sample pass; a query without one takes every real row of the tile, slots reserved with one atomic per (query, tile));
// a row at or above a query's threshold is appended with one warp-aggregated atomic.  Correctness does not depend on the
// threshold's value: the proof in fast_final_kernel bounds whatever was not emitted by it.
// =====================================================================================================================
constexpr int NARROW_EPI_THREADS = 256;
constexpr int NARROW_THREADS = 64 + NARROW_EPI_THREADS;
template <int NS>
struct NarrowCfg {
    static constexpr int A_BYTES = NS * kFastImg;                  // one k-block of the row tile, every digit
    static constexpr int B_SLICE = kNarrowSlots * kFastKB;         // one k-block of 32 query slots, one digit (4 KB)
    static constexpr int B_BYTES = NS * B_SLICE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NSTAGE = NS == 2 ? 5 : 3;
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024;
    static constexpr int ACC_COLS = kNarrowSlots;       // TMEM columns of one digit-product accumulator: [hi.hi | hi.lo | lo.hi | lo.lo] (row digit . query digit)
    static constexpr int SET_COLS = 4 * kNarrowSlots;   // columns of one accumulator set
    static constexpr int TMEM_COLS = 8 * kNarrowSlots;  // two sets (a power of two >= 32)
    static constexpr int GPT = kNarrowSlots / 16;       // groups of 8 query columns per epilogue thread (groups chalf, chalf + 2, ...)
    static_assert(kNarrowSlots == 16 || kNarrowSlots == 32, "narrow units hold 16 or 32 query slots");
};
// D = S32, A = B = signed int8, K-major, M = 128, N = 64: the B operand is the unit's 32 query slots TWICE — their high
// digits (rows 0-31 of the tile) stacked on their low digits (rows 32-63) — so one MMA per row digit yields both digit
// products and every row image is read from shared memory once per k-step instead of twice (NS = 2 only).
constexpr uint32_t kIdescI8Narrow = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * kNarrowSlots) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

template <int NS>
__global__ void __launch_bounds__(NARROW_THREADS, 1) tc_narrow_kernel(const TcParams P) {
    using Cfg = NarrowCfg<NS>;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t s_full[Cfg::NSTAGE];
    __shared__ __align__(8) uint64_t s_empty[Cfg::NSTAGE];
    __shared__ __align__(8) uint64_t s_tmem_full[2];
    __shared__ __align__(8) uint64_t s_tmem_empty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ int32_t s_q[2][kNarrowSlots];     // per item parity: query of each slot (-1: unused)
    __shared__ int32_t s_rel0[2][kNarrowSlots];  // position of the (query, list) segment in the query's concatenation
    __shared__ float s_thr[2][kNarrowSlots];
    __shared__ int32_t s_base[2][kNarrowSlots];  // >= 0: every real row of the tile is a candidate, its slots start here

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage0 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const int kbn = P.kbn;
    const int32_t *item0 = P.unit_item0n;
    const int total_items = item0[P.nunits];
    if (total_items == 0) return;
    const int item_begin = (int)((int64_t)total_items * blockIdx.x / gridDim.x);
    const int item_end = (int)((int64_t)total_items * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::NSTAGE; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&s_tmem_full[s]), 1);
            mbar_init(smem_u32(&s_tmem_empty[s]), NARROW_EPI_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "n"(Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ===== producer: per k-block the row tile's digit images (NS x 16 KB, from HBM) + the unit's 32 query slots (NS x 4 KB, L2).
        // One lane, large copies: splitting a stage into 96 pieces of <= 1 KB issued by the whole warp was measured 2.3x slower
        // (the copy engine takes ~60 clocks per operation). =====
        if (lane == 0 && item_begin < item_end) {
            int stage = 0;
            uint32_t phase = 0;
            long long t_wait = 0;
            int u = find_unit(item0, P.nunits, item_begin);
            for (int item = item_begin; item < item_end; ++item) {
                while (item >= item0[u + 1]) ++u;
                const int t = (P.unit_tile0 ? P.unit_tile0[u] : P.tile_start) + (item - item0[u]) * P.tile_stride;
                const int64_t btile = P.tile_off[P.unit_list[u]] + t;
                const int8_t *qsrc = P.aimg + (int64_t)u * kbn * NS * kFastImg;
                const int8_t *rsrc = P.bimg + btile * kbn * NS * kFastImg;
                // only the 8-slot groups the unit uses are copied (and were packed): the MMA multiplies whatever the rest of the
                // B tile holds, and the epilogue never looks at those columns
                const uint32_t bsz = (uint32_t)((P.unit_nsel_all[u] + 7) >> 3) * 8u * kFastKB;
                for (int kb = 0; kb < kbn; ++kb) {
                    const long long c0 = P.timing ? clock64() : 0;
                    mbar_wait(smem_u32(&s_empty[stage]), phase ^ 1u);
                    if (P.timing) t_wait += clock64() - c0;
                    const uint32_t full = smem_u32(&s_full[stage]);
                    const uint32_t dst = stage0 + (uint32_t)stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(full, Cfg::A_BYTES + NS * bsz);
                    bulk_g2s(dst, rsrc + (int64_t)kb * NS * kFastImg, Cfg::A_BYTES, full);
#pragma unroll
                    for (int sl = 0; sl < NS; ++sl)
                        bulk_g2s(dst + Cfg::A_BYTES + sl * Cfg::B_SLICE, qsrc + ((int64_t)kb * NS + sl) * kFastImg, bsz, full);
                    if (++stage == Cfg::NSTAGE) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
            if (P.timing) P.timing[(gridDim.x + blockIdx.x) * 8 + 7] = t_wait;
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[rows x queries] (+)= A[rows] * B[queries]^T, accumulator sets alternate per item =====
        if (lane == 0 && item_begin < item_end) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t tphase[2] = {0, 0};
            int it = 0;
            long long t_tmem = 0, t_full = 0;
            const long long t_all0 = clock64();
            for (int item = item_begin; item < item_end; ++item, ++it) {
                const int set = it & 1;
                long long c0 = P.timing ? clock64() : 0;
                mbar_wait(smem_u32(&s_tmem_empty[set]), tphase[set] ^ 1u);
                if (P.timing) t_tmem += clock64() - c0;
                tphase[set] ^= 1u;
                tc_fence_after();
                const uint32_t dbase = tmem_base + (uint32_t)set * Cfg::SET_COLS;
                for (int kb = 0; kb < kbn; ++kb) {
                    c0 = P.timing ? clock64() : 0;
                    mbar_wait(smem_u32(&s_full[stage]), phase);
                    if (P.timing) t_full += clock64() - c0;
                    tc_fence_after();
                    const uint32_t abase = stage0 + (uint32_t)stage * Cfg::STAGE_BYTES;
                    const uint32_t bbase = abase + Cfg::A_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < kFastKB / 32; ++k4) {
                        const uint64_t bd = smem_desc(bbase + k4 * 32);  // 64 rows: the slots' high digits, then their low digits
#pragma unroll
                        for (int sr = 0; sr < 2; ++sr) {  // digit of the row
                            const uint64_t ad = smem_desc(abase + sr * kFastImg + k4 * 32);
                            umma_i8(dbase + (uint32_t)sr * 2 * Cfg::ACC_COLS, ad, bd, kIdescI8Narrow, (kb == 0 && k4 == 0) ? 0u : 1u);
                        }
                    }
                    umma_commit(smem_u32(&s_empty[stage]));
                    if (++stage == Cfg::NSTAGE) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(smem_u32(&s_tmem_full[set]));
            }
            if (P.timing) {  // debug (narrow entries live behind the dense kernel's)
                long long *T = P.timing + (gridDim.x + blockIdx.x) * 8;
                T[0] = clock64() - t_all0;
                T[1] = t_tmem;
                T[2] = t_full;
                T[3] = item_end - item_begin;
            }
        }
    } else {
        // ===== epilogue: 8 warps; thread = (row of the tile = TMEM lane, every other group of 8 query columns) =====
        // A lone warp per scheduler issues a dependent instruction every ~5 clocks, so the work per item is kept short: only
        // the column groups the unit really uses are read and scored, the unit's slot data is fetched once per unit (not per
        // item), one compare per column decides (thr_eff: +inf unused, -inf take every real row).
        const int quarter = warp & 3;
        const int chalf = (warp - 2) >> 2;      // column groups chalf, chalf + 2
        const int et = (warp - 2) * 32 + lane;  // 0..255; threads 0..31 stage the unit's slots
        const int row = quarter * 32 + lane;    // row of the tile = accumulator lane
        uint32_t tphase[2] = {0, 0};
        int u = item_begin < item_end ? find_unit(item0, P.nunits, item_begin) : 0;
        int it = 0;
        long long t_meta = 0, t_wait = 0, t_proc = 0;
        // per-unit state, refreshed when the unit changes
        int cur_u = -1, u_end = 0, u_item0 = 0, u_tile0 = 0, nsel = 0, list_len = 0, my_q = -1, my_rel0 = 0;
        int64_t tile0 = 0, list_row0 = 0;
        float my_thr = INFINITY;
        for (int item = item_begin; item < item_end; ++item, ++it) {
            const long long e0 = (P.timing && et == 0) ? clock64() : 0;
            if (cur_u < 0 || item >= u_end) {
                while (item >= item0[u + 1]) ++u;
                cur_u = u;
                u_item0 = item0[u];
                u_end = item0[u + 1];
                u_tile0 = P.unit_tile0 ? P.unit_tile0[u] : P.tile_start;
                const int l = P.unit_list[u];
                nsel = P.unit_nsel_all[u];
                tile0 = P.tile_off[l];
                list_row0 = P.list_off[l];
                list_len = (int)(P.list_off[l + 1] - list_row0);
                if (et < kNarrowSlots) {
                    my_q = P.slot_query[(int64_t)u * kFastTile + et];
                    my_rel0 = my_q >= 0 ? P.slot_rel0[(int64_t)u * kFastTile + et] : 0;
                    my_thr = my_q >= 0 ? P.thr[my_q] : INFINITY;
                }
            }
            const int t = u_tile0 + (item - u_item0) * P.tile_stride;
            const int64_t btile = tile0 + t;
            const int par = it & 1, set = it & 1;
            const int nvalid = min(max(list_len - t * kFastTile, 0), kFastTile);
            if (et < kNarrowSlots) {
                int base = -1;
                // a nearest-list slot emits nothing for the tiles the sample pass already took (its candidates were kept)
                const bool sampled = my_rel0 == 0 && t < P.skip_tiles;
                // no bound yet: every real row of the tile is a candidate of this query, its slots reserved with one atomic
                if (my_q >= 0 && my_thr == -INFINITY && nvalid > 0 && !sampled) base = atomicAdd(&P.cnt[my_q], nvalid);
                s_q[par][et] = my_q;
                s_rel0[par][et] = my_rel0;
                s_thr[par][et] = sampled ? INFINITY : my_thr;
                s_base[par][et] = base;
            }
            const float rs = P.rs[btile * kFastTile + row] * 256.0f;
            const bool valid = row < nvalid;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const long long e1 = (P.timing && et == 0) ? clock64() : 0;
            mbar_wait(smem_u32(&s_tmem_full[set]), tphase[set]);
            const long long e2 = (P.timing && et == 0) ? clock64() : 0;
            tphase[set] ^= 1u;
            tc_fence_after();
            const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)set * Cfg::SET_COLS;
            // this thread's columns: group g0 = chalf (columns 8 g0 ..) and g1 = chalf + 2, where the unit reaches them
            constexpr int GPT = Cfg::GPT, NC = 8 * GPT;  // this thread's columns
            const bool use0 = chalf * 8 < nsel, use1 = GPT > 1 && (chalf + 2) * 8 < nsel;
            uint32_t c0[NC], c1[NC], c1b[NC], c2[NC];  // hi.hi, hi.lo, lo.hi, lo.lo
            if (use0) {
                tmem_ld8(tl + chalf * 8, &c0[0]);
                tmem_ld8(tl + Cfg::ACC_COLS + chalf * 8, &c1[0]);
                tmem_ld8(tl + 2 * Cfg::ACC_COLS + chalf * 8, &c1b[0]);
                tmem_ld8(tl + 3 * Cfg::ACC_COLS + chalf * 8, &c2[0]);
            }
            if (GPT > 1 && use1) {
                tmem_ld8(tl + (chalf + 2) * 8, &c0[NC - 8]);
                tmem_ld8(tl + Cfg::ACC_COLS + (chalf + 2) * 8, &c1[NC - 8]);
                tmem_ld8(tl + 2 * Cfg::ACC_COLS + (chalf + 2) * 8, &c1b[NC - 8]);
                tmem_ld8(tl + 3 * Cfg::ACC_COLS + (chalf + 2) * 8, &c2[NC - 8]);
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_tmem_empty[set]));
            if (use0) {
                const int pos_in_list = t * kFastTile + row;
                // column c (0..NC-1) of this thread is slot col_of(c) of the unit
                float sc[NC];
                uint32_t passbits = 0;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (c >= 8 && !use1) break;
                    const int slot = (chalf + (c >> 3) * 2) * 8 + (c & 7);
                    const int lo = (int)c1[c] + (int)c1b[c] + ((int)c2[c] >> 8);
                    sc[c] = fmaf((float)(int)c0[c], 256.0f, (float)lo) * rs;
                    // +inf: unused slot, -inf: every real row is taken (s_base >= 0)
                    const float th = s_base[par][slot] >= 0 ? -INFINITY : (s_q[par][slot] >= 0 ? s_thr[par][slot] : INFINITY);
                    passbits |= ((valid && sc[c] >= th) ? 1u : 0u) << c;
                }
                // lane c of the warp reserves the slots of column c with ONE atomic for the warp's 32 rows: all columns' atomics
                // are in flight together
                uint32_t mymask = 0;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (c >= 8 && !use1) break;
                    const uint32_t m = __ballot_sync(0xffffffffu, (passbits >> c) & 1u);
                    if (lane == c) mymask = m;
                }
                const int myslot = (chalf + ((lane & 15) >> 3) * 2) * 8 + (lane & 7);
                int mybase = 0;
                if (lane < NC && mymask != 0 && s_base[par][myslot] < 0) mybase = atomicAdd(&P.cnt[s_q[par][myslot]], __popc(mymask));
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (c >= 8 && !use1) break;
                    const uint32_t m = __shfl_sync(0xffffffffu, mymask, c);
                    if (m == 0) continue;  // uniform over the warp
                    const int b = __shfl_sync(0xffffffffu, mybase, c);
                    if ((passbits >> c) & 1u) {
                        const int slot = (chalf + (c >> 3) * 2) * 8 + (c & 7);
                        const int fb = s_base[par][slot];
                        const int idx = fb >= 0 ? fb + row : b + __popc(m & ((1u << lane) - 1u));
                        if (idx < P.cap) {
                            const int64_t o = (int64_t)s_q[par][slot] * P.cap + idx;
                            P.cand_negv[o] = -(double)sc[c];
                            P.cand_rel[o] = s_rel0[par][slot] + pos_in_list;
                            P.cand_pos[o] = (int32_t)(list_row0 + pos_in_list);
                        }
                    }
                }
            }
            if (P.timing && et == 0) {
                t_meta += e1 - e0;
                t_wait += e2 - e1;
                t_proc += clock64() - e2;
            }
        }
        if (P.timing && et == 0) {
            long long *T = P.timing + (gridDim.x + blockIdx.x) * 8;
            T[4] = t_wait;
            T[5] = t_meta;
            T[6] = t_proc;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

template <int NS>
void narrow_launch(const TcParams &P) {
    auto kernel = tc_narrow_kernel<NS>;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NarrowCfg<NS>::SMEM_BYTES));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(g_num_sms, P.items_hint > 0 ? P.items_hint : (int64_t)P.nunits * 8));
    kernel<<<grid, NARROW_THREADS, NarrowCfg<NS>::SMEM_BYTES, g_stream>>>(P);
    HB_LAUNCH_CHECK();
}

template <int NS, int MODE, int NB>
void tc_launch(const TcParams &P, int total_items) {
    auto kernel = tc_pass_kernel<NS, MODE, NB>;
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
    if (P.unit_item0n != nullptr) {  // the units with <= kNarrowSlots selections (unit_item0 then covers the others only)
        HB_REQUIRE(mode == FAST_EMIT && ns == 2, "narrow units serve 2-digit EMIT passes only");
        narrow_launch<2>(P);
    }
    HB_REQUIRE(P.kk == 64 || P.kk == 128 || mode == FAST_DUMP, "candidates re-scored per query must be 64 or 128");
    // a flat scan of a few units still has thousands of items: size the grid by the items when the host knows them
    const int hint = (int)std::min<int64_t>(1 << 30, P.items_hint > 0 ? P.items_hint : (int64_t)P.nunits * 8);
#define HB_TC(NS_)                                                          \
    do {                                                                    \
        if (mode == FAST_DUMP) tc_launch<NS_, FAST_DUMP, 16>(P, hint);      \
        else if (P.kk == 64) tc_launch<NS_, FAST_EMIT, 16>(P, hint);        \
        else tc_launch<NS_, FAST_EMIT, 32>(P, hint);                        \
    } while (0)
    if (ns == 2) HB_TC(2);
    else HB_TC(3);
#undef HB_TC
}

}  // namespace hb
