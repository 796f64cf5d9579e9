// hb_api.cu — the C ABI of libhnswb200.so (include/hnswb200.h): handles, staging of host/device buffers and
// the orchestration of the kernels behind hnsw-clj's build-index / search-knn / search-batch* surface.
#include <errno.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <cuda_profiler_api.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "hb_build.cuh"
#include "hb_comm.cuh"
#include "hb_fast.cuh"
#include "hb_hnsw.cuh"
#include "hb_kernels.cuh"
#include "hb_lanes.cuh"
#include "hb_validate.cuh"

namespace hb {

int64_t g_launches = 0;
cudaStream_t g_stream = 0;
int g_num_sms = 148;
static int g_mode = HB_MODE_EXACT;
static bool g_inited = false;
static int g_device = -1;
static std::mutex g_mu;
static thread_local std::string t_err;

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
const void *stage_in(const void *p, size_t bytes, DevBuf &stage) {
    if (bytes == 0) return p;
    if (is_device_ptr(p)) return p;
    void *d = stage.get(bytes);
    HB_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, g_stream));
    return d;
}
OutStage stage_out(void *user, size_t bytes, DevBuf &stage) {
    OutStage o;
    o.user = user;
    o.bytes = bytes;
    if (!user || bytes == 0) return o;
    o.host = !is_device_ptr(user);
    o.dev = o.host ? stage.get(bytes) : user;
    return o;
}
void finish_out(const OutStage &o) {
    if (o.user && o.host && o.bytes) HB_CUDA(cudaMemcpyAsync(o.user, o.dev, o.bytes, cudaMemcpyDeviceToHost, g_stream));
}

static size_t g_scratch_budget = 0;
static size_t scratch_budget() {
    if (!g_scratch_budget) {
        const char *e = getenv("HB_SCRATCH_MB");
        g_scratch_budget = (e && atoll(e) > 0) ? (size_t)atoll(e) << 20 : (size_t)8 << 30;
    }
    return g_scratch_budget;
}

static void ensure_init(int device = -1) {
    if (g_inited && (device < 0 || device == g_device)) return;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw Error(HB_ERR_NO_DEVICE, "no CUDA device is visible (libhnswb200 has no CPU fallback)");
    }
    int dev = device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    }
    if (dev >= count) throw Error(HB_ERR_NO_DEVICE, "device index out of range");
    cudaDeviceProp prop;
    HB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        throw Error(HB_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100 (kernels are built for sm_100a only)");
    HB_CUDA(cudaSetDevice(dev));
    g_device = dev;
    g_num_sms = prop.multiProcessorCount;
    g_inited = true;
}

// all transient device buffers; grown on demand, released by hb_shutdown
struct Workspace {
    DevBuf in_a, in_b, in_c, in_d, out_a, out_b, out_c;
    DevBuf qnorm, scratch, sel_val, sel_pos, cand_val, cand_id, plan, probes, pair_out, qsel, lq_off, tile_prefix, tmp,
        misc, misc2, misc3, misc4, vflag, sh_ids, sh_dist, sh_sums, sh_counts, sh_small;
    void release() {
        DevBuf *all[] = {&in_a, &in_b, &in_c, &in_d, &out_a, &out_b, &out_c, &qnorm, &scratch, &sel_val, &sel_pos, &cand_val,
                         &cand_id, &plan, &probes, &pair_out, &qsel, &lq_off, &tile_prefix, &tmp, &misc, &misc2, &misc3, &misc4, &vflag,
                         &sh_ids, &sh_dist, &sh_sums, &sh_counts, &sh_small};
        for (DevBuf *b : all) b->release();
    }
};
static Workspace g_ws;

template <typename F>
static int guarded(F &&f) {
    try {
        std::lock_guard<std::mutex> lk(g_mu);
        // the CUDA current device is per host thread: a caller thread that did not run hb_init (JVM pool threads, the
        // micro-batcher's leader of the moment) would otherwise allocate and launch on device 0
        if (g_inited) HB_CUDA(cudaSetDevice(g_device));
        f();
        return HB_OK;
    } catch (const Error &e) {
        t_err = e.what();
        return e.status;
    } catch (const std::exception &e) {
        t_err = e.what();
        return HB_ERR_CUDA;
    }
}

static void sync_stream() { HB_CUDA(cudaStreamSynchronize(g_stream)); }

// Range checks of caller-supplied index arrays (hb_validate.cu): the checks OR into one device flag, require() reads it.
struct Validator {
    int32_t *flag;
    Validator() : flag(g_ws.vflag.as<int32_t>(1)) { HB_CUDA(cudaMemsetAsync(flag, 0, 4, g_stream)); }
    void range_i32(const int32_t *p, int64_t n, int64_t lo, int64_t hi) { launch_check_range_i32(p, n, lo, hi, flag); }
    void range_i64(const int64_t *p, int64_t n, int64_t lo, int64_t hi) { launch_check_range_i64(p, n, lo, hi, flag); }
    void offsets(const int64_t *off, int64_t count, int64_t total) { launch_check_offsets(off, count, total, flag); }
    void csr(const int64_t *off, const int32_t *ids, int64_t n) { launch_check_csr(off, ids, n, 0, flag); }
    bool ok() {
        int32_t h = 0;
        HB_CUDA(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaStreamSynchronize(g_stream));
        HB_CUDA(cudaMemsetAsync(flag, 0, 4, g_stream));
        return h == 0;
    }
    void require(const char *msg) {
        if (!ok()) throw Error(HB_ERR_INVALID, msg);
    }
};

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events on the launching stream ------
enum ProfTag { PROF_SCAN = 0, PROF_COARSE = 1, PROF_SELECT = 2, PROF_PLAN = 3, PROF_ASSIGN = 4, PROF_TC = 5, PROF_PACK = 6, PROF_RESCORE = 7, PROF_TC_SAMPLE = 8, PROF_HNSW = 9,
               PROF_EXCHANGE = 10, PROF_MERGE = 11, PROF_ALLREDUCE = 12, PROF_UPDATE = 13, PROF_NTAGS = 14 };
static bool g_profile = false;
struct ProfSpan {
    cudaEvent_t a, b;
    int tag;
};
static std::vector<ProfSpan> g_spans;
static double g_prof_ms[PROF_NTAGS] = {0};
static int64_t g_prof_n[PROF_NTAGS] = {0};
struct Prof {
    bool on;
    ProfSpan sp;
    explicit Prof(int tag) : on(g_profile && tag >= 0) {
        if (!on) return;
        sp.tag = tag;
        cudaEventCreate(&sp.a);
        cudaEventCreate(&sp.b);
        cudaEventRecord(sp.a, g_stream);
    }
    ~Prof() {
        if (!on) return;
        cudaEventRecord(sp.b, g_stream);
        g_spans.push_back(sp);
    }
};
static void prof_collect() {
    for (auto &sp : g_spans) {
        cudaEventSynchronize(sp.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, sp.a, sp.b);
        g_prof_ms[sp.tag] += ms;
        g_prof_n[sp.tag] += 1;
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    g_spans.clear();
}
double fp64_peak_tflops();  // hb_microbench.cu
double mma_clocks_per_instr(int kind);

// ---- java.util.Random (needed by kmeans-plus-plus-init, ivf_flat.clj:37: (Random. 42)) -------------
// Implemented from the published LCG definition of java.util.Random (JDK javadoc).
struct JavaRandom {
    uint64_t s;
    explicit JavaRandom(int64_t seed) : s(((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) {}
    int32_t next(int bits) {
        s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)(uint32_t)(s >> (48 - bits));
    }
    int32_t next_int(int32_t bound) {
        int32_t r = next(31);
        const int32_t m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        for (int32_t u = r;; u = next(31)) {
            r = u % bound;
            if ((int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m) >= 0) break;
        }
        return r;
    }
    double next_double() {
        const int64_t hi = next(26), lo = next(27);
        return (double)((hi << 27) + lo) * 0x1.0p-53;
    }
};

// StrictMath.log, i.e. fdlibm's __ieee754_log (e_log.c): java.util.Random.nextGaussian goes through it, and the LSH
// projection matrices (src/hnsw/ann/hash/hybrid_lsh.clj:24-31) are nextGaussian draws.  libm's log is not bit-identical.
static double strict_log(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                        Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    uint64_t u;
    memcpy(&u, &x, 8);
    int32_t hx = (int32_t)(u >> 32), k = 0;
    const uint32_t lx = (uint32_t)u;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -INFINITY;
        if (hx < 0) return NAN;
        k -= 54;
        x *= two54;
        memcpy(&u, &x, 8);
        hx = (int32_t)(u >> 32);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    memcpy(&u, &x, 8);
    u = (u & 0xffffffffull) | ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32);
    memcpy(&x, &u, 8);
    k += (i >> 20);
    const double f = x - 1.0, dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) return k == 0 ? 0.0 : dk * ln2_hi + dk * ln2_lo;
        const double R = f * f * (0.5 - 0.33333333333333333 * f);
        return k == 0 ? f - R : dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    const double s = f / (2.0 + f), z = s * s, w = z * z;
    i = hx - 0x6147a;
    const int32_t j = 0x6b851 - hx;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6)), t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    const double R = t2 + t1;
    if (i > 0) {
        const double hfsq = 0.5 * f * f;
        return k == 0 ? f - (hfsq - s * (hfsq + R)) : dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    return k == 0 ? f - s * (f - R) : dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}
// java.util.Random.nextGaussian: Marsaglia's polar method, one spare value kept
struct JavaGaussian {
    JavaRandom rng;
    bool have = false;
    double spare = 0.0;
    explicit JavaGaussian(int64_t seed) : rng(seed) {}
    double next() {
        if (have) {
            have = false;
            return spare;
        }
        double v1, v2, s;
        do {
            v1 = 2.0 * rng.next_double() - 1.0;
            v2 = 2.0 * rng.next_double() - 1.0;
            s = v1 * v1 + v2 * v2;
        } while (s >= 1.0 || s == 0.0);
        const double mul = std::sqrt(-2.0 * strict_log(s) / s);
        spare = v2 * mul;
        have = true;
        return v1 * mul;
    }
};

static void metric_to_epi(int metric, bool guarded_cos, bool &l2, int &epi) {
    l2 = metric == HB_L2;
    if (metric == HB_COSINE) epi = guarded_cos ? EPI_COS_GUARD : EPI_COS;
    else if (metric == HB_L2) epi = EPI_L2;
    else if (metric == HB_IP) epi = EPI_NEGDOT;
    else throw Error(HB_ERR_INVALID, "unknown metric");
}
static void check_dtype(int dtype) {
    HB_REQUIRE(dtype == HB_F32 || dtype == HB_BF16 || dtype == HB_F64, "unknown dtype");
}

// digit images of one set of vectors (hb_fast.cuh)
struct FastSideBufs {
    DevBuf img, rs, ro, tile_off, stats, list_off;
    int ns = 0, kbn = 0;
    int64_t ntiles = 0, nrows = 0;
    bool built = false, usable = false;
    size_t bytes() const { return img.cap + rs.cap + ro.cap + tile_off.cap + stats.cap + list_off.cap; }
    void release() {
        img.release();
        rs.release();
        ro.release();
        tile_off.release();
        stats.release();
        list_off.release();
        built = false;
    }
};

static bool g_fast_debug = false;
static bool g_tc_interleave = true;
static int g_fast_ns = 2;            // digits per element in FAST mode (2: 16-bit mantissas, 3: 24-bit)
static int g_hnsw_prefetch = -1;     // hb_set_option("hnsw_prefetch", lines): -1 = sized to the L2 (hnsw_search)
static bool g_tc_half_m = true;      // EMIT passes: units with <= 64 selections run with M = 64 (hb_tc.cu)
static bool g_fast_carry = true;     // IVF list scans: the sample pass's candidates are kept, its tiles are not scanned twice
static bool g_tc_narrow = true;      // IVF list scans: units with <= 32 selections run with the rows on the M side (tc_narrow_kernel)
static bool g_fast_prune = true;     // IVF scan: drop (query, probed list) pairs that cannot reach the query's threshold
static int64_t g_fast_probe_pairs = 0;
// profiling counters that live on the device (a search makes no host round trip for them): [pruned (query, list) pairs,
// rows they held, units / items / row tiles of the IVF main candidate pass]
enum { DS_PRUNED_PAIRS = 0, DS_PRUNED_ROWS = 1, DS_TC_UNITS = 2, DS_TC_ITEMS = 3, DS_TC_TILES = 4, DS_TC_HALF_UNITS = 5,
       DS_TC_NARROW_UNITS = 6, DS_TC_NARROW_ITEMS = 7, DS_TC_NARROW_SLOTS = 8, DS_COUNT = 9 };
static DevBuf g_dev_stats;
static unsigned long long *dev_stats() {
    const bool fresh = g_dev_stats.p == nullptr;
    unsigned long long *p = g_dev_stats.as<unsigned long long>(DS_COUNT);
    if (fresh) HB_CUDA(cudaMemsetAsync(p, 0, DS_COUNT * 8, g_stream));
    return p;
}
static double dev_stat(int i) {
    unsigned long long h[DS_COUNT];
    HB_CUDA(cudaMemcpyAsync(h, dev_stats(), sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    HB_CUDA(cudaStreamSynchronize(g_stream));
    return (double)h[i];
}
static bool g_prof_coarse = false;   // diagnostics: the stage timers (pack / tc / select / re-score) record the COARSE job of an IVF search
                                     // instead of its list scan
static bool g_fast_set_only = true;  // IVF coarse routing proves the probed set only (FastJob::set_only)
static bool g_fast_dense = true;     // short flat scans (<= 2048 rows) select from the dumped score matrix
static int g_fast_level_min = 33;     // flat scans of at least this many row tiles run in levels (fast_topk)
static int g_fast_level_dense = 8;    // ... whose first level dumps the scores of this many row tiles (<= 16) instead of emitting them; 0: off
static int g_fast_level_ratio = 16;   // each level covers (ratio - 1) x the tiles before it: it emits about kk x (ratio - 1) rows per query
static int g_fast_sample_tiles = 2;  // IVF: row tiles of the nearest list scored by the threshold-seeding pass
static int64_t g_fast_queries = 0;   // queries answered in FAST mode ...
static int64_t g_fast_fallbacks = 0; // ... of which recomputed by the exact path (proof failed)
static int64_t g_coarse_fallbacks = 0;  // sharded coarse routing: (query, rank) pairs whose slice ranking took the exact kernels
static int64_t g_hnsw_scored = 0;    // (query, row) pairs scored by the HNSW search since "profile" was set
static int64_t g_hnsw_overflows = 0; // queries re-run with their candidate queue in global memory
static int g_hnsw_cand_cap = 0;      // shared-memory candidate queue slots (0: 4 * ef, at least 256)
static int g_micro_batch = 64;       // hb_set_option("micro_batch", n): queries per combined batch of concurrent small calls, 0 = off
static int64_t g_mb_batches = 0, g_mb_requests = 0;
static int g_comm_p2p = 1;           // hb_set_option("comm_p2p", 0): exchange the local top-k by ncclAllGather instead of peer windows (2: the peer-window kernel even with one rank)

}  // namespace hb

using namespace hb;

struct hb_index {
    int type = HB_INDEX_FLAT, dtype = HB_F32, metric = HB_COSINE, d = 0;
    int64_t n = 0;
    int mode = -1;        // hb_index_set_mode: HB_MODE_EXACT / HB_MODE_FAST for searches of this index, -1 = the process default
    int64_t id_base = 0;  // multi-GPU: global row of local row 0 (hb_sharded_search returns id_base + local row)
    // multi-GPU row shard of a global IVF-FLAT index (hb_sharded_ivf_build / hb_index_set_coarse_sharded): the coarse routing
    // of hb_sharded_search is split over the ranks too — rank r ranks the centroids [r nlist/G, (r+1) nlist/G) exactly and the
    // per-rank top-nprobe lists are exchanged and merged like the final results
    bool coarse_sharded = false;
    int cents_slice0 = 0, cents_slice1 = 0;  // the slice fast_cents was built over (0, 0: all centroids)
    // rows as given (flat / hnsw) or list-major slab (ivf)
    DevBuf rows, norms;
    // ivf
    int nlist = 0;
    int64_t max_list = 0;
    DevBuf cents, cent_norm, list_off, list_rows, assign;
    // hnsw
    int max_level = 0, entry = -1;
    DevBuf levels;
    std::vector<DevBuf> adj_off, adj_ids;
    DevBuf adj_off_ptrs, adj_ids_ptrs;                                   // device arrays of the per-level pointers
    DevBuf hn_visited, hn_vlist, hn_ctrl, hn_over, hn_gcand_d, hn_gcand_i;  // search state (hb_hnsw.cu), grown on demand
    int64_t hn_slots = 0;
    // HB_MODE_FAST: digit images of the rows (all index types) and of the centroids (IVF), built on first use
    hb::FastSideBufs fast_rows, fast_cents;
    DevBuf radius;  // IVF, built on first use: per list the largest angle between a row and the centroid (probe pruning)
    bool radius_built = false;
    int64_t device_bytes() const {
        size_t b = rows.cap + norms.cap + cents.cap + cent_norm.cap + list_off.cap + list_rows.cap + assign.cap + levels.cap;
        for (auto &x : adj_off) b += x.cap;
        for (auto &x : adj_ids) b += x.cap;
        b += adj_off_ptrs.cap + adj_ids_ptrs.cap + hn_visited.cap + hn_vlist.cap + hn_ctrl.cap + hn_over.cap + hn_gcand_d.cap +
             hn_gcand_i.cap;
        b += fast_rows.bytes() + fast_cents.bytes() + radius.cap;
        return (int64_t)b;
    }
    void release() {
        rows.release();
        norms.release();
        cents.release();
        cent_norm.release();
        list_off.release();
        list_rows.release();
        assign.release();
        levels.release();
        for (auto &x : adj_off) x.release();
        for (auto &x : adj_ids) x.release();
        for (DevBuf *x : {&adj_off_ptrs, &adj_ids_ptrs, &hn_visited, &hn_vlist, &hn_ctrl, &hn_over, &hn_gcand_d, &hn_gcand_i}) x->release();
        hn_slots = 0;
        fast_rows.release();
        fast_cents.release();
        radius.release();
        radius_built = false;
    }
};

namespace hb {

static void fill_empty_results(int64_t *ids, double *dist, int64_t count) {
    // small helper on the host side of the staging buffers: used only for degenerate calls (n == 0)
    std::vector<int64_t> hi((size_t)count, -1);
    std::vector<double> hd((size_t)count, INFINITY);
    if (ids) HB_CUDA(cudaMemcpyAsync(ids, hi.data(), count * 8, cudaMemcpyHostToDevice, g_stream));
    if (dist) HB_CUDA(cudaMemcpyAsync(dist, hd.data(), count * 8, cudaMemcpyHostToDevice, g_stream));
    sync_stream();
}

// uploads the trivial one-list plan {list_off, lq_off, tile_prefix} for a dense rows x queries scan
static void dense_plan(int64_t nrows, int64_t nq, int64_t *&list_off, int64_t *&lq_off, int64_t *&tile_prefix, DevBuf &buf) {
    int64_t h[6] = {0, nrows, 0, nq, 0, ceil_div(nrows, kTileRows) * ceil_div(nq, kTileQ)};
    int64_t *dptr = buf.as<int64_t>(6);
    HB_CUDA(cudaMemcpyAsync(dptr, h, sizeof(h), cudaMemcpyHostToDevice, g_stream));
    list_off = dptr;
    lq_off = dptr + 2;
    tile_prefix = dptr + 4;
}

// Exact flat search over rows [n x d] (device) — compute-exact-knn, src/hnsw/bench.clj:72-84, and
// top-k-distances, src/hnsw/simd_optimized.clj:271-280.  ids/dist are device buffers [nq x k].
static void flat_search_exact(const void *rows, int rdtype, const double *row_norm, int64_t n, int d, int metric,
                              const void *queries, int qdtype, int64_t nq, int k, int64_t *ids, double *dist) {
    if (nq == 0 || k == 0) return;
    if (n == 0) {
        fill_empty_results(ids, dist, nq * k);
        return;
    }
    bool l2;
    int epi;
    metric_to_epi(metric, false, l2, epi);
    const double *qn = nullptr;
    if (metric == HB_COSINE) {
        double *q = g_ws.qnorm.as<double>(nq);
        launch_row_norms(queries, qdtype, nq, d, q);
        qn = q;
    }
    // rows per pass so that the distance scratch [nq x rc] stays inside the budget
    int64_t rc = (int64_t)(scratch_budget() / 8 / (size_t)nq);
    rc = std::max<int64_t>(kTileRows, rc / kTileRows * kTileRows);
    rc = std::min(rc, n);
    const int64_t npass = ceil_div(n, rc);
    // few queries + long rows: split each query's range so the select kernel fills the machine
    int nsub = 1;
    int64_t sub_len = rc;
    if (nq < 2 * g_num_sms && rc > 32768) {
        nsub = (int)std::min<int64_t>(ceil_div(2 * g_num_sms, nq), ceil_div(rc, 16384));
        sub_len = ceil_div(rc, nsub);
        nsub = (int)ceil_div(rc, sub_len);
    }
    if (nq <= kSmallScanQ && rc > select_small_max()) {
        // a small batch: sub-ranges short enough for the shared-memory selection (select_small_kernel), as long as possible
        // so that the merge of their [nsub x k] survivors fits it too
        const int64_t sl = (rc / 2048 + 1) * k <= select_small_max() ? 2048 : select_small_max();
        if (nq * ceil_div(rc, sl) <= 4 * g_num_sms) {
            sub_len = sl;
        } else {  // too many sub-ranges for one wave of CTAs: the streaming selection over ~4 CTAs per SM
            sub_len = ceil_div(rc, std::min<int64_t>(ceil_div(4 * g_num_sms, nq), ceil_div(rc, 1024)));
        }
        nsub = (int)ceil_div(rc, sub_len);
    }
    const int64_t parts = npass * nsub;
    double *scratch = g_ws.scratch.as<double>((size_t)nq * rc);
    double *cval = parts > 1 ? g_ws.cand_val.as<double>((size_t)nq * parts * k) : dist;
    int64_t *cid = parts > 1 ? g_ws.cand_id.as<int64_t>((size_t)nq * parts * k) : ids;
    int64_t *plan = g_ws.plan.as<int64_t>((size_t)6 * npass);
    const size_t esz = dtype_size(rdtype);
    std::vector<int64_t> hplan((size_t)6 * npass);
    for (int64_t pass = 0; pass < npass; ++pass) {
        const int64_t len = std::min(rc, n - pass * rc);
        int64_t *h = &hplan[(size_t)6 * pass];
        h[0] = 0, h[1] = len, h[2] = 0, h[3] = nq, h[4] = 0, h[5] = ceil_div(len, kTileRows) * ceil_div(nq, kTileQ);
    }
    HB_CUDA(cudaMemcpyAsync(plan, hplan.data(), hplan.size() * 8, cudaMemcpyHostToDevice, g_stream));
    for (int64_t pass = 0; pass < npass; ++pass) {
        const int64_t r0 = pass * rc, len = std::min(rc, n - r0);
        ScanParams S;
        S.rows = (const char *)rows + (size_t)r0 * d * esz;
        S.row_norm = row_norm ? row_norm + r0 : nullptr;
        S.queries = queries;
        S.q_norm = qn;
        S.d = d;
        S.nlist = 1;
        S.list_off = plan + 6 * pass;
        S.lq_off = plan + 6 * pass + 2;
        S.tile_prefix = plan + 6 * pass + 4;
        S.out_stride = len;
        S.out = scratch;
        S.epi = epi;
        {
            Prof pr(PROF_SCAN);
            launch_pairscan(S, rdtype, qdtype, l2, (int)std::min<int64_t>(nq, 1 << 20));
        }
        Prof pr2(PROF_SELECT);
        SelectParams L;
        L.vals = scratch;
        L.nseg = nq;
        L.seg_stride = len;
        L.seg_len_const = len;
        L.nsub = nsub;
        L.sub_len = sub_len;
        L.out_seg_stride = parts;
        L.out_slot_base = pass * nsub;
        L.k = k;
        L.out_val = cval;
        L.out_pos = cid;
        launch_select(L);
        // positions within this pass -> global row ids
        launch_offset_ids(cid + (size_t)pass * nsub * k, nq, (int64_t)nsub * k, parts * k, r0);
    }
    if (parts > 1) {
        // candidates of a query are laid out [pass][sub][k]: ascending row blocks, each sorted by
        // (distance, row) — so (distance, position) order here is (distance, row) order again
        SelectParams L;
        L.vals = cval;
        L.nseg = nq;
        L.seg_stride = parts * k;
        L.seg_len_const = parts * k;
        L.k = k;
        L.out_val = dist;
        int64_t *pos = g_ws.sel_pos.as<int64_t>((size_t)nq * k);
        L.out_pos = pos;
        launch_select(L);
        launch_lookup_ids(pos, nq, k, cid, parts * k, ids);
    }
}

// search-ivf-flat, src/hnsw/ann/partition/ivf_flat.clj:236-294, for a batch of queries.
// probes_out (optional, device) [nq x nprobe] int32.
static void ivf_search_exact(hb_index *ix, const void *queries, int qdtype, int64_t nq, int k, int nprobe,
                             int64_t *ids, double *dist, int32_t *probes_out) {
    if (nq == 0) return;
    const int d = ix->d, nlist = ix->nlist;
    if (ix->n == 0 || nlist == 0) {
        if (k > 0) fill_empty_results(ids, dist, nq * k);
        if (probes_out) HB_CUDA(cudaMemsetAsync(probes_out, 0xFF, (size_t)nq * nprobe * 4, g_stream));
        return;
    }
    const int np_eff = std::min(nprobe, nlist);
    HB_REQUIRE(np_eff >= 1, "num-probes must be >= 1");
    // queries in chunks so the candidate scratch stays inside the budget
    int64_t qc = (int64_t)(scratch_budget() / 8 / ((size_t)np_eff * (size_t)std::max<int64_t>(ix->max_list, 1)));
    qc = std::max<int64_t>(1, std::min(qc, nq));
    qc = std::min<int64_t>(qc, ((1ll << 31) - 1) / np_eff);
    const size_t qsz = dtype_size(qdtype);

    double *qn = g_ws.qnorm.as<double>(nq);
    launch_row_norms(queries, qdtype, nq, d, qn);  // :275-278
    int64_t *plan = g_ws.plan.as<int64_t>(6);
    double *coarse = g_ws.cand_val.as<double>((size_t)qc * nlist);
    double *pval = g_ws.sel_val.as<double>((size_t)qc * np_eff);
    int64_t *ppos = g_ws.sel_pos.as<int64_t>((size_t)qc * std::max(np_eff, k));
    int32_t *probes = g_ws.probes.as<int32_t>((size_t)qc * np_eff);
    int64_t *pair_out = g_ws.pair_out.as<int64_t>((size_t)qc * np_eff + 1);
    int32_t *qsel = g_ws.qsel.as<int32_t>((size_t)qc * np_eff);
    int64_t *lq_off = g_ws.lq_off.as<int64_t>(nlist + 1);
    int64_t *tile_prefix = g_ws.tile_prefix.as<int64_t>(nlist + 1);
    bool cl2;
    int cepi;
    metric_to_epi(ix->metric == HB_IP ? HB_COSINE : ix->metric, true, cl2, cepi);

    for (int64_t q0 = 0; q0 < nq; q0 += qc) {
        const int64_t nqc = std::min(qc, nq - q0);
        const void *qptr = (const char *)queries + (size_t)q0 * d * qsz;
        // coarse quantiser (:261-269): distance-fn(query, centroid) for all centroids, stable sort, take
        {
            int64_t h[6] = {0, nlist, 0, nqc, 0, ceil_div(nlist, kTileRows) * ceil_div(nqc, kTileQ)};
            HB_CUDA(cudaMemcpyAsync(plan, h, sizeof(h), cudaMemcpyHostToDevice, g_stream));
            ScanParams S;
            S.rows = ix->cents.p;
            S.row_norm = (const double *)ix->cent_norm.p;
            S.queries = qptr;
            S.q_norm = qn + q0;
            S.d = d;
            S.nlist = 1;
            S.list_off = plan;
            S.lq_off = plan + 2;
            S.tile_prefix = plan + 4;
            S.out_stride = nlist;
            S.out = coarse;
            S.epi = cepi;
            Prof pr(PROF_COARSE);
            launch_pairscan(S, HB_F64, qdtype, cl2, (int)std::min<int64_t>(nqc, 1 << 20));
            SelectParams L;
            L.vals = coarse;
            L.nseg = nqc;
            L.seg_stride = nlist;
            L.seg_len_const = nlist;
            L.k = np_eff;
            L.out_val = pval;
            L.out_pos = ppos;
            launch_select(L);
        }
        const int64_t np = nqc * np_eff;
        {
            Prof prp(PROF_PLAN);
            ivf_plan(ppos, np, nlist, (const int64_t *)ix->list_off.p, probes, pair_out, qsel, lq_off, tile_prefix,
                     kTileRows, kTileQ, g_ws.tmp);
        }
        if (probes_out) {
            if (np_eff == nprobe) {
                HB_CUDA(cudaMemcpyAsync(probes_out + (size_t)q0 * nprobe, probes, (size_t)np * 4, cudaMemcpyDeviceToDevice, g_stream));
            } else {
                HB_CUDA(cudaMemsetAsync(probes_out + (size_t)q0 * nprobe, 0xFF, (size_t)nqc * nprobe * 4, g_stream));
                HB_CUDA(cudaMemcpy2DAsync(probes_out + (size_t)q0 * nprobe, (size_t)nprobe * 4, probes, (size_t)np_eff * 4,
                                          (size_t)np_eff * 4, (size_t)nqc, cudaMemcpyDeviceToDevice, g_stream));
            }
        }
        if (k == 0) continue;
        // candidates of the chunk: read back to size the scratch, unless the bound nprobe x longest list is small anyway
        // (small batches: one host round trip less per call)
        int64_t total = np * std::max<int64_t>(ix->max_list, 1);
        if (total > (1ll << 22)) {
            HB_CUDA(cudaMemcpyAsync(&total, pair_out + np, 8, cudaMemcpyDeviceToHost, g_stream));
            sync_stream();
        }
        double *scratch = g_ws.scratch.as<double>((size_t)std::max<int64_t>(total, 1));
        // list scan (:217-234): always cosine via precomputed norms, no zero guard; lists in probe order (:281-288)
        ScanParams S;
        S.rows = ix->rows.p;
        S.row_norm = (const double *)ix->norms.p;
        S.queries = qptr;
        S.q_norm = qn + q0;
        S.d = d;
        S.nlist = nlist;
        S.list_off = (const int64_t *)ix->list_off.p;
        S.lq_off = lq_off;
        S.qsel = qsel;
        S.pair_div = np_eff;
        S.pair_out = pair_out;
        S.tile_prefix = tile_prefix;
        S.out = scratch;
        S.epi = EPI_COS;
        {
            Prof pr(PROF_SCAN);
            // a query probes a list once; a small batch's queries are rows 0..nqc-1 of qptr
            launch_pairscan(S, ix->dtype, qdtype, false, (int)std::min<int64_t>(nqc, 1 << 20), nqc <= kSmallScanQ ? (int)nqc : 0);
        }
        Prof prs(PROF_SELECT);
        // merge (:291-294): stable sort of the concatenation in probe order, take k
        SelectParams L;
        L.vals = scratch;
        L.nseg = nqc;
        L.seg_off = pair_out;
        L.seg_off_stride = np_eff;
        L.k = k;
        L.out_val = dist + (size_t)q0 * k;
        L.out_pos = ppos;
        const int64_t max_seg = (int64_t)np_eff * std::max<int64_t>(ix->max_list, 1);
        if (nqc <= kSmallScanQ && max_seg > select_small_max()) {
            // a small batch: one CTA per concatenation is latency-bound; short sub-ranges first (their k best in
            // (distance, position) order, select_small_kernel), then the same selection over the [nsub x k] survivors,
            // which are again in position order among equal distances
            int64_t sub_len = (max_seg / 2048 + 1) * k <= select_small_max() ? 2048 : select_small_max();
            if (nqc * ceil_div(max_seg, sub_len) > 4 * g_num_sms)  // more than a wave: the streaming selection, ~4 CTAs per SM
                sub_len = ceil_div(max_seg, std::min<int64_t>(ceil_div(4 * g_num_sms, nqc), ceil_div(max_seg, 1024)));
            const int nsub = (int)ceil_div(max_seg, sub_len);
            double *cval = g_ws.misc3.as<double>((size_t)nqc * nsub * k);
            int64_t *cpos = g_ws.misc4.as<int64_t>((size_t)nqc * nsub * k);
            L.nsub = nsub;
            L.sub_len = sub_len;
            L.out_seg_stride = nsub;
            L.out_val = cval;
            L.out_pos = cpos;
            launch_select(L);
            SelectParams M;
            M.vals = cval;
            M.nseg = nqc;
            M.seg_stride = (int64_t)nsub * k;
            M.seg_len_const = (int64_t)nsub * k;
            M.k = k;
            M.out_val = dist + (size_t)q0 * k;
            int64_t *mpos = g_ws.misc2.as<int64_t>((size_t)nqc * k);
            M.out_pos = mpos;
            launch_select(M);
            launch_lookup_ids(mpos, nqc, k, cpos, (int64_t)nsub * k, ppos);
        } else {
            launch_select(L);
        }
        launch_ivf_resolve(ppos, nqc, k, np_eff, probes, pair_out, (const int64_t *)ix->list_off.p,
                           (const int64_t *)ix->list_rows.p, ids + (size_t)q0 * k);
    }
}

// partition-vectors-kmeans, src/hnsw/ann/partition/ivf_flat.clj:92-131, on device rows.
// cents [nlist x d] fp64 (device), assign [n] int32 (device); seeds (device int64[nlist]) are filled by
// k-means++ unless given.
// assign-to-nearest-centroid over all rows (ivf_flat.clj:79-90); defined with the FAST-mode orchestration below
static void assign_rows(const void *rows, int dtype, const double *row_norm, int64_t n, int d, const double *cents,
                        const double *cnorm, int nlist, bool l2, int32_t *assign);

// kmeans-plus-plus-init (ivf_flat.clj:32-60; linear: Lightning's d_i-weighted walk, lightning.clj:86-109): `seeds` receives
// the nlist chosen rows (device int64).  norm = fp64 row norms (device).
static bool g_kpp_scale = true;  // hb_set_option("kpp_scale", 0): the one-thread prefix walk + exhaustive distance pass (same seeds)
static int64_t g_kpp_scored = 0, g_kpp_walked = 0, g_kpp_steps = 0;
static DevBuf g_kpp_buf;
static void kpp_seeds(const void *rows, int dtype, int64_t n, int d, const double *norm, bool l2, int nlist, int64_t seed, bool linear,
                      int64_t *seeds) {
    JavaRandom rng(seed);
    std::vector<double> u((size_t)nlist);
    const int64_t first = rng.next_int((int32_t)n);
    for (int t = 1; t < nlist; ++t) u[t] = rng.next_double();
    // (the staged update kernel keeps the newest seed, widened, in shared memory: rows of more than ~20,000 dimensions take
    //  the plain path)
    if (g_kpp_scale && (size_t)d * 8 + 40 * 1024 <= 200 * 1024) {
        // hb_kpp.cu: the same picks, with the distance pass pruned by the triangle inequality and the ordered fp64 sum done
        // as integer adds per chunk
        const int64_t nchunks = ceil_div(n, kKppChunk);
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        size_t off = 0;
        auto take = [&](size_t bytes) {
            const size_t o = off;
            off += al(bytes);
            return o;
        };
        const size_t o_u = take((size_t)nlist * 8), o_mind = take((size_t)n * 8), o_near = take((size_t)n * 4), o_theta = take((size_t)n * 4),
                     o_bound = take((size_t)nlist * 8), o_ap = take((size_t)nchunks * 8), o_pf = take((size_t)nchunks * 8),
                     o_st = take((size_t)(nchunks + 1) * 8), o_ex = take((size_t)nchunks * 4), o_q = take((size_t)nchunks * 8),
                     o_misc = take(64);
        char *base = (char *)g_kpp_buf.get(off);
        KppScaleParams K;
        K.rows = rows;
        K.dtype = dtype;
        K.n = n;
        K.d = d;
        K.row_norm = norm;
        K.l2 = l2;
        K.linear = linear;
        K.mind = (double *)(base + o_mind);
        K.near = (int32_t *)(base + o_near);
        K.theta = (float *)(base + o_theta);
        K.seeds = seeds;
        K.seeds_rw = seeds;
        K.bound = (double *)(base + o_bound);
        K.c_approx = (double *)(base + o_ap);
        K.c_prefix = (double *)(base + o_pf);
        K.c_start = (double *)(base + o_st);
        K.c_exp = (int *)(base + o_ex);
        K.c_q = (long long *)(base + o_q);
        K.total = (double *)(base + o_misc);
        K.pick = (int64_t *)(base + o_misc + 8);
        unsigned long long *counters = (unsigned long long *)(base + o_misc + 16);
        if (g_profile) {
            HB_CUDA(cudaMemsetAsync(counters, 0, 16, g_stream));
            K.n_scored = counters;
            K.n_walked = counters + 1;
        }
        double *ud = (double *)(base + o_u);
        HB_CUDA(cudaMemcpyAsync(ud, u.data(), (size_t)nlist * 8, cudaMemcpyHostToDevice, g_stream));
        HB_CUDA(cudaMemcpyAsync(seeds, &first, 8, cudaMemcpyHostToDevice, g_stream));
        launch_kpp_init_state(n, K.mind, K.near, K.theta);
        for (int t = 1; t < nlist; ++t) {
            K.t = t;
            K.u = ud + t;
            launch_kpp_scale_step(K);
        }
        if (g_profile) {
            unsigned long long h[2];
            HB_CUDA(cudaMemcpyAsync(h, counters, 16, cudaMemcpyDeviceToHost, g_stream));
            sync_stream();
            g_kpp_scored += (int64_t)h[0];
            g_kpp_walked += (int64_t)h[1];
            g_kpp_steps += nlist - 1;
        }
        sync_stream();  // u / first (host) were sources of async copies
        return;
    }
    double *ud = g_ws.misc3.as<double>((size_t)nlist + 2 + 2 * (size_t)n);
    double *total = ud + nlist;
    double *mind = ud + nlist + 2;
    double *cum = mind + n;
    int64_t *pick = g_ws.misc4.as<int64_t>(1);
    HB_CUDA(cudaMemcpyAsync(ud, u.data(), (size_t)nlist * 8, cudaMemcpyHostToDevice, g_stream));
    HB_CUDA(cudaMemcpyAsync(pick, &first, 8, cudaMemcpyHostToDevice, g_stream));
    HB_CUDA(cudaMemcpyAsync(seeds, &first, 8, cudaMemcpyHostToDevice, g_stream));
    launch_fill_f64(mind, n, DBL_MAX);
    KppParams K;
    K.rows = rows;
    K.dtype = dtype;
    K.n = n;
    K.d = d;
    K.row_norm = norm;
    K.l2 = l2;
    K.linear = linear;
    K.mind = mind;
    K.cum = cum;
    K.total = total;
    K.pick = pick;
    for (int t = 1; t < nlist; ++t) {
        K.u = ud + t;
        K.out_seed = seeds + t;
        launch_kpp_step(K);
    }
    sync_stream();  // u (host vector) must outlive the copy; also bounds queue depth
}

static void kmeans_exact(const void *rows, int dtype, int64_t n, int d, int metric, int nlist, int iters, int64_t seed,
                         const int64_t *seed_rows_dev, double *cents, int32_t *assign, int64_t *seeds_out_dev) {
    HB_REQUIRE(n >= 1 && nlist >= 1, "k-means needs at least one row and one partition");
    HB_REQUIRE(n < (1ll << 31), "n must be < 2^31");
    HB_REQUIRE(metric == HB_COSINE || metric == HB_L2, "k-means distance-fn must be cosine or euclidean");
    const bool l2 = metric == HB_L2;
    double *norm = g_ws.misc.as<double>(n);
    launch_row_norms(rows, dtype, n, d, norm);
    int64_t *seeds = seeds_out_dev ? seeds_out_dev : g_ws.misc2.as<int64_t>(nlist);
    if (seed_rows_dev) {
        if (seeds != seed_rows_dev) HB_CUDA(cudaMemcpyAsync(seeds, seed_rows_dev, (size_t)nlist * 8, cudaMemcpyDeviceToDevice, g_stream));
    } else {
        kpp_seeds(rows, dtype, n, d, norm, l2, nlist, seed, false, seeds);
    }
    launch_init_centroids(rows, dtype, d, seeds, nlist, cents);
    double *cnorm = g_ws.misc3.as<double>((size_t)nlist + 2 + 2 * (size_t)n);  // reuse: [nlist] norms
    int64_t *list_off = g_ws.lq_off.as<int64_t>(nlist + 1);
    int64_t *list_rows = g_ws.cand_id.as<int64_t>(n);
    for (int it = 0; it <= iters; ++it) {
        launch_row_norms(cents, HB_F64, nlist, d, cnorm);
        {
            Prof pr(PROF_ASSIGN);
            assign_rows(rows, dtype, norm, n, d, cents, cnorm, nlist, l2, assign);
        }
        if (it == iters) break;  // final assignment pass (:119-131)
        build_lists(assign, n, nlist, list_off, list_rows, g_ws.tmp);
        launch_update_centroids(rows, dtype, d, list_off, list_rows, nlist, cents, nullptr, nullptr);
    }
}

// partition-vectors-kmeans with the rows sharded over the ranks (contiguous global row blocks, this rank holds
// [first_row, first_row + n)): centroids replicated; per round every rank assigns its rows, sums its members per cluster in
// row order, one all-reduce(sum) of [nlist x d] fp64 sums + [nlist] counts, divide (an empty cluster keeps its centroid).
// Everything is enqueued on the library's stream: no host synchronisation inside a round.  The all-reduce regroups the
// reference's row-order sum, so a centroid can differ from the single-GPU build in the last ulp.
static void kmeans_sharded(const void *rows, int dtype, int64_t n, int d, int metric, int nlist, int iters,
                           const int64_t *seed_rows_dev, int64_t first_row, double *cents, int32_t *assign) {
    HB_REQUIRE(n >= 1 && nlist >= 1, "k-means needs at least one row per shard and one partition");
    HB_REQUIRE(n < (1ll << 31), "rows per shard must be < 2^31");
    HB_REQUIRE(metric == HB_COSINE || metric == HB_L2, "k-means distance-fn must be cosine or euclidean");
    HB_REQUIRE(seed_rows_dev, "sharded k-means needs seed rows");
    const bool l2 = metric == HB_L2;
    double *norm = g_ws.misc.as<double>(n);
    launch_row_norms(rows, dtype, n, d, norm);
    // centroids start as data rows (ivf_flat.clj:40,58): the owner of a seed row contributes it, the others zeros
    launch_init_centroids_sharded(rows, dtype, d, seed_rows_dev, nlist, first_row, n, cents);
    comm_allreduce_sum_f64(cents, (int64_t)nlist * d);
    double *cnorm = g_ws.misc3.as<double>((size_t)nlist + 2 + 2 * (size_t)n);
    int64_t *list_off = g_ws.lq_off.as<int64_t>(nlist + 1);
    int64_t *list_rows = g_ws.cand_id.as<int64_t>(n);
    double *sums = g_ws.sh_sums.as<double>((size_t)nlist * d);
    int64_t *counts = g_ws.sh_counts.as<int64_t>((size_t)nlist);
    for (int it = 0; it <= iters; ++it) {
        launch_row_norms(cents, HB_F64, nlist, d, cnorm);
        {
            Prof pr(PROF_ASSIGN);
            assign_rows(rows, dtype, norm, n, d, cents, cnorm, nlist, l2, assign);
        }
        if (it == iters) break;
        {
            Prof pr(PROF_UPDATE);
            build_lists(assign, n, nlist, list_off, list_rows, g_ws.tmp);
            launch_update_centroids(rows, dtype, d, list_off, list_rows, nlist, nullptr, sums, counts);
        }
        Prof pr(PROF_ALLREDUCE);
        comm_allreduce_sum_f64(sums, (int64_t)nlist * d);
        comm_allreduce_sum_i64(counts, nlist);
        launch_divide_centroids(sums, counts, nlist, d, cents);
    }
}

// Turns (rows, centroids, assignments) on the device into the list-major index layout.
static void ivf_finalize(hb_index *ix, const void *rows_dev, const double *row_norm_dev) {
    const int64_t n = ix->n;
    const int d = ix->d, nlist = ix->nlist;
    int64_t *list_off = ix->list_off.as<int64_t>(nlist + 1);
    int64_t *list_rows = ix->list_rows.as<int64_t>(std::max<int64_t>(n, 1));
    build_lists((const int32_t *)ix->assign.p, n, nlist, list_off, list_rows, g_ws.tmp);
    void *slab = ix->rows.get(std::max<size_t>((size_t)n * d * dtype_size(ix->dtype), 16));
    double *snorm = ix->norms.as<double>(std::max<int64_t>(n, 1));
    launch_gather_rows(rows_dev, ix->dtype, d, list_rows, n, slab, row_norm_dev, snorm);
    double *cn = ix->cent_norm.as<double>(nlist);
    launch_row_norms(ix->cents.p, HB_F64, nlist, d, cn);
    std::vector<int64_t> off((size_t)nlist + 1);
    HB_CUDA(cudaMemcpyAsync(off.data(), list_off, off.size() * 8, cudaMemcpyDeviceToHost, g_stream));
    sync_stream();
    ix->max_list = 0;
    for (int l = 0; l < nlist; ++l) ix->max_list = std::max(ix->max_list, off[l + 1] - off[l]);
}


// =================================================================================================
// HB_MODE_FAST orchestration (scheme: hb_fast.cuh)
// =================================================================================================
struct FastWs {
    DevBuf dig, q64, pslot, srow, ptotal, qu, ql1, qscale, qeps, qmargin, thr, cnt, cnegv, crel, cpos, selval, selpos, pq, pr, exact;
    DevBuf aimg, aimg0, u_list, u_sel0, u_nsel, u_ntile, u_item0, u_item0n, u_tile0, u_slotq, u_slotrel;
    DevBuf t_list, t_sel0, t_nsel, t_ntile, t_item0, t_item0n, t_slotq, t_slotrel;
    DevBuf a_ids, a_dist, a_norm, a_tmp, dump, timing, simub, pruned;
    DevBuf sc_pos, sc_dist, sc_gdist, ok_c;
    DevBuf ppos, ppos0, ok_a, ok_b, relk, probes0, qsel0, lq_off0, uprefix0, uprefix, flat_plan, idx, gq, gids, gdist,
        tmp2;
    void release() {
        DevBuf *all[] = {&dig, &q64, &pslot, &srow, &ptotal, &qu, &ql1, &qscale, &qeps, &qmargin, &thr, &cnt, &cnegv, &crel, &cpos, &selval, &selpos, &pq, &pr, &exact,
                         &aimg, &aimg0, &u_list, &u_sel0, &u_nsel, &u_ntile, &u_item0, &u_item0n, &u_tile0, &u_slotq, &u_slotrel,
                         &t_list, &t_sel0, &t_nsel, &t_ntile, &t_item0, &t_item0n, &t_slotq, &t_slotrel,
                         &sc_pos, &sc_dist, &sc_gdist, &ok_c, &ppos, &ppos0, &ok_a, &ok_b, &relk, &probes0, &qsel0, &lq_off0, &uprefix0, &uprefix, &flat_plan,
                         &idx, &gq, &gids, &gdist, &tmp2, &a_ids, &a_dist, &a_norm, &a_tmp, &dump, &timing, &simub, &pruned};
        for (DevBuf *b : all) b->release();
    }
};
static FastWs g_fw;

constexpr int kFastMaxK = 112;  // largest k served by the candidate pass
// candidates selected (and at most re-scored) per query / candidate slots per query
static int fast_kk(int k) { return k <= 48 ? 64 : 128; }
static int fast_cap(int k) { return k <= 48 ? 2048 : 4096; }

static void build_fast_side(FastSideBufs &S, const void *rows, int dtype, int d, int nlist, const int64_t *list_off_dev,
                            const double *norm_dev, int ns) {
    S.kbn = (int)ceil_div(d, kFastKB);
    S.ns = ns;
    int64_t *toff = S.tile_off.as<int64_t>((size_t)nlist + 1);
    launch_tile_offsets(list_off_dev, nlist, toff);
    int64_t nt = 0, nr = 0;
    HB_CUDA(cudaMemcpyAsync(&nt, toff + nlist, 8, cudaMemcpyDeviceToHost, g_stream));
    HB_CUDA(cudaMemcpyAsync(&nr, list_off_dev + nlist, 8, cudaMemcpyDeviceToHost, g_stream));
    sync_stream();
    S.ntiles = nt;
    S.nrows = nr;
    const size_t img_bytes = (size_t)nt * S.kbn * ns * kFastImg;
    void *img = S.img.get(std::max<size_t>(img_bytes, 16));
    HB_CUDA(cudaMemsetAsync(img, 0, std::max<size_t>(img_bytes, 16), g_stream));
    float *rs = S.rs.as<float>((size_t)std::max<int64_t>(nt * kFastTile, 1));
    float *ro = S.ro.as<float>((size_t)std::max<int64_t>(nt * kFastTile, 1));
    HB_CUDA(cudaMemsetAsync(rs, 0, (size_t)std::max<int64_t>(nt * kFastTile, 1) * 4, g_stream));
    launch_fill_f32(ro, nt * kFastTile, -INFINITY);
    float *stats = S.stats.as<float>(4);
    HB_CUDA(cudaMemsetAsync(stats, 0, 16, g_stream));
    launch_quant_rows(rows, dtype, d, S.kbn, ns, nlist, list_off_dev, toff, norm_dev, (int8_t *)img, rs, ro, stats);
    float hs[4] = {0, 0, 0, 0};
    HB_CUDA(cudaMemcpyAsync(hs, stats, 16, cudaMemcpyDeviceToHost, g_stream));
    sync_stream();
    S.usable = hs[2] == 0.0f && std::isfinite(hs[0]) && std::isfinite(hs[1]);
    S.built = true;
}

// rows side of any index (IVF: slab lists; flat/HNSW: one list of all rows)
static FastSideBufs &fast_rows_side(hb_index *ix, bool cosine) {
    FastSideBufs &S = ix->fast_rows;
    if (S.built && S.ns == g_fast_ns) return S;
    S.release();
    if (ix->type == HB_INDEX_IVF_FLAT) {
        build_fast_side(S, ix->rows.p, ix->dtype, ix->d, ix->nlist, (const int64_t *)ix->list_off.p, (const double *)ix->norms.p,
                        g_fast_ns);
    } else {
        int64_t h[2] = {0, ix->n};
        int64_t *lo = S.list_off.as<int64_t>(2);
        HB_CUDA(cudaMemcpyAsync(lo, h, 16, cudaMemcpyHostToDevice, g_stream));
        sync_stream();
        build_fast_side(S, ix->rows.p, ix->dtype, ix->d, 1, lo, cosine ? (const double *)ix->norms.p : nullptr, g_fast_ns);
    }
    return S;
}
// digit images of the centroids [c0, c1) (default: all of them)
static FastSideBufs &fast_cents_side(hb_index *ix, int c0 = 0, int c1 = -1) {
    if (c1 < 0) c1 = ix->nlist;
    FastSideBufs &S = ix->fast_cents;
    if (S.built && S.ns == g_fast_ns && ix->cents_slice0 == c0 && ix->cents_slice1 == c1) return S;
    S.release();
    int64_t h[2] = {0, c1 - c0};
    int64_t *lo = S.list_off.as<int64_t>(2);
    HB_CUDA(cudaMemcpyAsync(lo, h, 16, cudaMemcpyHostToDevice, g_stream));
    sync_stream();
    build_fast_side(S, (const double *)ix->cents.p + (size_t)c0 * ix->d, HB_F64, ix->d, 1, lo, (const double *)ix->cent_norm.p + c0,
                    g_fast_ns);
    ix->cents_slice0 = c0;
    ix->cents_slice1 = c1;
    return S;
}

static const double *list_radius(hb_index *ix) {
    double *r = ix->radius.as<double>((size_t)std::max(ix->nlist, 1));
    if (!ix->radius_built) {
        launch_list_radius(ix->rows.p, ix->dtype, (const double *)ix->norms.p, (const double *)ix->cents.p,
                           (const double *)ix->cent_norm.p, (const int64_t *)ix->list_off.p, ix->nlist, ix->n, ix->d, r);
        ix->radius_built = true;
    }
    return r;
}

struct FastPlan {  // which (query, list) pairs a pass covers
    int nlist = 0;
    const int64_t *lq_off = nullptr;       // [nlist+1] selections per list
    const int64_t *unit_prefix = nullptr;  // [nlist+1] exclusive prefix of ceil(selections / 128)
    const int32_t *qsel = nullptr;         // selection -> pair (NULL: identity)
    const int64_t *pair_out = nullptr;     // pair -> offset of its segment (NULL: flat)
    int pair_div = 0;                      // query = pair / pair_div (0: pair == query)
    int nunits = 0;       // unit slots (interleaved layouts pad the real count up to a multiple of the SM count)
    int nunits_real = 0;
    int interleave = 0;
    int tile_limit = 0, tile_div = 0, tile_start = 0;
};
// pads the unit count so that the slots can be interleaved across the persistent CTAs (unit_plan_kernel)
static void set_units(FastPlan &F, int64_t real, bool interleave) {
    F.nunits_real = (int)real;
    F.interleave = (interleave && real > 2 * g_num_sms) ? g_num_sms : 0;
    F.nunits = F.interleave ? (int)(ceil_div(real, F.interleave) * F.interleave) : (int)real;
}
// The host only knows an upper bound of the unit count (the plan lives on the device): ceil(pairs / 128) full units plus
// one partial unit per list that has a selection.  unit_plan_kernel reads the real count; slots beyond it stay empty.
static void set_units_bound(FastPlan &F, int64_t pairs, int nlist, bool interleave) {
    const int64_t bound = ceil_div(pairs, kFastTile) + std::min<int64_t>(nlist, pairs);
    F.interleave = (interleave && bound > 2 * g_num_sms) ? g_num_sms : 0;
    F.nunits = F.interleave ? (int)(ceil_div(bound, F.interleave) * F.interleave) : (int)bound;
    F.nunits_real = -1;
}
struct FastJob {
    FastSideBufs *side = nullptr;
    const int64_t *list_off = nullptr;  // slab rows per list (B side)
    const void *rows_exact = nullptr;   // for the fp64 re-score
    int rdtype = HB_F32;
    const double *row_norm = nullptr;
    const void *queries = nullptr;
    int qdtype = HB_F32;
    const double *q64 = nullptr;  // the queries widened to fp64 (launch_widen_queries)
    int64_t nq = 0;
    const double *qn = nullptr;
    int d = 0;
    int metric = HB_COSINE;  // HB_COSINE or HB_IP: how scores relate to the exact distance
    int epi = EPI_COS;
    int k = 0;
    FastPlan emit, thresh;
    bool shared_units = false;  // thresh covers the same units as emit (fewer tiles)
    bool profile = true;        // record the per-stage events (the coarse job is timed as a whole by its caller)
    bool cover_stats = false;   // count what the main candidate pass covers ("tc_units" ... stats; the IVF list scan)
    // The caller needs the exact top-k SET only (IVF coarse routing: which lists to probe): candidates that are in it by
    // their bounds alone are not re-scored, out_rel lists them first, out_dist is not meaningful.
    bool set_only = false;
    const int64_t *tie_list_off = nullptr;  // see FinalParams
    int tie_nlist = 0;
    double *out_simub = nullptr;  // see FinalParams
    // 0: the whole job.  IVF list scan with probe pruning: 1 = query bounds + threshold sample only (needs `thresh`),
    // 2 = everything after it (needs `emit`, planned after the pruning; thresholds and bounds stay from phase 1)
    int phase = 0;
    int64_t *out_rel = nullptr;
    double *out_dist = nullptr;
    int32_t *out_ok = nullptr;
};

static UnitPlan make_units(const FastPlan &F, DevBuf &b_list, DevBuf &b_sel0, DevBuf &b_nsel, DevBuf &b_ntile, DevBuf &b_item0,
                           DevBuf &b_slotq, DevBuf &b_slotrel, const int64_t *tile_off, DevBuf *b_item0n = nullptr,
                           DevBuf *b_tile0 = nullptr, int skip_tiles = 0, unsigned long long *stats = nullptr) {
    UnitPlan U;
    const size_t nu = (size_t)std::max(F.nunits, 1);
    if (b_item0n) U.unit_item0n = b_item0n->as<int32_t>(nu + 1);
    if (b_tile0) U.unit_tile0 = b_tile0->as<int32_t>(nu);
    U.unit_list = b_list.as<int32_t>(nu);
    U.unit_sel0 = b_sel0.as<int32_t>(nu);
    U.unit_nsel = b_nsel.as<int32_t>(nu);
    U.unit_ntile = b_ntile.as<int32_t>(nu + 1);
    U.unit_item0 = b_item0.as<int32_t>(nu + 1);
    U.slot_query = b_slotq.as<int32_t>(nu * kFastTile);
    U.slot_rel0 = b_slotrel.as<int32_t>(nu * kFastTile);
    launch_unit_plan(F.nlist, F.lq_off, F.unit_prefix, tile_off, F.nunits, F.nunits_real, F.interleave, F.tile_limit, F.tile_div,
                     F.tile_start, F.qsel, F.pair_out, F.pair_div, nullptr, U, skip_tiles, stats);
    return U;
}

// queries must already be quantised into g_fw.dig / qu / ql1 (fast_quant_queries)
static void fast_quant_queries(const void *queries, int qdtype, int64_t nq, int d) {
    const int kbn = (int)ceil_div(d, kFastKB);
    int8_t *dig = g_fw.dig.as<int8_t>((size_t)nq * g_fast_ns * kbn * kFastKB);
    launch_quant_queries(queries, qdtype, nq, d, kbn, g_fast_ns, dig, g_fw.qu.as<double>(nq), g_fw.ql1.as<double>(nq));
}

static void fast_topk(const FastJob &J) {
    FastWs &W = g_fw;
    const FastSideBufs &S = *J.side;
    const int ns = S.ns, kbn = S.kbn, kk = fast_kk(J.k), cap = fast_cap(J.k);
    const int64_t nq = J.nq;
    const bool do_sample = J.phase != 2, do_main = J.phase != 1;
    // IVF list scans (their own threshold plan, EMIT passes, thresholds seeded by the sample): sparse units on tc_narrow_kernel
    // ... and a levelled flat scan of at most 32 queries (one narrow unit): 9..32 queries over a long list then cost one pass
    // over the digit images at HBM speed instead of the fp64 tiles (thresholds rise between the levels, not inside them)
    const bool leveled_flat = J.phase == 0 && J.shared_units && J.emit.nlist == 1 && S.ntiles >= g_fast_level_min;
    const bool narrow = g_tc_narrow && ns == 2 && (!J.shared_units || (leveled_flat && J.nq <= kNarrowSlots));
    // IVF list scans: the sample pass scores the first tiles of every query's nearest list; its kk best candidates are KEPT
    // (compacted in place) and the main pass does not read those tiles again for the units that only hold nearest-list queries
    const bool carry = g_fast_carry && !J.shared_units && J.thresh.tile_limit > 0 && J.thresh.tile_div <= 1;
    const int skip_tiles = carry ? J.thresh.tile_limit : 0;
    HB_REQUIRE(J.phase == 0 || !J.shared_units, "phased jobs need their own threshold plan");
    double *qscale = W.qscale.as<double>(nq), *qeps = W.qeps.as<double>(nq);
    float *qmargin = W.qmargin.as<float>(nq);
    float *thr = W.thr.as<float>(nq);
    int32_t *cnt = W.cnt.as<int32_t>(nq);
    double *cnegv = W.cnegv.as<double>((size_t)nq * cap);
    int32_t *crel = W.crel.as<int32_t>((size_t)nq * cap);
    int32_t *cpos = W.cpos.as<int32_t>((size_t)nq * cap);
    if (do_sample) {
        launch_query_bounds((const double *)W.qu.p, (const double *)W.ql1.p, J.qn, nq, ns, J.d, J.metric, (const float *)S.stats.p, qscale,
                            qeps, qmargin, thr, cnt);  // also thr <- -inf, cnt <- 0
    }

    UnitPlan U, T;
    int8_t *aimg = nullptr, *aimg0 = nullptr;
    {
        Prof pr(J.profile ? PROF_PACK : -1);
        if (do_main) {
            U = make_units(J.emit, W.u_list, W.u_sel0, W.u_nsel, W.u_ntile, W.u_item0, W.u_slotq, W.u_slotrel, (const int64_t *)S.tile_off.p,
                           narrow ? &W.u_item0n : nullptr, carry ? &W.u_tile0 : nullptr, skip_tiles,
                           (g_profile && J.cover_stats) ? dev_stats() + DS_TC_UNITS : nullptr);
            aimg = W.aimg.as<int8_t>((size_t)std::max(J.emit.nunits, 1) * kbn * ns * kFastImg);
            // an IVF list scan runs EMIT passes only: its M = 64 units (<= 64 selections) need half an image, its narrow
            // units (<= 32) a quarter
            launch_pack_units((const int8_t *)W.dig.p, kbn, ns, J.emit.nunits, U.slot_query, aimg,
                              (narrow || (g_tc_half_m && !J.shared_units)) ? U.unit_nsel : nullptr, narrow);
        }
        if (!do_sample) {
        } else if (J.shared_units) {
            T = U;
            T.unit_ntile = W.t_ntile.as<int32_t>((size_t)J.emit.nunits + 1);
            T.unit_item0 = W.t_item0.as<int32_t>((size_t)J.emit.nunits + 1);
            launch_unit_plan(J.thresh.nlist, J.thresh.lq_off, J.thresh.unit_prefix, (const int64_t *)S.tile_off.p, J.thresh.nunits,
                             J.thresh.nunits_real, J.thresh.interleave, J.thresh.tile_limit, J.thresh.tile_div, 0, J.thresh.qsel,
                             J.thresh.pair_out, J.thresh.pair_div, nullptr, T);
            aimg0 = aimg;
        } else {
            T = make_units(J.thresh, W.t_list, W.t_sel0, W.t_nsel, W.t_ntile, W.t_item0, W.t_slotq, W.t_slotrel,
                           (const int64_t *)S.tile_off.p, narrow ? &W.t_item0n : nullptr);
            aimg0 = W.aimg0.as<int8_t>((size_t)std::max(J.thresh.nunits, 1) * kbn * ns * kFastImg);
            launch_pack_units((const int8_t *)W.dig.p, kbn, ns, J.thresh.nunits, T.slot_query, aimg0,
                              (g_tc_half_m || narrow) ? T.unit_nsel : nullptr, narrow);
        }
    }
    TcParams P;
    P.bimg = (const int8_t *)S.img.p;
    P.kbn = kbn;
    P.tile_off = (const int64_t *)S.tile_off.p;
    P.list_off = J.list_off;
    P.rs = (const float *)S.rs.p;
    P.ro = (const float *)S.ro.p;
    P.thr = thr;
    P.margin = qmargin;
    P.k = J.k;
    P.kk = kk;
    P.cap = cap;
    if (g_fast_debug) {
        P.timing = (long long *)W.timing.get((size_t)g_num_sms * 8 * 8 * 2);  // [dense kernel | narrow kernel]
        HB_CUDA(cudaMemsetAsync(P.timing, 0, (size_t)g_num_sms * 8 * 8 * 2, g_stream));
    }
    P.cnt = cnt;
    P.cand_negv = cnegv;
    P.cand_rel = crel;
    P.cand_pos = cpos;
    double *selval = W.selval.as<double>((size_t)nq * kk);
    int64_t *selpos = W.selpos.as<int64_t>((size_t)nq * kk);
    auto select_candidates = [&] {
        Prof pr(J.profile ? PROF_SELECT : -1);
        launch_cand_select(cnegv, cnt, nq, kk, cap, selval, selpos);
    };
    // A long flat scan (one list, many row tiles) runs in levels over growing tile ranges [0,2), [2,32), [32,512), ...:
    // after each level the candidate lists are cut back to their kk best and the thresholds rise to the exact kk-th best
    // (k-th best - margin) so far, so a level emits about kk * 15 rows per query however long the scan is.
    const bool leveled = leveled_flat;
    // A short flat scan (coarse routing, k-means assignment over <= 2048 centroids) dumps its whole score matrix and lets
    // one warp per query pick the candidates from it: no thresholds, no sample pass, no sort.
    const int64_t flat_rows = J.emit.nlist == 1 ? S.nrows : 0;
    const bool dense = J.phase == 0 && g_fast_dense && J.shared_units && J.emit.nlist == 1 && flat_rows >= 1 && flat_rows <= 2048 && !leveled;
    if (dense) {
        Prof pr(J.profile ? PROF_TC : -1);
        P.aimg = aimg;
        P.nunits = J.emit.nunits;
        P.unit_list = U.unit_list;
        P.unit_nsel = g_tc_half_m ? U.unit_nsel : nullptr;
        P.unit_item0 = U.unit_item0;
        P.slot_query = U.slot_query;
        P.slot_rel0 = U.slot_rel0;
        P.tile_stride = 1;
        P.items_hint = (int64_t)J.emit.nunits * S.ntiles;
        P.dump = W.dump.as<float>((size_t)J.emit.nunits * S.ntiles * kFastTile * kFastTile);
        launch_tc_pass(P, ns, FAST_DUMP);
        P.items_hint = 0;
        launch_dense_select(P.dump, (int)S.ntiles, nq, (int)flat_rows, J.k, kk, cap, qmargin, cnegv, crel, cpos, cnt, thr, selval,
                            selpos);
    } else if (leveled) {
        P.aimg = aimg;
        P.nunits = J.emit.nunits;
        P.unit_list = U.unit_list;
        P.unit_nsel = g_tc_half_m ? U.unit_nsel : nullptr;
        P.slot_query = U.slot_query;
        P.slot_rel0 = U.slot_rel0;
        P.tile_stride = 1;
        UnitPlan Lp = U;
        Lp.unit_ntile = W.t_ntile.as<int32_t>((size_t)J.emit.nunits + 1);
        Lp.unit_item0 = W.t_item0.as<int32_t>((size_t)J.emit.nunits + 1);
        if (narrow) Lp.unit_item0n = W.t_item0n.as<int32_t>((size_t)J.emit.nunits + 1);
        P.unit_nsel_all = U.unit_nsel;
        int64_t t0 = 0;
        if (!narrow && g_fast_level_dense > 0) {
            // level 0 without thresholds would push every score of its tiles through the candidate lists (atomics, then a sort
            // of all of them per query): dump the scores of the first tiles instead and let one warp per query pick from the
            // matrix, as the short flat scans do — 8 tiles cost less than the 2 emitted ones did, and the next level starts
            // from the k-th best of 1024 rows instead of 256
            const int64_t t1 = std::min<int64_t>(S.ntiles, g_fast_level_dense);
            {
                Prof pr(J.profile ? PROF_PLAN : -1);
                launch_unit_plan(J.emit.nlist, J.emit.lq_off, J.emit.unit_prefix, (const int64_t *)S.tile_off.p, J.emit.nunits,
                                 J.emit.nunits_real, J.emit.interleave, (int)t1, 1, 0, J.emit.qsel, J.emit.pair_out, J.emit.pair_div,
                                 nullptr, Lp);
            }
            Prof pr(J.profile ? PROF_TC_SAMPLE : -1);
            P.unit_item0 = Lp.unit_item0;
            P.tile_start = 0;
            P.items_hint = (int64_t)J.emit.nunits * t1;
            P.dump = W.dump.as<float>((size_t)J.emit.nunits * t1 * kFastTile * kFastTile);
            launch_tc_pass(P, ns, FAST_DUMP);
            launch_dense_select(P.dump, (int)t1, nq, (int)std::min<int64_t>(S.nrows, t1 * kFastTile), J.k, kk, cap, qmargin, cnegv, crel,
                                cpos, cnt, thr, selval, selpos);
            t0 = t1;
        }
        while (t0 < S.ntiles) {
            const int64_t t1 = t0 == 0 ? 2 : std::min<int64_t>(S.ntiles, t0 * g_fast_level_ratio);
            {
                Prof pr(J.profile ? PROF_PLAN : -1);
                launch_unit_plan(J.emit.nlist, J.emit.lq_off, J.emit.unit_prefix, (const int64_t *)S.tile_off.p, J.emit.nunits,
                                 J.emit.nunits_real, J.emit.interleave, (int)(t1 - t0), 1, (int)t0, J.emit.qsel, J.emit.pair_out,
                                 J.emit.pair_div, nullptr, Lp);
            }
            P.unit_item0 = Lp.unit_item0;
            P.unit_item0n = Lp.unit_item0n;
            P.tile_start = (int)t0;
            P.items_hint = (int64_t)J.emit.nunits * (t1 - t0);
            {
                Prof pr(J.profile ? (t0 == 0 ? PROF_TC_SAMPLE : PROF_TC) : -1);
                launch_tc_pass(P, ns, FAST_EMIT);
            }
            {
                Prof pr(J.profile ? PROF_SELECT : -1);
                launch_cand_compact(cnegv, crel, cpos, cnt, nq, kk, cap, J.k, qmargin, thr, selval, selpos);
            }
            t0 = t1;
        }
        P.tile_start = 0;
        P.unit_item0n = nullptr;
        P.items_hint = 0;
    } else {
    if (do_sample) {
    {
        // sample pass: candidates of a subset of the rows (nearest list / every 8th tile) -> their kk-th best
        // score seeds the thresholds of the full pass
        Prof pr(J.profile ? PROF_TC_SAMPLE : -1);
        P.aimg = aimg0;
        P.nunits = J.thresh.nunits;
        P.unit_list = T.unit_list;
        P.unit_nsel = g_tc_half_m ? T.unit_nsel : nullptr;
        P.unit_item0 = T.unit_item0;
        P.unit_item0n = T.unit_item0n;
        P.unit_nsel_all = T.unit_nsel;
        P.slot_query = T.slot_query;
        P.slot_rel0 = T.slot_rel0;
        P.tile_stride = J.thresh.tile_div > 1 ? J.thresh.tile_div : 1;
        if (J.shared_units && J.emit.nlist == 1) P.items_hint = (int64_t)J.thresh.nunits * ceil_div(S.ntiles, P.tile_stride);
        launch_tc_pass(P, ns, FAST_EMIT);
        P.items_hint = 0;
    }
    if (carry) {
        // keep the kk best sample candidates in place (cnt = their number) and raise the thresholds to the kk-th best / the
        // k-th best - margin: what thr_from_sample did, without throwing the candidates away
        Prof pr(J.profile ? PROF_SELECT : -1);
        launch_cand_compact(cnegv, crel, cpos, cnt, nq, kk, cap, J.k, qmargin, thr, selval, selpos);
    } else {
        select_candidates();
        launch_thr_from_sample(selval, cnt, nq, kk, cap, J.k, qmargin, thr);
        HB_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nq * 4, g_stream));
    }
    }
    if (!do_main) return;
    {
        Prof pr(J.profile ? PROF_TC : -1);
        P.tile_stride = 1;
        P.aimg = aimg;
        P.nunits = J.emit.nunits;
        P.unit_list = U.unit_list;
        P.unit_nsel = g_tc_half_m ? U.unit_nsel : nullptr;
        P.unit_item0 = U.unit_item0;
        P.unit_item0n = U.unit_item0n;
        P.unit_nsel_all = U.unit_nsel;
        P.unit_tile0 = U.unit_tile0;
        P.skip_tiles = skip_tiles;
        P.slot_query = U.slot_query;
        P.slot_rel0 = U.slot_rel0;
        if (J.shared_units && J.emit.nlist == 1) P.items_hint = (int64_t)J.emit.nunits * S.ntiles;
        launch_tc_pass(P, ns, FAST_EMIT);
        P.items_hint = 0;
        P.unit_tile0 = nullptr;
        P.skip_tiles = 0;
    }
    select_candidates();
    }
    double *exact = W.exact.as<double>((size_t)nq * kk);
    {
        Prof pr(J.profile ? PROF_RESCORE : -1);
        int32_t *pq = W.pq.as<int32_t>((size_t)nq * kk), *prow = W.pr.as<int32_t>((size_t)nq * kk);
        int32_t *pslot = W.pslot.as<int32_t>((size_t)nq * kk), *srow = W.srow.as<int32_t>((size_t)nq * kk);
        int32_t *ptotal = W.ptotal.as<int32_t>(1);
        launch_rescore_pairs(selpos, selval, cpos, nq, kk, cap, J.k, qmargin, srow, exact, ptotal, pq, prow, pslot, J.set_only);
        launch_rescore(J.rows_exact, J.rdtype, J.row_norm, J.q64, J.qdtype == HB_F32, J.qn, J.d, pq, prow, pslot, ptotal, nq * kk, J.epi,
                       exact, J.set_only);
        FinalParams F;
        F.nq = nq;
        F.k = J.k;
        F.kk = kk;
        F.cap = cap;
        F.sel_pos = selpos;
        F.sel_negv = selval;
        F.exact = exact;
        F.pair_row = srow;
        F.cand_rel = crel;
        F.cnt = cnt;
        F.thr = thr;
        F.q_scale = qscale;
        F.q_eps = qeps;
        F.metric = J.metric;
        F.out_rel = J.out_rel;
        F.out_dist = J.out_dist;
        F.out_ok = J.out_ok;
        F.tie_list_off = J.tie_list_off;
        F.tie_nlist = J.tie_nlist;
        F.out_simub = J.out_simub;
        launch_fast_final(F);
    }
    if (g_fast_debug && P.timing) {
        std::vector<long long> ht((size_t)g_num_sms * 8);
        HB_CUDA(cudaMemcpy(ht.data(), P.timing, ht.size() * 8, cudaMemcpyDeviceToHost));
        double a[6] = {0, 0, 0, 0, 0, 0};
        for (int c = 0; c < g_num_sms; ++c)
            for (int j = 0; j < 6; ++j) a[j] += (double)ht[(size_t)c * 8 + j] / g_num_sms;
        fprintf(stderr, "[hb fast] last tc pass, mean per CTA: items %.1f, mma warp %.0f clk (wait accumulators free %.0f, wait operands %.0f), "
                        "epilogue: wait accumulators %.0f, phase A %.0f\n", a[3], a[0], a[1], a[2], a[4], a[5]);
        if (narrow) {
            std::vector<long long> hn((size_t)g_num_sms * 8);
            HB_CUDA(cudaMemcpy(hn.data(), P.timing + (size_t)g_num_sms * 8, hn.size() * 8, cudaMemcpyDeviceToHost));
            double b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int c = 0; c < g_num_sms; ++c)
                for (int j = 0; j < 8; ++j) b[j] += (double)hn[(size_t)c * 8 + j] / g_num_sms;
            fprintf(stderr, "[hb fast] last narrow pass, mean per CTA: items %.1f, mma warp %.0f clk (wait accumulators free %.0f, wait operands %.0f), "
                            "epilogue: wait accumulators %.0f, per-item setup %.0f, scores + emission %.0f; producer waits for a free stage %.0f\n",
                    b[3], b[0], b[1], b[2], b[4], b[5], b[6], b[7]);
        }
    }
    if (g_fast_debug) {
        std::vector<int32_t> hc((size_t)nq), hok((size_t)nq);
        std::vector<float> ht((size_t)nq);
        std::vector<double> he((size_t)nq), hs((size_t)nq), hsel((size_t)nq * kk), hex((size_t)nq * kk);
        HB_CUDA(cudaMemcpyAsync(hc.data(), cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaMemcpyAsync(hok.data(), J.out_ok, (size_t)nq * 4, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaMemcpyAsync(ht.data(), thr, (size_t)nq * 4, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaMemcpyAsync(he.data(), qeps, (size_t)nq * 8, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaMemcpyAsync(hs.data(), qscale, (size_t)nq * 8, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaMemcpyAsync(hsel.data(), selval, (size_t)nq * kk * 8, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaMemcpyAsync(hex.data(), exact, (size_t)nq * kk * 8, cudaMemcpyDeviceToHost, g_stream));
        sync_stream();
        std::vector<int32_t> hpr((size_t)nq * kk);
        HB_CUDA(cudaMemcpy(hpr.data(), W.srow.p, (size_t)nq * kk * 4, cudaMemcpyDeviceToHost));
        int64_t nres = 0;
        for (int32_t r : hpr) nres += r >= 0;
        fprintf(stderr, "[hb fast] re-scored %.1f of %d selected candidates per query\n", (double)nres / nq, kk);
        int64_t nok = 0, over = 0, cmin = 1 << 30, cmax = 0, csum = 0;
        for (int64_t q = 0; q < nq; ++q) {
            nok += hok[q];
            over += hc[q] > cap;
            cmin = std::min<int64_t>(cmin, hc[q]);
            cmax = std::max<int64_t>(cmax, hc[q]);
            csum += hc[q];
        }
        fprintf(stderr, "[hb fast] nq=%lld k=%d units=%d/%d ok=%lld overflow=%lld cand min/mean/max=%lld/%.1f/%lld\n", (long long)nq,
                J.k, J.thresh.nunits, J.emit.nunits, (long long)nok, (long long)over, (long long)cmin, (double)csum / nq, (long long)cmax);
        for (int64_t q = 0; q < std::min<int64_t>(nq, 3); ++q) {
            double kth = 0;
            std::vector<double> ex(hex.begin() + q * kk, hex.begin() + (q + 1) * kk);
            std::sort(ex.begin(), ex.end());
            kth = ex[std::min(J.k, kk) - 1];
            fprintf(stderr, "  q%lld cnt=%d thr=%g scale=%g eps=%g sel[kk-1]=%g -> t_sim=%g exact kth dist=%g ok=%d\n", (long long)q, hc[q],
                    ht[q], hs[q], he[q], hsel[q * kk + kk - 1], -hsel[q * kk + kk - 1] * hs[q], kth, hok[q]);
        }
    }
}

// single-list plan {lq_off = [0, nq], unit_prefix = [0, ceil(nq/128)]} on the device
static void flat_fast_plan(int64_t nq, FastPlan &E, FastPlan &T, DevBuf &buf) {
    int64_t h[4] = {0, nq, 0, ceil_div(nq, kFastTile)};
    int64_t *dptr = buf.as<int64_t>(4);
    launch_set_i64x4(dptr, h[0], h[1], h[2], h[3]);  // kernel arguments: no host buffer to keep alive, no sync
    E.nlist = 1;
    E.lq_off = dptr;
    E.unit_prefix = dptr + 2;
    set_units(E, h[3], false);
    T = E;
    T.tile_div = 8;
}

// Re-runs `exact_fn(sub_queries, nsub, sub_ids, sub_dist)` for the queries whose proof failed and scatters the results.
template <typename F>
static void fast_fallback(const int32_t *ok_dev, const void *queries, int qdtype, int64_t nq, int d, int k, int64_t *ids,
                          double *dist, F &&exact_fn) {
    // the failed queries are compacted on the device; the host reads one word (how many) to size the exact call
    int32_t *idx = g_fw.idx.as<int32_t>((size_t)nq + 1);
    launch_compact_failed(ok_dev, nq, idx);
    int32_t nbad = 0;
    HB_CUDA(cudaMemcpyAsync(&nbad, idx + nq, 4, cudaMemcpyDeviceToHost, g_stream));
    sync_stream();
    g_fast_queries += nq;
    g_fast_fallbacks += nbad;
    if (nbad == 0) return;
    const int64_t nb = nbad;
    const size_t qsz = dtype_size(qdtype);
    void *gq = g_fw.gq.get((size_t)nb * d * qsz);
    launch_gather_bytes(queries, idx, nb, (int64_t)d * qsz, gq);
    int64_t *gids = g_fw.gids.as<int64_t>((size_t)nb * k);
    double *gdist = g_fw.gdist.as<double>((size_t)nb * k);
    exact_fn(gq, nb, gids, gdist);
    launch_scatter_rows64(gids, idx, nb, k, ids);
    launch_scatter_rows64(gdist, idx, nb, k, dist);
}

static bool fast_metric_ok(int metric) { return metric == HB_COSINE || metric == HB_IP; }

// compute-exact-knn / top-k-distances in FAST mode
static void flat_search_fast(hb_index *ix, const void *queries, int qdtype, int64_t nq, int k, int64_t *ids, double *dist) {
    const bool cosine = ix->metric == HB_COSINE;
    if (nq == 0 || k == 0) return;
    // One query, or a few over rows that fit a couple of hundred MB: the HBM-bound exact scan (rowstream_kernel) is the faster
    // path.  2..8 queries over a long list read half the bytes as digit images on the levelled narrow scan (1M x 768: 8
    // queries 1.34 ms exact, 0.56 ms here).
    const bool long_list = (size_t)ix->n * ix->d * dtype_size(ix->dtype) >= ((size_t)1 << 30) && ceil_div(ix->n, kFastTile) >= g_fast_level_min;
    if (ix->n == 0 || k > kFastMaxK || !fast_metric_ok(ix->metric) || ix->n >= (1ll << 31) || nq < 2 || (nq <= kSmallScanQ && !long_list)) {
        flat_search_exact(ix->rows.p, ix->dtype, (const double *)ix->norms.p, ix->n, ix->d, ix->metric, queries, qdtype, nq, k, ids, dist);
        return;
    }
    FastSideBufs &S = fast_rows_side(ix, cosine);
    if (!S.usable) {
        flat_search_exact(ix->rows.p, ix->dtype, (const double *)ix->norms.p, ix->n, ix->d, ix->metric, queries, qdtype, nq, k, ids, dist);
        return;
    }
    const int d = ix->d;
    const size_t qsz = dtype_size(qdtype);
    const int64_t qc = 16384;
    for (int64_t q0 = 0; q0 < nq; q0 += qc) {
        const int64_t nqc = std::min(qc, nq - q0);
        const void *qptr = (const char *)queries + (size_t)q0 * d * qsz;
        double *qn = g_ws.qnorm.as<double>(nqc);
        launch_row_norms(qptr, qdtype, nqc, d, qn);
        fast_quant_queries(qptr, qdtype, nqc, d);
        const double *q64 = launch_widen_queries(qptr, qdtype, nqc * d, g_fw.q64.as<double>((size_t)nqc * d));
        FastJob J;
        J.side = &S;
        J.list_off = (const int64_t *)S.list_off.p;
        J.rows_exact = ix->rows.p;
        J.rdtype = ix->dtype;
        J.row_norm = cosine ? (const double *)ix->norms.p : nullptr;
        J.queries = qptr;
        J.qdtype = qdtype;
        J.q64 = q64;
        J.nq = nqc;
        J.qn = qn;
        J.d = d;
        J.metric = ix->metric;
        J.epi = cosine ? EPI_COS : EPI_NEGDOT;
        J.k = k;
        flat_fast_plan(nqc, J.emit, J.thresh, g_fw.flat_plan);
        J.shared_units = true;
        J.out_rel = ids + (size_t)q0 * k;
        J.out_dist = dist + (size_t)q0 * k;
        J.out_ok = g_fw.ok_a.as<int32_t>(nq) + q0;
        fast_topk(J);
    }
    fast_fallback((const int32_t *)g_fw.ok_a.p, queries, qdtype, nq, d, k, ids, dist,
                  [&](const void *gq, int64_t nb, int64_t *gids, double *gdist) {
                      flat_search_exact(ix->rows.p, ix->dtype, (const double *)ix->norms.p, ix->n, ix->d, ix->metric, gq, qdtype, nb,
                                        k, gids, gdist);
                  });
}

// assign-to-nearest-centroid (ivf_flat.clj:79-90) over all rows.  EXACT mode: the fp64 tile kernel.  FAST mode
// (cosine): the rows are the queries of a k = 1 candidate pass over the centroids (quantised afresh for every Lloyd
// round), the surviving centroids are re-scored in fp64 and a row whose nearest centroid is not proven falls back to
// the fp64 kernel -- same assignments, ties to the lowest centroid index (strict <, :87-89).
static FastSideBufs g_assign_side;
static void assign_rows(const void *rows, int dtype, const double *row_norm, int64_t n, int d, const double *cents,
                        const double *cnorm, int nlist, bool l2, int32_t *assign) {
    auto exact = [&](const void *r, const double *rn, int64_t cnt, int32_t *out) {
        AssignParams A;
        A.rows = r;
        A.row_norm = rn;
        A.n = cnt;
        A.d = d;
        A.cents = cents;
        A.cent_norm = cnorm;
        A.nlist = nlist;
        A.epi = l2 ? EPI_L2 : EPI_COS_GUARD;
        A.out_assign = out;
        launch_assign(A, dtype, l2);
    };
    const bool try_fast = g_mode == HB_MODE_FAST && !l2 && (dtype == HB_F32 || dtype == HB_F64) && nlist >= 2 * kFastTile && n >= 1;
    FastSideBufs &S = g_assign_side;
    if (try_fast) {
        int64_t h[2] = {0, nlist};
        int64_t *lo = S.list_off.as<int64_t>(2);
        HB_CUDA(cudaMemcpyAsync(lo, h, 16, cudaMemcpyHostToDevice, g_stream));
        sync_stream();
        build_fast_side(S, cents, HB_F64, d, 1, lo, cnorm, g_fast_ns);
    }
    if (!try_fast || !S.usable) {
        exact(rows, row_norm, n, assign);
        return;
    }
    const size_t esz = dtype_size(dtype);
    int64_t *ids64 = g_fw.a_ids.as<int64_t>(n);
    double *dist64 = g_fw.a_dist.as<double>(n);
    int32_t *ok = g_fw.ok_a.as<int32_t>(n);
    const int64_t qc = 16384;
    for (int64_t q0 = 0; q0 < n; q0 += qc) {
        const int64_t nqc = std::min(qc, n - q0);
        const void *qptr = (const char *)rows + (size_t)q0 * d * esz;
        fast_quant_queries(qptr, dtype, nqc, d);
        const double *q64 = launch_widen_queries(qptr, dtype, nqc * d, g_fw.q64.as<double>((size_t)nqc * d));
        FastJob J;
        J.side = &S;
        J.list_off = (const int64_t *)S.list_off.p;
        J.rows_exact = cents;
        J.rdtype = HB_F64;
        J.row_norm = cnorm;
        J.queries = qptr;
        J.qdtype = dtype;
        J.q64 = q64;
        J.nq = nqc;
        J.qn = row_norm + q0;
        J.d = d;
        J.metric = HB_COSINE;
        J.epi = EPI_COS_GUARD;
        J.k = 1;
        J.profile = false;
        flat_fast_plan(nqc, J.emit, J.thresh, g_fw.flat_plan);
        J.shared_units = true;
        J.out_rel = ids64 + q0;
        J.out_dist = dist64 + q0;
        J.out_ok = ok + q0;
        fast_topk(J);
    }
    fast_fallback(ok, rows, dtype, n, d, 1, ids64, dist64, [&](const void *gq, int64_t nb, int64_t *gids, double *gdist) {
        double *gn = g_fw.a_norm.as<double>(nb);
        int32_t *ga = g_fw.a_tmp.as<int32_t>(nb);
        launch_row_norms(gq, dtype, nb, d, gn);
        exact(gq, gn, nb, ga);
        launch_i32_to_i64(ga, nb, gids);
        HB_CUDA(cudaMemsetAsync(gdist, 0, (size_t)nb * 8, g_stream));
    });
    launch_pos_to_i32(ids64, n, assign);
}

// Host queries arrive in blocks on a copy stream while the per-query stages (norms, digits, coarse routing) of the blocks
// already on the device run: the H2D copy of a 10k x 768 batch is 0.55 ms, as long as the whole coarse stage.
struct HostFeed {
    int64_t block = 0;                // queries per block
    std::vector<cudaEvent_t> ready;   // ready[b]: block b is on the device
};
static cudaStream_t g_copy_stream = nullptr;
// hb_set_option("host_feed", block).  Off by default: measured at configs[1] (10k x 768 fp32 queries) the per-block coarse
// stage costs more than the hidden copy saves (e2e 3.52 ms with one copy, 3.80 ms in blocks of 2048, 4.67 ms in blocks of 1024).
static int64_t g_host_feed_block = 0;

// ---- coarse routing split over the ranks (row shards of ONE global IVF-FLAT index) -------------------------------------------
// Every rank holds all centroids, but ranking nlist = 65,536 centroids for every query on every rank is replicated work that
// grows with the number of GPUs.  The QUERIES are split instead: rank r ranks all centroids for the queries
// [r per, (r + 1) per), per = ceil(nq / G) — EXACTLY: candidate pass, fp64 re-score, proof, exact fallback for the queries that
// fail it, in (distance, centroid index) order, the stable sort of ivf_flat.clj:263-269 — and one all-gather replicates the probe
// lists.  The probe ORDER is then exact as well, so no cross-list tie needs the exact path.  (The first version split the
// CENTROIDS: every rank ranked its slice for all queries and the G local top-nprobe lists were merged.  That repeats the per-query
// stages — selection, 35 fp64 re-scores, proof — on every rank: 1.9 ms per 10,000 queries on 8 GPUs, most of it those stages.)
static bool sharded_coarse_on(const hb_index *ix, int np_eff) {
    const CommInfo &c = comm_info();
    return c.inited && c.nranks > 1 && ix->coarse_sharded && ix->metric == HB_COSINE && np_eff <= kFastMaxK &&
           ix->nlist >= 2 * kFastTile && ix->nlist >= np_eff;
}
// exact top-np of the centroids [c0, c1) for nb queries (device): the reference's arithmetic, ids = position inside the range
static void coarse_slice_exact(hb_index *ix, int c0, int c1, const void *q, int qdtype, const double *qn, int64_t nb, int np,
                               int64_t *out_pos, double *out_dist) {
    const int d = ix->d, ns = c1 - c0;
    int64_t *plan = g_ws.plan.as<int64_t>(6);
    double *coarse = g_ws.cand_val.as<double>((size_t)nb * ns);
    int64_t h[6] = {0, ns, 0, nb, 0, ceil_div(ns, kTileRows) * ceil_div(nb, kTileQ)};
    HB_CUDA(cudaMemcpyAsync(plan, h, sizeof(h), cudaMemcpyHostToDevice, g_stream));
    sync_stream();
    ScanParams Sc;
    Sc.rows = (const double *)ix->cents.p + (size_t)c0 * d;
    Sc.row_norm = (const double *)ix->cent_norm.p + c0;
    Sc.queries = q;
    Sc.q_norm = qn;
    Sc.d = d;
    Sc.nlist = 1;
    Sc.list_off = plan;
    Sc.lq_off = plan + 2;
    Sc.tile_prefix = plan + 4;
    Sc.out_stride = ns;
    Sc.out = coarse;
    Sc.epi = EPI_COS_GUARD;
    launch_pairscan(Sc, HB_F64, qdtype, false);
    SelectParams L;
    L.vals = coarse;
    L.nseg = nb;
    L.seg_stride = ns;
    L.seg_len_const = ns;
    L.k = np;
    L.out_val = out_dist;
    L.out_pos = out_pos;
    launch_select(L);
}
// Per query the global top-np_eff centroids in exact (distance, index) order -> ppos (list ids), simub (upper bound of the
// cosine similarity to each probed centroid).  Collective: every rank calls it with the same queries; qn / q64 hold the norms
// and the widened copy of all nqc queries.  Leaves the digit images of this rank's query block in the workspace: the caller
// quantises the whole batch again before the list scan.
static void sharded_coarse(hb_index *ix, const void *qptr, int qdtype, const double *q64, const double *qn, int64_t nqc, int np_eff,
                           int64_t *ppos, double *simub) {
    const CommInfo &c = comm_info();
    const int nlist = ix->nlist, d = ix->d;
    const int64_t per = ceil_div(nqc, (int64_t)c.nranks);
    const int64_t q_lo = std::min(nqc, per * c.rank), nb = std::min(nqc, q_lo + per) - q_lo;
    const size_t blk = (size_t)per * np_eff;
    FastWs &W = g_fw;
    // [pos | dist] of this rank's block, then the gathered blocks of every rank
    char *mine = (char *)W.sc_pos.get(blk * 16);
    char *all = (char *)W.sc_dist.get(blk * 16 * c.nranks);
    int64_t *lpos = (int64_t *)mine;
    double *ldist = (double *)(mine + blk * 8);
    if (nb < per) HB_CUDA(cudaMemsetAsync(mine, 0, blk * 16, g_stream));  // the padding rows travel too
    if (nb > 0) {
        const size_t qsz = dtype_size(qdtype);
        const void *bptr = (const char *)qptr + (size_t)q_lo * d * qsz;
        FastSideBufs &C = fast_cents_side(ix);
        if (C.usable) {
            fast_quant_queries(bptr, qdtype, nb, d);
            int32_t *ok = W.ok_b.as<int32_t>(nb);
            FastJob J;
            J.side = &C;
            J.list_off = (const int64_t *)C.list_off.p;
            J.rows_exact = ix->cents.p;
            J.rdtype = HB_F64;
            J.row_norm = (const double *)ix->cent_norm.p;
            J.queries = bptr;
            J.qdtype = qdtype;
            J.q64 = q64 + (size_t)q_lo * d;
            J.nq = nb;
            J.qn = qn + q_lo;
            J.d = d;
            J.metric = HB_COSINE;
            J.epi = EPI_COS_GUARD;
            J.profile = false;
            J.k = np_eff;
            flat_fast_plan(nb, J.emit, J.thresh, W.flat_plan);
            J.shared_units = true;
            J.out_rel = lpos;
            J.out_dist = ldist;
            J.out_ok = ok;
            fast_topk(J);
            const int64_t served0 = g_fast_queries, fell0 = g_fast_fallbacks;
            fast_fallback(ok, bptr, qdtype, nb, d, np_eff, lpos, ldist, [&](const void *gq, int64_t ng, int64_t *gids, double *gd) {
                double *gn = g_fw.a_norm.as<double>(ng);
                launch_row_norms(gq, qdtype, ng, d, gn);
                coarse_slice_exact(ix, 0, nlist, gq, qdtype, gn, ng, np_eff, gids, gd);
            });
            g_coarse_fallbacks += g_fast_fallbacks - fell0;  // "fast_queries" / "fast_fallbacks" count whole searches, not this stage
            g_fast_queries = served0;
            g_fast_fallbacks = fell0;
        } else {
            coarse_slice_exact(ix, 0, nlist, bptr, qdtype, qn + q_lo, nb, np_eff, lpos, ldist);
        }
    }
    comm_allgather_bytes(mine, all, (int64_t)(blk * 16));
    launch_unpack_probe_blocks(all, c.nranks, per, nqc, np_eff, ppos, simub);
}

// search-ivf-flat in FAST mode: coarse routing and the probed-list scan both run the candidate pass
static void ivf_search_fast(hb_index *ix, const void *queries, int qdtype, int64_t nq, int k, int nprobe, int64_t *ids,
                            double *dist, const HostFeed *feed = nullptr) {
    if (nq == 0 || k == 0) return;
    auto wait_all = [&] {
        if (feed)
            for (cudaEvent_t e : feed->ready) HB_CUDA(cudaStreamWaitEvent(g_stream, e, 0));
    };
    const int d = ix->d, nlist = ix->nlist;
    const int np_eff = std::min(nprobe, nlist);
    if (ix->n == 0 || nlist == 0 || k > kFastMaxK || ix->n >= (1ll << 31) || np_eff < 1 || nq <= kSmallScanQ) {
        wait_all();  // (small batches: see flat_search_fast)
        ivf_search_exact(ix, queries, qdtype, nq, k, nprobe, ids, dist, nullptr);
        return;
    }
    FastSideBufs &S = fast_rows_side(ix, true);
    const bool shard_coarse = sharded_coarse_on(ix, np_eff);  // the same on every rank
    const size_t qsz = dtype_size(qdtype);
    int64_t qc = std::max<int64_t>(kFastTile, (1ll << 20) / np_eff);
    qc = std::min(qc, nq);
    if (!S.usable) {
        wait_all();
        if (shard_coarse) {
            // this shard's rows cannot take the candidate pass (a zero / non-finite norm), but the other ranks wait for its
            // part of the coarse exchange: contribute it, then answer from the exact path (which ranks all centroids itself)
            for (int64_t q0 = 0; q0 < nq; q0 += qc) {
                const int64_t nqc = std::min(qc, nq - q0);
                const void *qptr = (const char *)queries + (size_t)q0 * d * qsz;
                double *qn = g_ws.qnorm.as<double>(nqc);
                launch_row_norms(qptr, qdtype, nqc, d, qn);
                const double *q64 = launch_widen_queries(qptr, qdtype, nqc * d, g_fw.q64.as<double>((size_t)nqc * d));
                sharded_coarse(ix, qptr, qdtype, q64, qn, nqc, np_eff, g_fw.ppos.as<int64_t>((size_t)nqc * np_eff),
                               g_fw.simub.as<double>((size_t)nqc * np_eff));
            }
        }
        ivf_search_exact(ix, queries, qdtype, nq, k, nprobe, ids, dist, nullptr);
        return;
    }
    const bool fast_coarse = !shard_coarse && ix->metric == HB_COSINE && np_eff <= kFastMaxK && nlist >= 2 * kFastTile;
    FastSideBufs *C = fast_coarse ? &fast_cents_side(ix) : nullptr;
    const bool coarse_tc = shard_coarse || (C && C->usable);
    if (feed) {  // query chunks must start on a block boundary of the feed
        if (qc >= feed->block && qc < nq) qc -= qc % feed->block;
        else if (qc < feed->block) {
            wait_all();
            feed = nullptr;
        }
    }
    FastWs &W = g_fw;
    int32_t *ok_all = W.ok_a.as<int32_t>(nq);
    for (int64_t q0 = 0; q0 < nq; q0 += qc) {
        const int64_t nqc = std::min(qc, nq - q0);
        const void *qptr = (const char *)queries + (size_t)q0 * d * qsz;
        double *qn = g_ws.qnorm.as<double>(nqc);
        double *q64buf = g_fw.q64.as<double>((size_t)nqc * d);
        const double *q64 = qdtype == HB_F64 ? (const double *)qptr : q64buf;
        int64_t *ppos = W.ppos.as<int64_t>((size_t)nqc * np_eff);
        double *simub = W.simub.as<double>((size_t)nqc * np_eff);
        int32_t *ok_c = W.ok_c.as<int32_t>(nqc);
        if (shard_coarse) {
            Prof pr(PROF_COARSE);
            wait_all();
            launch_row_norms(qptr, qdtype, nqc, d, qn);
            launch_widen_queries(qptr, qdtype, nqc * d, q64buf);
            sharded_coarse(ix, qptr, qdtype, q64, qn, nqc, np_eff, ppos, simub);  // (quantises this rank's query block)
            fast_quant_queries(qptr, qdtype, nqc, d);  // the list scan packs its units from the whole batch's digits
            {
                float one_bits;
                const int32_t one = 1;
                memcpy(&one_bits, &one, 4);
                launch_fill_f32((float *)ok_c, nqc, one_bits);  // every probe list is exact: int32 1 in every slot
            }
        } else if (coarse_tc) {
            Prof pr(PROF_COARSE);
            // per-query stages in blocks (the blocks of a host batch as they arrive; a device batch is one block)
            const int64_t blk = (feed && feed->block > 0) ? feed->block : nqc;
            for (int64_t b0 = 0; b0 < nqc; b0 += blk) {
                const int64_t nb = std::min(blk, nqc - b0);
                if (feed) HB_CUDA(cudaStreamWaitEvent(g_stream, feed->ready[(size_t)((q0 + b0) / feed->block)], 0));
                const void *bptr = (const char *)qptr + (size_t)b0 * d * qsz;
                launch_row_norms(bptr, qdtype, nb, d, qn + b0);
                fast_quant_queries(bptr, qdtype, nb, d);
                const double *b64 = launch_widen_queries(bptr, qdtype, nb * d, q64buf + (size_t)b0 * d);
                FastJob J;
                J.side = C;
                J.list_off = (const int64_t *)C->list_off.p;
                J.rows_exact = ix->cents.p;
                J.rdtype = HB_F64;
                J.row_norm = (const double *)ix->cent_norm.p;
                J.queries = bptr;
                J.qdtype = qdtype;
                J.q64 = b64;
                J.nq = nb;
                J.qn = qn + b0;
                J.d = d;
                J.metric = HB_COSINE;
                J.epi = EPI_COS_GUARD;
                J.profile = g_prof_coarse;
                J.k = np_eff;
                J.out_simub = simub + (size_t)b0 * np_eff;
                J.set_only = g_fast_set_only;  // which lists to probe; their order only numbers the candidates (ties: tie_list_off below)
                flat_fast_plan(nb, J.emit, J.thresh, W.flat_plan);
                J.shared_units = true;
                J.out_rel = ppos + (size_t)b0 * np_eff;
                J.out_dist = W.tmp2.as<double>((size_t)nb * np_eff);
                J.out_ok = ok_c + b0;
                fast_topk(J);
            }
            if (blk < nqc) fast_quant_queries(qptr, qdtype, nqc, d);  // the list scan packs its units from the whole batch's digits
        } else {
            Prof pr(PROF_COARSE);
            wait_all();
            launch_row_norms(qptr, qdtype, nqc, d, qn);
            fast_quant_queries(qptr, qdtype, nqc, d);
            launch_widen_queries(qptr, qdtype, nqc * d, q64buf);
            bool cl2;
            int cepi;
            metric_to_epi(ix->metric == HB_IP ? HB_COSINE : ix->metric, true, cl2, cepi);
            int64_t *plan = g_ws.plan.as<int64_t>(6);
            double *coarse = g_ws.cand_val.as<double>((size_t)nqc * nlist);
            double *pval = g_ws.sel_val.as<double>((size_t)nqc * np_eff);
            int64_t h[6] = {0, nlist, 0, nqc, 0, ceil_div(nlist, kTileRows) * ceil_div(nqc, kTileQ)};
            HB_CUDA(cudaMemcpyAsync(plan, h, sizeof(h), cudaMemcpyHostToDevice, g_stream));
            sync_stream();
            ScanParams Sc;
            Sc.rows = ix->cents.p;
            Sc.row_norm = (const double *)ix->cent_norm.p;
            Sc.queries = qptr;
            Sc.q_norm = qn;
            Sc.d = d;
            Sc.nlist = 1;
            Sc.list_off = plan;
            Sc.lq_off = plan + 2;
            Sc.tile_prefix = plan + 4;
            Sc.out_stride = nlist;
            Sc.out = coarse;
            Sc.epi = cepi;
            launch_pairscan(Sc, HB_F64, qdtype, cl2);
            SelectParams L;
            L.vals = coarse;
            L.nseg = nqc;
            L.seg_stride = nlist;
            L.seg_len_const = nlist;
            L.k = np_eff;
            L.out_val = pval;
            L.out_pos = ppos;
            launch_select(L);
            {
                float one_bits;
                const int32_t one = 1;
                memcpy(&one_bits, &one, 4);
                launch_fill_f32((float *)ok_c, nqc, one_bits);  // int32 1 in every slot
            }
        }
        // plans: all (query, probe) pairs for EMIT, the nearest list of every query for THRESH
        const int64_t np = nqc * np_eff;
        int32_t *probes = g_ws.probes.as<int32_t>((size_t)np);
        int64_t *pair_out = g_ws.pair_out.as<int64_t>((size_t)np + 1);
        int32_t *qsel = g_ws.qsel.as<int32_t>((size_t)np);
        int64_t *lq_off = g_ws.lq_off.as<int64_t>(nlist + 1);
        int64_t *uprefix = W.uprefix.as<int64_t>(nlist + 1);
        int32_t *probes0 = W.probes0.as<int32_t>(nqc);
        int32_t *qsel0 = W.qsel0.as<int32_t>(nqc);
        int64_t *lq_off0 = W.lq_off0.as<int64_t>(nlist + 1);
        int64_t *uprefix0 = W.uprefix0.as<int64_t>(nlist + 1);
        // Probe pruning (needs the coarse stage's similarity bounds): sample the nearest list first, then drop every
        // (query, probed list) pair whose rows cannot reach the query's threshold, and plan the scan over the rest.
        const bool prune = coarse_tc && g_fast_prune && np_eff > 1;
        auto plan_emit = [&] {
            ivf_plan_fast(ppos, np_eff, nqc, np_eff, nlist, (const int64_t *)ix->list_off.p, probes, pair_out, qsel, lq_off, uprefix, kFastTile,
                          g_ws.tmp);
        };
        {
            Prof prp(PROF_PLAN);
            if (!prune) plan_emit();
            ivf_plan_fast(ppos, np_eff, nqc, 1, nlist, (const int64_t *)ix->list_off.p, probes0, nullptr, qsel0, lq_off0, uprefix0, kFastTile,
                          g_ws.tmp);
        }
        // no host round trip: unit counts stay on the device, the host sizes buffers and grids by their bounds
        int64_t *relk = W.relk.as<int64_t>((size_t)nqc * k);
        FastJob J;
        J.side = &S;
        J.list_off = (const int64_t *)ix->list_off.p;
        J.rows_exact = ix->rows.p;
        J.rdtype = ix->dtype;
        J.row_norm = (const double *)ix->norms.p;
        J.queries = qptr;
        J.qdtype = qdtype;
        J.q64 = q64;
        J.nq = nqc;
        J.qn = qn;
        J.d = d;
        J.metric = HB_COSINE;  // the list scan is always cosine (ivf_flat.clj:217-234)
        J.epi = EPI_COS;
        J.k = k;
        J.emit.nlist = nlist;
        J.emit.lq_off = lq_off;
        J.emit.unit_prefix = uprefix;
        J.emit.qsel = qsel;
        J.emit.pair_out = pair_out;
        J.emit.pair_div = np_eff;
        J.thresh.nlist = nlist;
        J.thresh.lq_off = lq_off0;
        J.thresh.unit_prefix = uprefix0;
        J.thresh.qsel = qsel0;
        J.thresh.pair_out = nullptr;
        J.thresh.pair_div = 1;
        set_units_bound(J.thresh, nqc, nlist, false);
        J.thresh.tile_limit = g_fast_sample_tiles;
        J.shared_units = false;
        J.out_rel = relk;
        J.out_dist = dist + (size_t)q0 * k;
        J.out_ok = ok_all + q0;
        J.cover_stats = true;
        J.profile = !g_prof_coarse;
        if (coarse_tc && g_fast_set_only && !shard_coarse) {  // probe order is approximate: cross-list distance ties go to the exact path
            J.tie_list_off = (const int64_t *)ix->list_off.p;
            J.tie_nlist = nlist;
        }
        if (prune) {
            J.phase = 1;
            fast_topk(J);  // query bounds + thresholds from the nearest list
            {
                Prof prp(PROF_PLAN);
                unsigned long long *npruned = (unsigned long long *)W.pruned.get(16);
                HB_CUDA(cudaMemsetAsync(npruned, 0, 16, g_stream));
                launch_prune_probes(ppos, simub, list_radius(ix), (const float *)W.thr.p, (const double *)W.qscale.p,
                                    (const double *)W.qeps.p, nqc, np_eff, (const int64_t *)ix->list_off.p, npruned);
                plan_emit();
                if (g_profile) {
                    launch_accumulate_u64(dev_stats() + DS_PRUNED_PAIRS, npruned, 2);
                    g_fast_probe_pairs += np;
                }
            }
            set_units_bound(J.emit, np, nlist, g_tc_interleave);
            J.phase = 2;
            fast_topk(J);
        } else {
            set_units_bound(J.emit, np, nlist, g_tc_interleave);
            fast_topk(J);
        }
        // what the main candidate pass covered: units, items (unit x row tile), distinct row tiles
        launch_ivf_resolve(relk, nqc, k, np_eff, probes, pair_out, (const int64_t *)ix->list_off.p, (const int64_t *)ix->list_rows.p,
                           ids + (size_t)q0 * k, ok_all + q0, ok_c);  // and ok &= the coarse stage's flags
    }
    fast_fallback(ok_all, queries, qdtype, nq, d, k, ids, dist, [&](const void *gq, int64_t nb, int64_t *gids, double *gdist) {
        ivf_search_exact(ix, gq, qdtype, nb, k, nprobe, gids, gdist, nullptr);
    });
}

}  // namespace hb

// =================================================================================================
// C ABI
// =================================================================================================

// search-knn over an uploaded HNSW graph (src/hnsw/ultra_fast.clj:346-374), batched: one warp per query
static void hnsw_search(hb_index *ix, const void *queries, int qdtype, int64_t nq, int k, int ef_param, int64_t *ids, double *dist) {
    if (nq == 0 || k == 0) return;
    if (ix->n == 0 || ix->entry < 0) {
        fill_empty_results(ids, dist, nq * k);
        return;
    }
    HB_REQUIRE(nq < (1ll << 31) && ix->n < (1ll << 31), "too many queries / rows for the HNSW search");
    const int ef = ef_param > 0 ? ef_param : std::max(k, 50);  // :355
    HB_REQUIRE(ef <= 4096, "ef > 4096 is not supported");
    const int d = ix->d;
    bool l2;
    int epi;
    metric_to_epi(ix->metric, true, l2, epi);
    const double *qn = nullptr;
    if (ix->metric == HB_COSINE) {
        double *x = g_ws.qnorm.as<double>(nq);
        launch_row_norms(queries, qdtype, nq, d, x);
        qn = x;
    }
    HnswSearchParams P;
    P.rows = ix->rows.p;
    P.row_norm = ix->metric == HB_COSINE ? (const double *)ix->norms.p : nullptr;
    P.d = d;
    P.max_level = ix->max_level;
    P.entry = ix->entry;
    P.adj_off = (const int64_t *const *)ix->adj_off_ptrs.p;
    P.adj_ids = (const int32_t *const *)ix->adj_ids_ptrs.p;
    P.queries = queries;
    P.q_norm = qn;
    P.ef = ef;
    P.k = k;
    P.epi = epi;
    P.ef_cap = ((ef + 1 + 3) / 4) * 4;
    int cc = g_hnsw_cand_cap > 0 ? g_hnsw_cand_cap : std::max(256, 4 * ef);
    const int W = hnsw_warps_per_cta();
    const size_t smem_max = 200 * 1024;
    while (cc > 16 && (size_t)W * hnsw_warp_smem(P.ef_cap, cc) > smem_max) cc /= 2;
    P.cand_cap_smem = cc;
    P.warp_smem = hnsw_warp_smem(P.ef_cap, cc);
    HB_REQUIRE((size_t)W * P.warp_smem <= 227 * 1024, "ef too large for the shared-memory queues");
    const int ctas_per_sm = std::max<int>(1, std::min<size_t>(4, (220 * 1024) / ((size_t)W * P.warp_smem + 1024)));
    const int grid = (int)std::min<int64_t>((int64_t)g_num_sms * ctas_per_sm, ceil_div(nq, W));
    // rows in flight: every resident warp gathers up to 32 rows at once; keep their prefetched lines within ~48 MB of the L2
    {
        const int64_t warps = (int64_t)grid * W;
        const int64_t lines = (48ll << 20) / (warps * 32 * 128);
        P.pf_window = g_hnsw_prefetch >= 0 ? g_hnsw_prefetch : (int)std::max<int64_t>(2, std::min<int64_t>(lines, 32));
    }
    const int64_t slots = (int64_t)g_num_sms * 4 * W;  // upper bound of grid * W, so the buffers are allocated once
    P.vwords = ceil_div(ix->n, 32);
    P.vcap = 8192;
    if (ix->hn_slots != slots) {
        uint32_t *v = ix->hn_visited.as<uint32_t>((size_t)slots * P.vwords);
        HB_CUDA(cudaMemsetAsync(v, 0, (size_t)slots * P.vwords * 4, g_stream));  // the kernel leaves it all zero
        ix->hn_vlist.as<int32_t>((size_t)slots * P.vcap);
        ix->hn_slots = slots;
    }
    P.visited = (uint32_t *)ix->hn_visited.p;
    P.vlist = (int32_t *)ix->hn_vlist.p;
    // ctrl: [0] work counter, [1] overflow count, [2..3] scored pairs (u64)
    int32_t *ctrl = ix->hn_ctrl.as<int32_t>(4);
    HB_CUDA(cudaMemsetAsync(ctrl, 0, 16, g_stream));
    int32_t *over = ix->hn_over.as<int32_t>(nq);
    P.next_work = ctrl;
    P.n_overflow = ctrl + 1;
    P.n_scored = (unsigned long long *)(ctrl + 2);
    P.overflow_list = over;
    P.nwork = (int)nq;
    P.out_ids = ids;
    P.out_dist = dist;
    {
        Prof pr(PROF_HNSW);
        launch_hnsw_search(P, ix->dtype, qdtype, l2, grid);
    }
    int32_t h[4];
    HB_CUDA(cudaMemcpyAsync(h, ctrl, 16, cudaMemcpyDeviceToHost, g_stream));
    sync_stream();
    if (h[1] > 0) {
        // the rare query whose candidate queue outgrew shared memory: same kernel, queue in global memory (a node is
        // offered at most once per layer, so n slots always suffice)
        const int nover = h[1];
        g_hnsw_overflows += nover;
        const int grid2 = (int)std::min<int64_t>(grid, ceil_div(nover, W));
        P.g_cand_cap = (int)std::min<int64_t>(ix->n + 1, INT32_MAX);
        P.g_cand_d = ix->hn_gcand_d.as<double>((size_t)grid2 * W * P.g_cand_cap);
        P.g_cand_i = ix->hn_gcand_i.as<int32_t>((size_t)grid2 * W * P.g_cand_cap);
        P.work_list = over;
        P.nwork = nover;
        HB_CUDA(cudaMemsetAsync(ctrl, 0, 8, g_stream));
        P.overflow_list = (int32_t *)g_ws.misc.as<int32_t>(nover);
        Prof pr(PROF_HNSW);
        launch_hnsw_search(P, ix->dtype, qdtype, l2, grid2);
        HB_CUDA(cudaMemcpyAsync(h, ctrl, 16, cudaMemcpyDeviceToHost, g_stream));
        sync_stream();
        if (h[1] > 0) throw Error(HB_ERR_CUDA, "HNSW candidate queue overflowed its global-memory capacity");
    }
    unsigned long long sc;
    memcpy(&sc, &h[2], 8);
    g_hnsw_scored += (int64_t)sc;
}

extern "C" {

HB_API int hb_init(int device) {
    return guarded([&] { ensure_init(device < 0 ? 0 : device); });
}
HB_API int hb_shutdown(void) {
    return guarded([&] {
        if (g_inited) cudaStreamSynchronize(g_stream);
        g_ws.release();
        g_fw.release();
        g_dev_stats.release();
        g_assign_side.release();
        g_kpp_buf.release();
    });
}
HB_API const char *hb_last_error(void) { return t_err.c_str(); }
HB_API int hb_version(void) { return 100; }
HB_API int hb_set_stream(void *cuda_stream) {
    return guarded([&] { g_stream = (cudaStream_t)cuda_stream; });
}
HB_API int hb_set_mode(int mode) {
    return guarded([&] {
        HB_REQUIRE(mode == HB_MODE_EXACT || mode == HB_MODE_FAST, "unknown mode");
        g_mode = mode;
    });
}
HB_API int hb_set_option(const char *name, int64_t value) {
    return guarded([&] {
        HB_REQUIRE(name, "null option name");
        if (!strcmp(name, "scratch_mb")) {
            HB_REQUIRE(value >= 1, "scratch_mb must be >= 1");
            g_scratch_budget = (size_t)value << 20;
        } else if (!strcmp(name, "tc_interleave")) {
            g_tc_interleave = value != 0;
        } else if (!strcmp(name, "fast_debug")) {
            g_fast_debug = value != 0;
        } else if (!strcmp(name, "fast_digits")) {
            HB_REQUIRE(value == 2 || value == 3, "fast_digits must be 2 or 3");
            g_fast_ns = (int)value;
        } else if (!strcmp(name, "fast_dense")) {
            g_fast_dense = value != 0;
        } else if (!strcmp(name, "prof_coarse")) {
            g_prof_coarse = value != 0;
        } else if (!strcmp(name, "fast_level_dense")) {
            HB_REQUIRE(value >= 0 && value <= 16, "fast_level_dense must be 0..16 row tiles");
            g_fast_level_dense = (int)value;
        } else if (!strcmp(name, "fast_level_ratio")) {
            HB_REQUIRE(value >= 2 && value <= 64, "fast_level_ratio must be 2..64");
            g_fast_level_ratio = (int)value;
        } else if (!strcmp(name, "fast_level_min")) {
            HB_REQUIRE(value >= 3, "fast_level_min must be >= 3");
            g_fast_level_min = (int)value;
        } else if (!strcmp(name, "fast_sample_tiles")) {
            HB_REQUIRE(value >= 1 && value <= 64, "fast_sample_tiles must be 1..64");
            g_fast_sample_tiles = (int)value;
        } else if (!strcmp(name, "cuda_profiler")) {
            // brackets the region `ncu --profile-from-start off` captures (bench.py: the timed steps only)
            ensure_init();
            HB_CUDA(cudaDeviceSynchronize());
            if (value) HB_CUDA(cudaProfilerStart()); else HB_CUDA(cudaProfilerStop());
        } else if (!strcmp(name, "profile")) {
            prof_collect();
            g_profile = value != 0;
            for (int i = 0; i < PROF_NTAGS; ++i) g_prof_ms[i] = 0, g_prof_n[i] = 0;
            g_fast_queries = g_fast_fallbacks = g_coarse_fallbacks = 0;
            g_kpp_scored = g_kpp_walked = g_kpp_steps = 0;
            g_fast_probe_pairs = 0;
            HB_CUDA(cudaMemsetAsync(dev_stats(), 0, DS_COUNT * 8, g_stream));
            g_hnsw_scored = g_hnsw_overflows = 0;
        } else if (!strcmp(name, "hnsw_prefetch")) {
            HB_REQUIRE(value >= -1 && value <= 64, "hnsw_prefetch must be -1..64");
            g_hnsw_prefetch = (int)value;
        } else if (!strcmp(name, "tc_half_m")) {
            g_tc_half_m = value != 0;
        } else if (!strcmp(name, "fast_carry")) {
            g_fast_carry = value != 0;
        } else if (!strcmp(name, "kpp_scale")) {
            g_kpp_scale = value != 0;
        } else if (!strcmp(name, "tc_narrow")) {
            g_tc_narrow = value != 0;
        } else if (!strcmp(name, "rowstream")) {
            g_use_rowstream = value != 0;
        } else if (!strncmp(name, "stream_", 7)) {
            set_rowstream_option(name, (int)value);
        } else if (!strcmp(name, "fast_prune")) {
            g_fast_prune = value != 0;
        } else if (!strcmp(name, "fast_set_only")) {
            g_fast_set_only = value != 0;
        } else if (!strcmp(name, "host_feed")) {
            HB_REQUIRE(value == 0 || (value >= 128 && value % 128 == 0), "host_feed must be 0 or a multiple of 128");
            g_host_feed_block = value;
        } else if (!strcmp(name, "comm_p2p")) {
            HB_REQUIRE(value >= 0 && value <= 2, "comm_p2p must be 0, 1 or 2");
            g_comm_p2p = (int)value;
        } else if (!strcmp(name, "micro_batch")) {
            HB_REQUIRE(value >= 0 && value <= 4096, "micro_batch must be 0..4096");
            g_micro_batch = (int)value;
        } else if (!strcmp(name, "hnsw_cand_cap")) {
            HB_REQUIRE(value >= 0 && value <= 8192, "hnsw_cand_cap must be 0..8192");
            g_hnsw_cand_cap = (int)value;
        } else {
            throw Error(HB_ERR_INVALID, std::string("unknown option ") + name);
        }
    });
}
HB_API int hb_get_stat(const char *name, double *out) {
    return guarded([&] {
        HB_REQUIRE(name && out, "null argument");
        static const char *ms_names[PROF_NTAGS] = {"scan_ms", "coarse_ms", "select_ms", "plan_ms", "assign_ms", "tc_ms", "pack_ms", "rescore_ms", "tc_sample_ms", "hnsw_ms",
                                                   "exchange_ms", "merge_ms", "allreduce_ms", "update_ms"};
        static const char *n_names[PROF_NTAGS] = {"scan_count", "coarse_count", "select_count", "plan_count", "assign_count", "tc_count", "pack_count", "rescore_count", "tc_sample_count", "hnsw_count",
                                                  "exchange_count", "merge_count", "allreduce_count", "update_count"};
        prof_collect();
        for (int i = 0; i < PROF_NTAGS; ++i) {
            if (!strcmp(name, ms_names[i])) { *out = g_prof_ms[i]; return; }
            if (!strcmp(name, n_names[i])) { *out = (double)g_prof_n[i]; return; }
        }
        if (!strcmp(name, "fast_queries")) { *out = (double)g_fast_queries; return; }
        if (!strcmp(name, "fast_fallbacks")) { *out = (double)g_fast_fallbacks; return; }
        if (!strcmp(name, "coarse_fallbacks")) { *out = (double)g_coarse_fallbacks; return; }
        if (!strcmp(name, "kpp_rows_scored")) { *out = (double)g_kpp_scored; return; }
        if (!strcmp(name, "kpp_chunks_walked")) { *out = (double)g_kpp_walked; return; }
        if (!strcmp(name, "kpp_steps")) { *out = (double)g_kpp_steps; return; }
        if (!strcmp(name, "fast_pruned_pairs")) { *out = dev_stat(DS_PRUNED_PAIRS); return; }
        if (!strcmp(name, "tc_units")) { *out = dev_stat(DS_TC_UNITS); return; }
        if (!strcmp(name, "tc_items")) { *out = dev_stat(DS_TC_ITEMS); return; }
        if (!strcmp(name, "tc_tiles")) { *out = dev_stat(DS_TC_TILES); return; }
        if (!strcmp(name, "tc_half_units")) { *out = dev_stat(DS_TC_HALF_UNITS); return; }
        if (!strcmp(name, "tc_narrow_units")) { *out = dev_stat(DS_TC_NARROW_UNITS); return; }
        if (!strcmp(name, "tc_narrow_items")) { *out = dev_stat(DS_TC_NARROW_ITEMS); return; }
        if (!strcmp(name, "tc_narrow_slots")) { *out = dev_stat(DS_TC_NARROW_SLOTS); return; }  // query slots the narrow units read (groups of 8)
        if (!strcmp(name, "fast_pruned_rows")) { *out = dev_stat(DS_PRUNED_ROWS); return; }
        if (!strcmp(name, "fast_probe_pairs")) { *out = (double)g_fast_probe_pairs; return; }
        if (!strcmp(name, "hnsw_scored")) { *out = (double)g_hnsw_scored; return; }
        if (!strcmp(name, "hnsw_overflows")) { *out = (double)g_hnsw_overflows; return; }
        if (!strcmp(name, "micro_batches")) { *out = (double)g_mb_batches; return; }
        if (!strcmp(name, "micro_batch_requests")) { *out = (double)g_mb_requests; return; }
        if (!strncmp(name, "mma_clocks_", 11)) {
            ensure_init();
            const char *k = name + 11;
            const int kind = !strcmp(k, "i8") ? 0 : !strcmp(k, "bf16") ? 1 : !strcmp(k, "e4m3") ? 2 : !strcmp(k, "tf32") ? 3 : -1;
            HB_REQUIRE(kind >= 0, "mma_clocks_{i8,bf16,e4m3,tf32}");
            *out = mma_clocks_per_instr(kind);
            return;
        }
        if (!strcmp(name, "fp64_peak_tflops")) {
            ensure_init();
            *out = fp64_peak_tflops();
            return;
        }
        throw Error(HB_ERR_INVALID, std::string("unknown stat ") + name);
    });
}
HB_API int64_t hb_launch_count(int reset) {
    std::lock_guard<std::mutex> lk(g_mu);
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

HB_API int hb_row_norms(const void *rows, int64_t n, int32_t d, int dtype, double *out_norms) {
    return guarded([&] {
        ensure_init();
        check_dtype(dtype);
        HB_REQUIRE(n >= 0 && d >= 1 && (n == 0 || (rows && out_norms)), "bad arguments");
        if (n == 0) return;
        const void *r = stage_in(rows, (size_t)n * d * dtype_size(dtype), g_ws.in_a);
        OutStage o = stage_out(out_norms, (size_t)n * 8, g_ws.out_a);
        launch_row_norms(r, dtype, n, d, (double *)o.dev);
        finish_out(o);
        sync_stream();
    });
}

HB_API int hb_pairwise(const void *a, int64_t na, int adtype, const void *b, int64_t nb, int bdtype, int32_t d,
                       int metric, double *out) {
    return guarded([&] {
        ensure_init();
        check_dtype(adtype);
        check_dtype(bdtype);
        HB_REQUIRE(adtype != HB_BF16, "`a` must be fp32 or fp64");
        HB_REQUIRE(na >= 0 && nb >= 0 && d >= 1, "bad arguments");
        if (na == 0 || nb == 0) return;
        HB_REQUIRE(a && b && out, "null buffer");
        const void *qa = stage_in(a, (size_t)na * d * dtype_size(adtype), g_ws.in_a);
        const void *rb = stage_in(b, (size_t)nb * d * dtype_size(bdtype), g_ws.in_b);
        OutStage o = stage_out(out, (size_t)na * nb * 8, g_ws.out_a);
        bool l2;
        int epi;
        metric_to_epi(metric, true, l2, epi);
        if (metric == HB_IP) epi = EPI_DOT;
        const double *qn = nullptr, *rn = nullptr;
        if (metric == HB_COSINE) {
            double *x = g_ws.qnorm.as<double>(na);
            double *y = g_ws.misc.as<double>(nb);
            launch_row_norms(qa, adtype, na, d, x);
            launch_row_norms(rb, bdtype, nb, d, y);
            qn = x;
            rn = y;
        }
        int64_t *lo, *lq, *tp;
        dense_plan(nb, na, lo, lq, tp, g_ws.plan);
        ScanParams S;
        S.rows = rb;
        S.row_norm = rn;
        S.queries = qa;
        S.q_norm = qn;
        S.d = d;
        S.nlist = 1;
        S.list_off = lo;
        S.lq_off = lq;
        S.tile_prefix = tp;
        S.out_stride = nb;
        S.out = (double *)o.dev;
        S.epi = epi;
        launch_pairscan(S, bdtype, adtype, l2);
        finish_out(o);
        sync_stream();
    });
}

HB_API int hb_flat_create(const void *rows, int64_t n, int32_t d, int dtype, int metric, hb_index **out) {
    return guarded([&] {
        ensure_init();
        check_dtype(dtype);
        HB_REQUIRE(out, "null out");
        HB_REQUIRE(n >= 0 && d >= 1 && (n == 0 || rows), "bad arguments");
        HB_REQUIRE(metric == HB_COSINE || metric == HB_L2 || metric == HB_IP, "unknown metric");
        hb_index *ix = new hb_index();
        try {
            ix->type = HB_INDEX_FLAT;
            ix->dtype = dtype;
            ix->metric = metric;
            ix->d = d;
            ix->n = n;
            const size_t bytes = (size_t)n * d * dtype_size(dtype);
            void *r = ix->rows.get(std::max<size_t>(bytes, 16));
            if (bytes)
                HB_CUDA(cudaMemcpyAsync(r, rows, bytes, is_device_ptr(rows) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, g_stream));
            double *nm = ix->norms.as<double>(std::max<int64_t>(n, 1));
            launch_row_norms(r, dtype, n, d, nm);
            sync_stream();
        } catch (...) {
            ix->release();
            delete ix;
            throw;
        }
        *out = ix;
    });
}

static void ivf_common_checks(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist) {
    check_dtype(dtype);
    HB_REQUIRE(n >= 1 && d >= 1 && rows, "build needs at least one row");
    HB_REQUIRE(nlist >= 1, "num-partitions must be >= 1");
    HB_REQUIRE(metric == HB_COSINE || metric == HB_L2, "IVF distance-fn must be cosine or euclidean");
}

HB_API int hb_ivf_build(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist, int32_t iters,
                        int64_t seed, hb_index **out) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(out, "null out");
        ivf_common_checks(rows, n, d, dtype, metric, nlist);
        HB_REQUIRE(iters >= 0, "max-iterations must be >= 0");
        hb_index *ix = new hb_index();
        try {
            ix->type = HB_INDEX_IVF_FLAT;
            ix->dtype = dtype;
            ix->metric = metric;
            ix->d = d;
            ix->n = n;
            ix->nlist = nlist;
            const void *r = stage_in(rows, (size_t)n * d * dtype_size(dtype), g_ws.in_a);
            double *cents = ix->cents.as<double>((size_t)nlist * d);
            int32_t *assign = ix->assign.as<int32_t>(n);
            kmeans_exact(r, dtype, n, d, metric, nlist, iters, seed, nullptr, cents, assign, nullptr);
            ivf_finalize(ix, r, (const double *)g_ws.misc.p);
        } catch (...) {
            ix->release();
            delete ix;
            throw;
        }
        *out = ix;
    });
}

// build-lightning-index with :smart-partition? true (src/hnsw/ann/partition/lightning.clj:84-130)
HB_API int hb_lightning_build(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist, int64_t seed,
                              hb_index **out) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(out, "null out");
        ivf_common_checks(rows, n, d, dtype, metric, nlist);
        HB_REQUIRE(n < (1ll << 31), "n must be < 2^31");
        hb_index *ix = new hb_index();
        try {
            ix->type = HB_INDEX_IVF_FLAT;
            ix->dtype = dtype;
            ix->metric = metric;
            ix->d = d;
            ix->n = n;
            ix->nlist = nlist;
            const bool l2 = metric == HB_L2;
            const void *r = stage_in(rows, (size_t)n * d * dtype_size(dtype), g_ws.in_a);
            double *norm = g_ws.misc.as<double>(n);
            launch_row_norms(r, dtype, n, d, norm);
            int64_t *seeds = g_ws.misc2.as<int64_t>(nlist);
            kpp_seeds(r, dtype, n, d, norm, l2, nlist, seed, true, seeds);  // :86-109
            double *cents = ix->cents.as<double>((size_t)nlist * d);
            int32_t *assign = ix->assign.as<int32_t>(n);
            launch_init_centroids(r, dtype, d, seeds, nlist, cents);
            double *cnorm = g_ws.misc3.as<double>((size_t)nlist + 2 + 2 * (size_t)n);
            launch_row_norms(cents, HB_F64, nlist, d, cnorm);
            assign_rows(r, dtype, norm, n, d, cents, cnorm, nlist, l2, assign);  // assign-to-partition over the seeds, :111-115
            int64_t *list_off = g_ws.lq_off.as<int64_t>(nlist + 1);
            int64_t *list_rows = g_ws.cand_id.as<int64_t>(n);
            build_lists(assign, n, nlist, list_off, list_rows, g_ws.tmp);
            // routing centroids = partition means, an empty partition gets the zero vector (:122-126)
            HB_CUDA(cudaMemsetAsync(cents, 0, (size_t)nlist * d * 8, g_stream));
            launch_update_centroids(r, dtype, d, list_off, list_rows, nlist, cents, nullptr, nullptr);
            ivf_finalize(ix, r, norm);
        } catch (...) {
            ix->release();
            delete ix;
            throw;
        }
        *out = ix;
    });
}

HB_API int hb_ivf_import(const void *rows, int64_t n, int32_t d, int dtype, int metric, const double *centroids,
                         int32_t nlist, const int32_t *assignments, hb_index **out) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(out && centroids && assignments, "null buffer");
        ivf_common_checks(rows, n, d, dtype, metric, nlist);
        hb_index *ix = new hb_index();
        try {
            ix->type = HB_INDEX_IVF_FLAT;
            ix->dtype = dtype;
            ix->metric = metric;
            ix->d = d;
            ix->n = n;
            ix->nlist = nlist;
            const void *r = stage_in(rows, (size_t)n * d * dtype_size(dtype), g_ws.in_a);
            double *cents = ix->cents.as<double>((size_t)nlist * d);
            int32_t *assign = ix->assign.as<int32_t>(n);
            HB_CUDA(cudaMemcpyAsync(cents, centroids, (size_t)nlist * d * 8, cudaMemcpyDefault, g_stream));
            HB_CUDA(cudaMemcpyAsync(assign, assignments, (size_t)n * 4, cudaMemcpyDefault, g_stream));
            {
                Validator v;
                v.range_i32(assign, n, 0, nlist);
                v.require("hb_ivf_import: assignment out of range");
            }
            double *norm = g_ws.misc.as<double>(n);
            launch_row_norms(r, dtype, n, d, norm);
            ivf_finalize(ix, r, norm);
        } catch (...) {
            ix->release();
            delete ix;
            throw;
        }
        *out = ix;
    });
}

HB_API int hb_ivf_export(const hb_index *index, double *out_centroids, int32_t *out_assignments) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(index && index->type == HB_INDEX_IVF_FLAT, "not an IVF-FLAT index");
        if (out_centroids)
            HB_CUDA(cudaMemcpyAsync(out_centroids, index->cents.p, (size_t)index->nlist * index->d * 8, cudaMemcpyDefault, g_stream));
        if (out_assignments)
            HB_CUDA(cudaMemcpyAsync(out_assignments, index->assign.p, (size_t)index->n * 4, cudaMemcpyDefault, g_stream));
        sync_stream();
    });
}

// The search behind hb_search / hb_sharded_search: stages the queries (host batches of an IVF FAST search optionally in
// blocks on a copy stream), dispatches on the index type and the effective mode, results into device buffers [nq x k].
static int effective_mode(const hb_index *index) { return index->mode >= 0 ? index->mode : g_mode; }
static void check_search_args(const hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t k, const void *out_ids,
                              const void *out_dist) {
    HB_REQUIRE(index, "null index");
    HB_REQUIRE(qdtype == HB_F32 || qdtype == HB_F64, "queries must be fp32 or fp64");
    HB_REQUIRE(nq >= 0 && k >= 0, "bad arguments");
    if (nq == 0 || k == 0) return;
    HB_REQUIRE(queries && out_ids && out_dist, "null buffer");
    HB_REQUIRE(k <= 1024, "k > 1024 is not supported");
}
static void search_core(hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t k, int32_t param, int64_t *ids_dev,
                        double *dist_dev) {
    const int mode = effective_mode(index);
    HostFeed feed;
    const void *q = nullptr;
    const size_t qrow = (size_t)index->d * dtype_size(qdtype);
    if (index->type == HB_INDEX_IVF_FLAT && mode == HB_MODE_FAST && !is_device_ptr(queries) && g_host_feed_block > 0 &&
        nq > g_host_feed_block) {
        // blocks of 2048 queries on a copy stream, one event each (pinned host memory makes the copies asynchronous)
        if (!g_copy_stream) HB_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        char *dq = (char *)g_ws.in_b.get((size_t)nq * qrow);
        feed.block = g_host_feed_block;
        for (int64_t b0 = 0; b0 < nq; b0 += feed.block) {
            const int64_t nb = std::min<int64_t>(feed.block, nq - b0);
            HB_CUDA(cudaMemcpyAsync(dq + (size_t)b0 * qrow, (const char *)queries + (size_t)b0 * qrow, (size_t)nb * qrow,
                                    cudaMemcpyHostToDevice, g_copy_stream));
            cudaEvent_t e;
            HB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            HB_CUDA(cudaEventRecord(e, g_copy_stream));
            feed.ready.push_back(e);
        }
        q = dq;
    } else {
        q = stage_in(queries, (size_t)nq * qrow, g_ws.in_b);
    }
    struct FeedGuard {
        HostFeed &f;
        ~FeedGuard() {
            for (cudaEvent_t e : f.ready) cudaEventDestroy(e);
        }
    } feed_guard{feed};
    const int saved_mode = g_mode;
    struct ModeGuard {  // the orchestration below reads g_mode (assign_rows etc.): make the index's mode the current one
        int saved;
        ~ModeGuard() { g_mode = saved; }
    } mode_guard{saved_mode};
    g_mode = mode;
    if (index->type == HB_INDEX_FLAT) {
        if (mode == HB_MODE_FAST) flat_search_fast(index, q, qdtype, nq, k, ids_dev, dist_dev);
        else
            flat_search_exact(index->rows.p, index->dtype, (const double *)index->norms.p, index->n, index->d, index->metric, q,
                              qdtype, nq, k, ids_dev, dist_dev);
    } else if (index->type == HB_INDEX_IVF_FLAT) {
        HB_REQUIRE(param >= 1, "num-probes must be >= 1");
        if (mode == HB_MODE_FAST) ivf_search_fast(index, q, qdtype, nq, k, param, ids_dev, dist_dev, feed.block ? &feed : nullptr);
        else ivf_search_exact(index, q, qdtype, nq, k, param, ids_dev, dist_dev, nullptr);
    } else if (index->type == HB_INDEX_HNSW) {
        HB_REQUIRE(param >= 0, "ef must be >= 0");
        hnsw_search(index, q, qdtype, nq, k, param, ids_dev, dist_dev);
    } else {
        throw Error(HB_ERR_UNSUPPORTED, "search on this index type is not implemented");
    }
}

static void search_call(hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t k, int32_t param, int64_t *out_ids,
                        double *out_dist) {
    ensure_init();
    check_search_args(index, queries, qdtype, nq, k, out_ids, out_dist);
    if (nq == 0 || k == 0) return;
    OutStage oi = stage_out(out_ids, (size_t)nq * k * 8, g_ws.out_a);
    OutStage od = stage_out(out_dist, (size_t)nq * k * 8, g_ws.out_b);
    search_core(index, queries, qdtype, nq, k, param, (int64_t *)oi.dev, (double *)od.dev);
    finish_out(oi);
    finish_out(od);
    sync_stream();
}

// ---- micro-batching of concurrent small calls (src/hnsw/helper/parallel_search.clj:15-49) ---------------------------------
// The reference fans single-query searches out over a thread pool (up to 50 threads on one index,
// src/hnsw/wip/31k-multithread-sb.clj:127-134).  On the device one query costs as much as eight, so concurrent small
// calls on HOST buffers are combined: while a search is in flight the callers that arrive queue up; the next leader takes
// every queued request that matches its index / k / param / dtype and answers them with ONE batched search (group commit,
// no timer: an idle library answers a lone caller at once).  Every path returns the same bits whatever the batch is.
struct PendingSearch {
    hb_index *index;
    const void *queries;
    int qdtype;
    int64_t nq;
    int32_t k, param;
    int64_t *out_ids;
    double *out_dist;
    int status = HB_OK;
    std::string err;
    bool done = false;
};
static std::mutex q_mu;
static std::condition_variable q_cv;
static std::vector<PendingSearch *> q_pending;
static bool q_leader = false;
static std::vector<char> g_mb_q;  // host staging of a combined batch
static std::vector<int64_t> g_mb_ids;
static std::vector<double> g_mb_dist;

static void run_batch(std::vector<PendingSearch *> &batch) {
    PendingSearch &f = *batch[0];
    int status = HB_OK;
    std::string err;
    if (batch.size() == 1) {
        status = guarded([&] { search_call(f.index, f.queries, f.qdtype, f.nq, f.k, f.param, f.out_ids, f.out_dist); });
        if (status != HB_OK) err = t_err;
    } else {
        status = guarded([&] {
            ensure_init();
            const size_t qrow = (size_t)f.index->d * dtype_size(f.qdtype);
            int64_t total = 0;
            for (PendingSearch *r : batch) total += r->nq;
            g_mb_q.resize((size_t)total * qrow);
            g_mb_ids.resize((size_t)total * f.k);
            g_mb_dist.resize((size_t)total * f.k);
            int64_t o = 0;
            for (PendingSearch *r : batch) {
                memcpy(g_mb_q.data() + (size_t)o * qrow, r->queries, (size_t)r->nq * qrow);
                o += r->nq;
            }
            search_call(f.index, g_mb_q.data(), f.qdtype, total, f.k, f.param, g_mb_ids.data(), g_mb_dist.data());
            o = 0;
            for (PendingSearch *r : batch) {
                memcpy(r->out_ids, g_mb_ids.data() + (size_t)o * f.k, (size_t)r->nq * f.k * 8);
                memcpy(r->out_dist, g_mb_dist.data() + (size_t)o * f.k, (size_t)r->nq * f.k * 8);
                o += r->nq;
            }
            g_mb_batches += 1;
            g_mb_requests += (int64_t)batch.size();
        });
        if (status != HB_OK) err = t_err;
    }
    for (PendingSearch *r : batch) {
        r->status = status;
        r->err = err;
    }
}

static int search_combined(PendingSearch &me) {
    std::unique_lock<std::mutex> lk(q_mu);
    q_pending.push_back(&me);
    while (!me.done) {
        if (q_leader) {
            q_cv.wait(lk);
            continue;
        }
        q_leader = true;
        // everything queued that can share a launch with the oldest request
        std::vector<PendingSearch *> batch, rest;
        PendingSearch &f = *q_pending.front();
        int64_t total = 0;
        for (PendingSearch *r : q_pending) {
            const bool same = r->index == f.index && r->k == f.k && r->param == f.param && r->qdtype == f.qdtype;
            if (same && (batch.empty() || total + r->nq <= g_micro_batch)) {
                batch.push_back(r);
                total += r->nq;
            } else {
                rest.push_back(r);
            }
        }
        q_pending.swap(rest);
        lk.unlock();
        run_batch(batch);
        lk.lock();
        for (PendingSearch *r : batch) r->done = true;
        q_leader = false;
        q_cv.notify_all();
    }
    if (me.status != HB_OK) t_err = me.err;
    return me.status;
}

HB_API int hb_search(hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t k, int32_t param,
                     int64_t *out_ids, double *out_dist) {
    // small calls on host buffers go through the combiner; everything else straight to the device
    if (g_micro_batch > 0 && index && queries && out_ids && out_dist && nq >= 1 && nq <= kSmallScanQ && k >= 1 && k <= 1024 &&
        (qdtype == HB_F32 || qdtype == HB_F64)) {
        bool host = false;
        const int st = guarded([&] {
            ensure_init();
            host = !is_device_ptr(queries) && !is_device_ptr(out_ids) && !is_device_ptr(out_dist);
        });
        if (st != HB_OK) return st;
        if (host) {
            PendingSearch me{index, queries, qdtype, nq, k, param, out_ids, out_dist};
            return search_combined(me);
        }
    }
    return guarded([&] { search_call(index, queries, qdtype, nq, k, param, out_ids, out_dist); });
}

// ---- multi-GPU: one process per GPU (hb_comm.cu) -------------------------------------------------------------------------

HB_API int hb_comm_unique_id(void *out_id) {
    return guarded([&] {
        HB_REQUIRE(out_id, "null buffer");
        comm_unique_id(out_id);
    });
}
HB_API int hb_comm_init(const void *id, int32_t nranks, int32_t rank) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(id, "null id");
        comm_init(id, nranks, rank, g_device);
    });
}
HB_API int hb_comm_info(int32_t *out_nranks, int32_t *out_rank, int32_t *out_p2p) {
    return guarded([&] {
        const CommInfo &c = comm_info();
        if (out_nranks) *out_nranks = c.inited ? c.nranks : 0;
        if (out_rank) *out_rank = c.inited ? c.rank : 0;
        if (out_p2p) *out_p2p = c.p2p ? 1 : 0;
    });
}
HB_API int hb_comm_shutdown(void) {
    return guarded([&] { comm_shutdown(); });
}
HB_API int hb_comm_broadcast(void *buf, int64_t bytes, int32_t root) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(bytes >= 0 && (bytes == 0 || buf), "bad arguments");
        if (bytes == 0) return;
        if (is_device_ptr(buf)) {
            comm_broadcast_bytes(buf, bytes, root);
        } else {
            void *d = g_ws.in_a.get((size_t)bytes);
            if (comm_info().rank == root) HB_CUDA(cudaMemcpyAsync(d, buf, (size_t)bytes, cudaMemcpyHostToDevice, g_stream));
            comm_broadcast_bytes(d, bytes, root);
            HB_CUDA(cudaMemcpyAsync(buf, d, (size_t)bytes, cudaMemcpyDeviceToHost, g_stream));
        }
        sync_stream();
    });
}
HB_API int hb_comm_allreduce_f64(double *buf, int64_t count, int32_t op) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(count >= 0 && (count == 0 || buf) && (op == 0 || op == 1), "bad arguments (op: 0 = sum, 1 = max)");
        if (count == 0) return;
        const bool dev = is_device_ptr(buf);
        double *d = dev ? buf : g_ws.in_a.as<double>((size_t)count);
        if (!dev) HB_CUDA(cudaMemcpyAsync(d, buf, (size_t)count * 8, cudaMemcpyHostToDevice, g_stream));
        if (op == 0) comm_allreduce_sum_f64(d, count);
        else comm_allreduce_max_f64(d, count);
        if (!dev) HB_CUDA(cudaMemcpyAsync(buf, d, (size_t)count * 8, cudaMemcpyDeviceToHost, g_stream));
        sync_stream();
    });
}

HB_API int hb_index_set_id_base(hb_index *index, int64_t first_global_row) {
    return guarded([&] {
        HB_REQUIRE(index && first_global_row >= 0, "bad arguments");
        index->id_base = first_global_row;
    });
}
HB_API int hb_index_set_coarse_sharded(hb_index *index, int on) {
    return guarded([&] {
        HB_REQUIRE(index && index->type == HB_INDEX_IVF_FLAT, "not an IVF-FLAT index");
        index->coarse_sharded = on != 0;
    });
}
HB_API int hb_index_set_mode(hb_index *index, int mode) {
    return guarded([&] {
        HB_REQUIRE(index && (mode == -1 || mode == HB_MODE_EXACT || mode == HB_MODE_FAST), "bad arguments");
        index->mode = mode;
    });
}

// Global top-k over the row shards of all ranks: local search (any index type) -> exchange of the local top-k ->
// merge, all on the library's stream.  partitioned_hnsw.clj:149-196 (search every partition, concat, sort-by :distance, take k).
HB_API int hb_sharded_search(hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t k, int32_t param,
                             int64_t *out_ids, double *out_dist) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(comm_info().inited, "hb_sharded_search before hb_comm_init");
        check_search_args(index, queries, qdtype, nq, k, out_ids, out_dist);
        if (nq == 0 || k == 0) return;
        OutStage oi = stage_out(out_ids, (size_t)nq * k * 8, g_ws.out_a);
        OutStage od = stage_out(out_dist, (size_t)nq * k * 8, g_ws.out_b);
        int64_t *lid = g_ws.sh_ids.as<int64_t>((size_t)nq * k);
        double *ldist = g_ws.sh_dist.as<double>((size_t)nq * k);
        search_core(index, queries, qdtype, nq, k, param, lid, ldist);
        ExchangeStats st;
        if (g_profile) {
            cudaEventCreate(&st.e0);
            cudaEventCreate(&st.e1);
            cudaEventCreate(&st.e2);
        }
        comm_topk_exchange_merge(ldist, lid, index->id_base, nq, k, (double *)od.dev, (int64_t *)oi.dev, g_comm_p2p, &st);
        finish_out(oi);
        finish_out(od);
        sync_stream();
        if (g_profile) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, st.e0, st.e1);
            cudaEventElapsedTime(&b, st.e1, st.e2);
            g_prof_ms[PROF_EXCHANGE] += a, g_prof_n[PROF_EXCHANGE] += 1;
            g_prof_ms[PROF_MERGE] += b, g_prof_n[PROF_MERGE] += 1;
            cudaEventDestroy(st.e0);
            cudaEventDestroy(st.e1);
            cudaEventDestroy(st.e2);
        }
        comm_check_exchange();
    });
}

HB_API int hb_ivf_probes(hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t nprobe,
                         int32_t *out_probes) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(index && index->type == HB_INDEX_IVF_FLAT, "not an IVF-FLAT index");
        HB_REQUIRE(qdtype == HB_F32 || qdtype == HB_F64, "queries must be fp32 or fp64");
        HB_REQUIRE(nq >= 0 && nprobe >= 1, "bad arguments");
        if (nq == 0) return;
        HB_REQUIRE(queries && out_probes, "null buffer");
        const void *q = stage_in(queries, (size_t)nq * index->d * dtype_size(qdtype), g_ws.in_b);
        OutStage op = stage_out(out_probes, (size_t)nq * nprobe * 4, g_ws.out_c);
        ivf_search_exact(index, q, qdtype, nq, 0, nprobe, nullptr, nullptr, (int32_t *)op.dev);
        finish_out(op);
        sync_stream();
    });
}

HB_API int hb_kmeanspp_init(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist, int64_t seed,
                            int64_t *out_seed_rows) {
    return guarded([&] {
        ensure_init();
        ivf_common_checks(rows, n, d, dtype, metric, nlist);
        HB_REQUIRE(out_seed_rows, "null buffer");
        const void *r = stage_in(rows, (size_t)n * d * dtype_size(dtype), g_ws.in_a);
        OutStage o = stage_out(out_seed_rows, (size_t)nlist * 8, g_ws.out_a);
        double *cents = g_ws.cand_val.as<double>((size_t)nlist * d);
        int32_t *assign = g_ws.probes.as<int32_t>(n);
        // iters = -1: seeds only (kmeans_exact's loop body runs `it <= iters`)
        kmeans_exact(r, dtype, n, d, metric, nlist, -1, seed, nullptr, cents, assign, (int64_t *)o.dev);
        finish_out(o);
        sync_stream();
    });
}

HB_API int hb_kmeans_assign(const void *rows, int64_t n, int32_t d, int dtype, int metric, const double *centroids,
                            int32_t nlist, int32_t *out_assign) {
    return guarded([&] {
        ensure_init();
        ivf_common_checks(rows, n, d, dtype, metric, nlist);
        HB_REQUIRE(centroids && out_assign, "null buffer");
        const void *r = stage_in(rows, (size_t)n * d * dtype_size(dtype), g_ws.in_a);
        const double *c = (const double *)stage_in(centroids, (size_t)nlist * d * 8, g_ws.in_c);
        OutStage o = stage_out(out_assign, (size_t)n * 4, g_ws.out_a);
        double *norm = g_ws.misc.as<double>(n);
        double *cnorm = g_ws.qnorm.as<double>(nlist);
        launch_row_norms(r, dtype, n, d, norm);
        launch_row_norms(c, HB_F64, nlist, d, cnorm);
        {
            Prof pr(PROF_ASSIGN);
            assign_rows(r, dtype, norm, n, d, c, cnorm, nlist, metric == HB_L2, (int32_t *)o.dev);
        }
        finish_out(o);
        sync_stream();
    });
}

HB_API int hb_kmeans_update(const void *rows, int64_t n, int32_t d, int dtype, const int32_t *assign, int32_t nlist,
                            double *centroids, double *out_sums, int64_t *out_counts) {
    return guarded([&] {
        ensure_init();
        ivf_common_checks(rows, n, d, dtype, HB_COSINE, nlist);
        HB_REQUIRE(assign && (centroids || out_sums), "null buffer");
        const void *r = stage_in(rows, (size_t)n * d * dtype_size(dtype), g_ws.in_a);
        const int32_t *a = (const int32_t *)stage_in(assign, (size_t)n * 4, g_ws.in_c);
        {
            Validator v;
            v.range_i32(a, n, 0, nlist);
            v.require("hb_kmeans_update: assignment out of range");
        }
        int64_t *list_off = g_ws.lq_off.as<int64_t>(nlist + 1);
        int64_t *list_rows = g_ws.cand_id.as<int64_t>(n);
        build_lists(a, n, nlist, list_off, list_rows, g_ws.tmp);
        if (out_sums) {
            OutStage os = stage_out(out_sums, (size_t)nlist * d * 8, g_ws.out_a);
            OutStage oc = stage_out(out_counts, (size_t)nlist * 8, g_ws.out_b);
            launch_update_centroids(r, dtype, d, list_off, list_rows, nlist, nullptr, (double *)os.dev, (int64_t *)oc.dev);
            finish_out(os);
            finish_out(oc);
        } else {
            const bool host = !is_device_ptr(centroids);
            double *c = centroids;
            if (host) {
                c = g_ws.out_a.as<double>((size_t)nlist * d);
                HB_CUDA(cudaMemcpyAsync(c, centroids, (size_t)nlist * d * 8, cudaMemcpyHostToDevice, g_stream));
            }
            launch_update_centroids(r, dtype, d, list_off, list_rows, nlist, c, nullptr, nullptr);
            if (host) HB_CUDA(cudaMemcpyAsync(centroids, c, (size_t)nlist * d * 8, cudaMemcpyDeviceToHost, g_stream));
        }
        sync_stream();
    });
}

HB_API int hb_kmeans(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist, int32_t iters,
                     int64_t seed, const int64_t *seed_rows, double *out_centroids, int32_t *out_assign) {
    return guarded([&] {
        ensure_init();
        ivf_common_checks(rows, n, d, dtype, metric, nlist);
        HB_REQUIRE(iters >= 0, "max-iterations must be >= 0");
        HB_REQUIRE(out_centroids && out_assign, "null buffer");
        const void *r = stage_in(rows, (size_t)n * d * dtype_size(dtype), g_ws.in_a);
        const int64_t *sr = seed_rows ? (const int64_t *)stage_in(seed_rows, (size_t)nlist * 8, g_ws.in_c) : nullptr;
        if (sr) {
            Validator v;
            v.range_i64(sr, nlist, 0, n);
            v.require("hb_kmeans: seed row out of range");
        }
        OutStage oc = stage_out(out_centroids, (size_t)nlist * d * 8, g_ws.out_a);
        OutStage oa = stage_out(out_assign, (size_t)n * 4, g_ws.out_b);
        kmeans_exact(r, dtype, n, d, metric, nlist, iters, seed, sr, (double *)oc.dev, (int32_t *)oa.dev, nullptr);
        finish_out(oc);
        finish_out(oa);
        sync_stream();
    });
}

// global seed rows on the device, checked against the global row count
static const int64_t *stage_global_seeds(const int64_t *seed_rows, int nlist, int64_t n_total) {
    HB_REQUIRE(seed_rows, "sharded k-means: seed_rows (global row ids) are required");
    const int64_t *sr = (const int64_t *)stage_in(seed_rows, (size_t)nlist * 8, g_ws.in_c);
    Validator v;
    v.range_i64(sr, nlist, 0, n_total);
    v.require("sharded k-means: seed row out of range");
    return sr;
}
static int64_t global_row_count(int64_t n_local, int64_t first_row) {
    // every rank passes its block [first_row, first_row + n_local): the total is the largest end
    double *x = g_ws.sh_small.as<double>(8);
    const double end = (double)(first_row + n_local);
    HB_CUDA(cudaMemcpyAsync(x, &end, 8, cudaMemcpyHostToDevice, g_stream));
    comm_allreduce_max_f64(x, 1);
    double tot = 0;
    HB_CUDA(cudaMemcpyAsync(&tot, x, 8, cudaMemcpyDeviceToHost, g_stream));
    sync_stream();
    return (int64_t)tot;
}

HB_API int hb_sharded_kmeans(const void *rows, int64_t n_local, int32_t d, int dtype, int metric, int32_t nlist, int32_t iters,
                             const int64_t *seed_rows, int64_t first_global_row, double *out_centroids, int32_t *out_assign) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(comm_info().inited, "hb_sharded_kmeans before hb_comm_init");
        ivf_common_checks(rows, n_local, d, dtype, metric, nlist);
        HB_REQUIRE(iters >= 0 && first_global_row >= 0, "bad arguments");
        HB_REQUIRE(out_centroids && out_assign, "null buffer");
        const void *r = stage_in(rows, (size_t)n_local * d * dtype_size(dtype), g_ws.in_a);
        const int64_t *sr = stage_global_seeds(seed_rows, nlist, global_row_count(n_local, first_global_row));
        OutStage oc = stage_out(out_centroids, (size_t)nlist * d * 8, g_ws.out_a);
        OutStage oa = stage_out(out_assign, (size_t)n_local * 4, g_ws.out_b);
        kmeans_sharded(r, dtype, n_local, d, metric, nlist, iters, sr, first_global_row, (double *)oc.dev, (int32_t *)oa.dev);
        finish_out(oc);
        finish_out(oa);
        sync_stream();
    });
}

HB_API int hb_sharded_ivf_build(const void *rows, int64_t n_local, int32_t d, int dtype, int metric, int32_t nlist, int32_t iters,
                                const int64_t *seed_rows, int64_t first_global_row, hb_index **out) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(comm_info().inited, "hb_sharded_ivf_build before hb_comm_init");
        HB_REQUIRE(out, "null out");
        ivf_common_checks(rows, n_local, d, dtype, metric, nlist);
        HB_REQUIRE(iters >= 0 && first_global_row >= 0, "bad arguments");
        hb_index *ix = new hb_index();
        try {
            ix->type = HB_INDEX_IVF_FLAT;
            ix->dtype = dtype;
            ix->metric = metric;
            ix->d = d;
            ix->n = n_local;
            ix->nlist = nlist;
            ix->id_base = first_global_row;
            ix->coarse_sharded = true;
            const void *r = stage_in(rows, (size_t)n_local * d * dtype_size(dtype), g_ws.in_a);
            const int64_t *sr = stage_global_seeds(seed_rows, nlist, global_row_count(n_local, first_global_row));
            double *cents = ix->cents.as<double>((size_t)nlist * d);
            int32_t *assign = ix->assign.as<int32_t>(n_local);
            kmeans_sharded(r, dtype, n_local, d, metric, nlist, iters, sr, first_global_row, cents, assign);
            ivf_finalize(ix, r, (const double *)g_ws.misc.p);
        } catch (...) {
            ix->release();
            delete ix;
            throw;
        }
        *out = ix;
    });
}

HB_API int hb_hnsw_create(const void *rows, int64_t n, int32_t d, int dtype, int metric, const int32_t *levels,
                          int32_t max_level, int32_t entry_point, const int64_t *const *level_offsets,
                          const int32_t *const *level_ids, hb_index **out) {
    return guarded([&] {
        ensure_init();
        check_dtype(dtype);
        HB_REQUIRE(out, "null out");
        HB_REQUIRE(n >= 0 && d >= 1 && (n == 0 || (rows && levels)), "bad arguments");
        HB_REQUIRE(metric == HB_COSINE || metric == HB_L2, "HNSW distance-fn must be cosine or euclidean");
        HB_REQUIRE(n == 0 || (max_level >= 0 && entry_point >= 0 && entry_point < n && level_offsets && level_ids), "bad graph");
        hb_index *ix = new hb_index();
        try {
            ix->type = HB_INDEX_HNSW;
            ix->dtype = dtype;
            ix->metric = metric;
            ix->d = d;
            ix->n = n;
            ix->max_level = n ? max_level : 0;
            ix->entry = n ? entry_point : -1;
            const size_t bytes = (size_t)n * d * dtype_size(dtype);
            void *r = ix->rows.get(std::max<size_t>(bytes, 16));
            if (bytes) HB_CUDA(cudaMemcpyAsync(r, rows, bytes, cudaMemcpyDefault, g_stream));
            double *nm = ix->norms.as<double>(std::max<int64_t>(n, 1));
            launch_row_norms(r, dtype, n, d, nm);
            if (n) {
                int32_t *lv = ix->levels.as<int32_t>(n);
                HB_CUDA(cudaMemcpyAsync(lv, levels, (size_t)n * 4, cudaMemcpyDefault, g_stream));
                Validator val;
                val.range_i32(lv, n, 0, (int64_t)max_level + 1);
                val.require("hb_hnsw_create: node level out of range");
                ix->adj_off.resize((size_t)max_level + 1);
                ix->adj_ids.resize((size_t)max_level + 1);
                for (int l = 0; l <= max_level; ++l) {
                    HB_REQUIRE(level_offsets[l] && level_ids[l], "null adjacency level");
                    int64_t *o = ix->adj_off[l].as<int64_t>(n + 1);
                    HB_CUDA(cudaMemcpyAsync(o, level_offsets[l], (size_t)(n + 1) * 8, cudaMemcpyDefault, g_stream));
                    int64_t tot = 0;
                    if (is_device_ptr(level_offsets[l])) {
                        HB_CUDA(cudaMemcpyAsync(&tot, level_offsets[l] + n, 8, cudaMemcpyDeviceToHost, g_stream));
                        sync_stream();
                    } else {
                        tot = level_offsets[l][n];
                    }
                    HB_REQUIRE(tot >= 0 && tot <= n * (int64_t)4096, "hb_hnsw_create: bad adjacency size");
                    val.offsets(o, n + 1, tot);
                    val.require("hb_hnsw_create: adjacency offsets must start at 0 and not decrease");
                    int32_t *a = ix->adj_ids[l].as<int32_t>(std::max<int64_t>(tot, 1));
                    if (tot) HB_CUDA(cudaMemcpyAsync(a, level_ids[l], (size_t)tot * 4, cudaMemcpyDefault, g_stream));
                    val.csr(o, a, n);
                    val.require("hb_hnsw_create: neighbour id out of range");
                }
                std::vector<const void *> po((size_t)max_level + 1), pi((size_t)max_level + 1);
                for (int l = 0; l <= max_level; ++l) po[l] = ix->adj_off[l].p, pi[l] = ix->adj_ids[l].p;
                HB_CUDA(cudaMemcpyAsync(ix->adj_off_ptrs.as<const void *>(po.size()), po.data(), po.size() * sizeof(void *),
                                        cudaMemcpyHostToDevice, g_stream));
                HB_CUDA(cudaMemcpyAsync(ix->adj_ids_ptrs.as<const void *>(pi.size()), pi.data(), pi.size() * sizeof(void *),
                                        cudaMemcpyHostToDevice, g_stream));
                sync_stream();  // po / pi are locals
            }
            sync_stream();
        } catch (...) {
            ix->release();
            delete ix;
            throw;
        }
        *out = ix;
    });
}

HB_API int hb_gather_score(hb_index *index, const void *queries, int qdtype, int64_t nq, const int32_t *pair_query,
                           const int32_t *pair_row, int64_t npairs, double *out_scores) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(index, "null index");
        HB_REQUIRE(qdtype == HB_F32 || qdtype == HB_F64, "queries must be fp32 or fp64");
        HB_REQUIRE(nq >= 0 && npairs >= 0, "bad arguments");
        if (npairs == 0) return;
        HB_REQUIRE(queries && pair_query && pair_row && out_scores, "null buffer");
        HB_REQUIRE(index->type != HB_INDEX_IVF_FLAT, "gather-score needs rows in original order (flat or HNSW index)");
        const int d = index->d;
        const void *q = stage_in(queries, (size_t)nq * d * dtype_size(qdtype), g_ws.in_b);
        const int32_t *pq = (const int32_t *)stage_in(pair_query, (size_t)npairs * 4, g_ws.in_c);
        const int32_t *pr = (const int32_t *)stage_in(pair_row, (size_t)npairs * 4, g_ws.in_d);
        OutStage o = stage_out(out_scores, (size_t)npairs * 8, g_ws.out_a);
        {  // an unknown id is an exception in the reference (ultra_fast.clj:189-192), never an out-of-bounds read here
            Validator v;
            v.range_i32(pq, npairs, 0, nq);
            v.range_i32(pr, npairs, 0, index->n);
            v.require("hb_gather_score: pair_query / pair_row out of range");
        }
        bool l2;
        int epi;
        metric_to_epi(index->metric, true, l2, epi);
        if (index->metric == HB_IP) epi = EPI_DOT;
        const double *qn = nullptr;
        if (index->metric == HB_COSINE) {
            double *x = g_ws.qnorm.as<double>(nq);
            launch_row_norms(q, qdtype, nq, d, x);
            qn = x;
        }
        launch_gather_score(index->rows.p, index->dtype, (const double *)index->norms.p, q, qdtype, qn, d, pq, pr, npairs, l2, epi,
                            (double *)o.dev);
        finish_out(o);
        sync_stream();
    });
}

HB_API int hb_topk_merge(const double *dist, const int64_t *ids, int32_t nparts, int64_t nq, int32_t k, int64_t *out_ids,
                         double *out_dist) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(nparts >= 1 && nq >= 0 && k >= 0, "bad arguments");
        if (nq == 0 || k == 0) return;
        HB_REQUIRE(dist && ids && out_ids && out_dist, "null buffer");
        HB_REQUIRE(k <= 1024, "k > 1024 is not supported");
        const size_t cnt = (size_t)nparts * nq * k;
        const double *dd = (const double *)stage_in(dist, cnt * 8, g_ws.in_a);
        const int64_t *di = (const int64_t *)stage_in(ids, cnt * 8, g_ws.in_b);
        OutStage oi = stage_out(out_ids, (size_t)nq * k * 8, g_ws.out_a);
        OutStage od = stage_out(out_dist, (size_t)nq * k * 8, g_ws.out_b);
        double *cval = g_ws.cand_val.as<double>(cnt);
        int64_t *cid = g_ws.cand_id.as<int64_t>(cnt);
        launch_parts_to_query_major(dd, di, nparts, nq, k, cval, cid);
        // unused slots of a part carry id -1 / +inf: they sort last among equal distances only if real +inf
        // distances are absent, which holds for every finite input
        SelectParams L;
        L.vals = cval;
        L.nseg = nq;
        L.seg_stride = (int64_t)nparts * k;
        L.seg_len_const = (int64_t)nparts * k;
        L.k = k;
        L.out_val = (double *)od.dev;
        int64_t *pos = g_ws.sel_pos.as<int64_t>((size_t)nq * k);
        L.out_pos = pos;
        launch_select(L);
        launch_lookup_ids(pos, nq, k, cid, (int64_t)nparts * k, (int64_t *)oi.dev);
        finish_out(oi);
        finish_out(od);
        sync_stream();
    });
}

HB_API int hb_lsh_matrices(int32_t d, int32_t ntables, int32_t proj_dim, int64_t seed, double *out) {
    return guarded([&] {
        HB_REQUIRE(d >= 1 && ntables >= 1 && proj_dim >= 1 && out, "bad arguments");
        JavaGaussian g(seed);  // host arithmetic only: no device is needed
        const int64_t count = (int64_t)ntables * proj_dim * d;
        for (int64_t i = 0; i < count; ++i) out[i] = g.next();
    });
}

// ---- a4: the float[] Vector-API variants (src/hnsw/simd.clj:18-115) and PCAF (src/hnsw/ann/dimreduct/pcaf.clj) ------------
HB_API int hb_pairwise_f32lanes(const float *a, int64_t na, const float *b, int64_t nb, int32_t d, int metric, int32_t lanes,
                                double *out) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(na >= 0 && nb >= 0 && d >= 1 && lanes >= 1 && lanes <= 64, "bad arguments");
        HB_REQUIRE(metric == HB_COSINE || metric == HB_L2 || metric == HB_IP, "unknown metric");
        if (na == 0 || nb == 0) return;
        HB_REQUIRE(a && b && out, "null buffer");
        const float *da = (const float *)stage_in(a, (size_t)na * d * 4, g_ws.in_a);
        const float *db = (const float *)stage_in(b, (size_t)nb * d * 4, g_ws.in_b);
        OutStage o = stage_out(out, (size_t)na * nb * 8, g_ws.out_a);
        launch_lanes_pairwise(da, na, db, nb, d, metric, lanes, (double *)o.dev, nb);
        finish_out(o);
        sync_stream();
    });
}

HB_API int hb_pcaf_matrix(int32_t original_dim, int32_t target_dim, int64_t seed, float *out) {
    return guarded([&] {
        HB_REQUIRE(original_dim >= 1 && target_dim >= 1 && out, "bad arguments");
        // create-random-projection (pcaf.clj:33-46): scale = (float)(1 / sqrt(target)); two floats multiply in double in Clojure
        // and aset narrows the product
        JavaGaussian g(seed);
        const float scale = (float)(1.0 / std::sqrt((double)target_dim));
        const int64_t count = (int64_t)original_dim * target_dim;
        for (int64_t i = 0; i < count; ++i) {
            const float x = (float)g.next();
            out[i] = (float)((double)scale * (double)x);
        }
    });
}

HB_API int hb_pcaf_project(const float *matrix, int32_t original_dim, int32_t target_dim, const float *rows, int64_t n,
                           int32_t lanes, float *out) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(original_dim >= 1 && target_dim >= 1 && n >= 0 && lanes >= 1 && lanes <= 64, "bad arguments");
        if (n == 0) return;
        HB_REQUIRE(matrix && rows && out, "null buffer");
        const float *dm = (const float *)stage_in(matrix, (size_t)original_dim * target_dim * 4, g_ws.in_a);
        const float *dr = (const float *)stage_in(rows, (size_t)n * original_dim * 4, g_ws.in_b);
        OutStage o = stage_out(out, (size_t)n * target_dim * 4, g_ws.out_a);
        launch_lanes_project(dm, target_dim, dr, n, original_dim, lanes, (float *)o.dev);
        finish_out(o);
        sync_stream();
    });
}

// search-pcaf-parallel (pcaf.clj:195-253) for a batch: low-dimensional scan of every row, stable top min(k_filter, 3k),
// full-dimension re-rank of those, stable sort in candidate order, take k.
HB_API int hb_pcaf_search(hb_index *high, hb_index *low, const float *queries, const float *low_queries, int64_t nq, int32_t k,
                          int32_t k_filter, int32_t lanes, int64_t *out_ids, double *out_dist) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(high && low && high->type == HB_INDEX_FLAT && low->type == HB_INDEX_FLAT, "PCAF needs two flat indexes");
        HB_REQUIRE(high->dtype == HB_F32 && low->dtype == HB_F32 && high->n == low->n, "PCAF stores float[] rows (doubles-to-floats, pcaf.clj:83-90)");
        HB_REQUIRE(nq >= 0 && k >= 0 && k_filter >= 1 && lanes >= 1 && lanes <= 64, "bad arguments");
        if (nq == 0 || k == 0) return;
        HB_REQUIRE(queries && low_queries && out_ids && out_dist, "null buffer");
        HB_REQUIRE(k <= 1024 && k_filter <= 1024, "k / k-filter > 1024 is not supported");
        const int64_t n = high->n;
        const int d = high->d, t = low->d;
        OutStage oi = stage_out(out_ids, (size_t)nq * k * 8, g_ws.out_a);
        OutStage od = stage_out(out_dist, (size_t)nq * k * 8, g_ws.out_b);
        if (n == 0) {
            fill_empty_results((int64_t *)oi.dev, (double *)od.dev, nq * k);
            finish_out(oi);
            finish_out(od);
            sync_stream();
            return;
        }
        const float *q = (const float *)stage_in(queries, (size_t)nq * d * 4, g_ws.in_b);
        const float *lq = (const float *)stage_in(low_queries, (size_t)nq * t * 4, g_ws.in_c);
        const int c = (int)std::min<int64_t>(std::min<int64_t>(k_filter, 3ll * k), n);  // (take (min k-filter (* 3 k)) ...), :229-230
        int64_t qc = (int64_t)(scratch_budget() / 8 / (size_t)n);
        qc = std::max<int64_t>(1, std::min(qc, nq));
        double *scratch = g_ws.scratch.as<double>((size_t)qc * n);
        double *cval = g_ws.cand_val.as<double>((size_t)qc * c);
        int64_t *cpos = g_ws.cand_id.as<int64_t>((size_t)qc * c);
        double *exact = g_ws.sel_val.as<double>((size_t)qc * c);
        int64_t *pos2 = g_ws.sel_pos.as<int64_t>((size_t)qc * k);
        for (int64_t q0 = 0; q0 < nq; q0 += qc) {
            const int64_t nqc = std::min(qc, nq - q0);
            // phase 1 (:214-230): cosine-distance-simd of the projected query against every projected row
            launch_lanes_pairwise(lq + q0 * t, nqc, (const float *)low->rows.p, n, t, HB_COSINE, lanes, scratch, n);
            SelectParams L;
            L.vals = scratch;
            L.nseg = nqc;
            L.seg_stride = n;
            L.seg_len_const = n;
            L.k = c;
            L.out_val = cval;
            L.out_pos = cpos;
            launch_select(L);
            // phase 2 (:236-243): exact distance in the full dimension, candidates in phase-1 order
            launch_lanes_gather(q + q0 * d, (const float *)high->rows.p, d, lanes, cpos, nqc, c, exact);
            SelectParams M;  // stable sort of the candidate list + take k (:246-252): ties keep the phase-1 order
            M.vals = exact;
            M.nseg = nqc;
            M.seg_stride = c;
            M.seg_len_const = c;
            M.k = k;
            M.out_val = (double *)od.dev + q0 * k;
            M.out_pos = pos2;
            launch_select(M);
            launch_lookup_ids(pos2, nqc, k, cpos, c, (int64_t *)oi.dev + q0 * k);
        }
        finish_out(oi);
        finish_out(od);
        sync_stream();
    });
}

// The ordered fp64 sum + pick of one k-means++ step on given weights (hb_kpp.cu), for the parity tests:
// total = w_0 + w_1 + ... added left to right in fp64 (ivf_flat.clj:51-52), pick = first i with running sum >= u * total (:53-58).
HB_API int hb_kpp_sum_pick(const double *weights, int64_t n, double u, double *out_total, int64_t *out_pick) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(weights && n >= 1 && out_total && out_pick, "bad arguments");
        const int64_t nchunks = ceil_div(n, kKppChunk);
        const double *w = (const double *)stage_in(weights, (size_t)n * 8, g_ws.in_a);
        char *base = (char *)g_kpp_buf.get((size_t)(5 * nchunks + 16) * 8 + 1024);
        KppScaleParams K;
        K.n = n;
        K.linear = true;  // the weights as they are
        K.mind = const_cast<double *>(w);
        K.c_approx = (double *)base;
        K.c_prefix = K.c_approx + nchunks;
        K.c_start = K.c_prefix + nchunks;
        K.c_q = (long long *)(K.c_start + nchunks + 1);
        K.c_exp = (int *)(K.c_q + nchunks);
        double *misc = (double *)(base + (size_t)(5 * nchunks + 8) * 8);
        HB_CUDA(cudaMemcpyAsync(misc, &u, 8, cudaMemcpyHostToDevice, g_stream));
        K.u = misc;
        K.total = misc + 1;
        K.pick = (int64_t *)(misc + 2);
        launch_kpp_sum_pick(K);
        HB_CUDA(cudaMemcpyAsync(out_total, K.total, 8, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaMemcpyAsync(out_pick, K.pick, 8, cudaMemcpyDeviceToHost, g_stream));
        sync_stream();
    });
}

HB_API int hb_fast_scores(hb_index *index, const void *queries, int qdtype, int64_t nq, float *out_scores, double *out_scale,
                          double *out_eps) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(index && index->type == HB_INDEX_FLAT, "hb_fast_scores needs a flat index");
        HB_REQUIRE(qdtype == HB_F32 || qdtype == HB_F64, "queries must be fp32 or fp64");
        HB_REQUIRE(fast_metric_ok(index->metric), "FAST mode serves cosine and inner-product");
        HB_REQUIRE(nq >= 1 && index->n >= 1 && queries && out_scores, "bad arguments");
        const int d = index->d;
        const bool cosine = index->metric == HB_COSINE;
        const void *q = stage_in(queries, (size_t)nq * d * dtype_size(qdtype), g_ws.in_b);
        FastSideBufs &S = fast_rows_side(index, cosine);
        double *qn = g_ws.qnorm.as<double>(nq);
        launch_row_norms(q, qdtype, nq, d, qn);
        fast_quant_queries(q, qdtype, nq, d);
        FastPlan E, T;
        flat_fast_plan(nq, E, T, g_fw.flat_plan);
        UnitPlan U = make_units(E, g_fw.u_list, g_fw.u_sel0, g_fw.u_nsel, g_fw.u_ntile, g_fw.u_item0, g_fw.u_slotq, g_fw.u_slotrel,
                                (const int64_t *)S.tile_off.p);
        int8_t *aimg = g_fw.aimg.as<int8_t>((size_t)E.nunits * S.kbn * S.ns * kFastImg);
        launch_pack_units((const int8_t *)g_fw.dig.p, S.kbn, S.ns, E.nunits, U.slot_query, aimg);
        const size_t cnt = (size_t)E.nunits * S.ntiles * kFastTile * kFastTile;
        OutStage os = stage_out(out_scores, cnt * 4, g_ws.out_a);
        OutStage oc = stage_out(out_scale, (size_t)nq * 8, g_ws.out_b);
        OutStage oe = stage_out(out_eps, (size_t)nq * 8, g_ws.out_c);
        double *qscale = g_fw.qscale.as<double>(nq), *qeps = g_fw.qeps.as<double>(nq);
        launch_query_bounds((const double *)g_fw.qu.p, (const double *)g_fw.ql1.p, qn, nq, S.ns, d, index->metric,
                            (const float *)S.stats.p, qscale, qeps, g_fw.qmargin.as<float>(nq));
        TcParams P;
        P.aimg = aimg;
        P.bimg = (const int8_t *)S.img.p;
        P.kbn = S.kbn;
        P.nunits = E.nunits;
        P.unit_list = U.unit_list;
        P.unit_nsel = g_tc_half_m ? U.unit_nsel : nullptr;
        P.unit_item0 = U.unit_item0;
        P.tile_off = (const int64_t *)S.tile_off.p;
        P.list_off = (const int64_t *)S.list_off.p;
        P.slot_query = U.slot_query;
        P.slot_rel0 = U.slot_rel0;
        P.rs = (const float *)S.rs.p;
        P.ro = (const float *)S.ro.p;
        P.dump = (float *)os.dev;
        launch_tc_pass(P, S.ns, FAST_DUMP);
        if (oc.dev) HB_CUDA(cudaMemcpyAsync(oc.dev, qscale, (size_t)nq * 8, cudaMemcpyDeviceToDevice, g_stream));
        if (oe.dev) HB_CUDA(cudaMemcpyAsync(oe.dev, qeps, (size_t)nq * 8, cudaMemcpyDeviceToDevice, g_stream));
        finish_out(os);
        finish_out(oc);
        finish_out(oe);
        sync_stream();
    });
}

// ---- persistence of the device layout -----------------------------------------------------------------
// The reference persists only the HNSW graph, as EDN text (pr-str of every double, src/hnsw/helper/index_io.clj:10-80:
// 492.9 MB for 31 k vectors, README.md:22) and has no save / load for IVF-FLAT at all.  Here the file is the device layout
// itself: a fixed header, then tagged sections holding the index's device arrays verbatim (list-major slab, fp64 norms,
// centroids, list offsets, ...), so loading is file -> pinned staging -> HBM with no re-clustering, no norm pass and no
// re-ordering.  The FAST-mode digit images are derived data and are rebuilt on first use.
namespace {
constexpr uint64_t kFileMagic = 0x3130304958494248ull;  // "HBIXI001" little-endian
constexpr size_t kIoChunk = 64u << 20;
struct FileHeader {
    uint64_t magic;
    int32_t version, type, dtype, metric, d, nlist, max_level, entry;
    int64_t n, max_list;
    int32_t nsections, reserved;
};
enum SectionTag : uint32_t { SEC_ROWS = 1, SEC_NORMS, SEC_CENTS, SEC_CENT_NORM, SEC_LIST_OFF, SEC_LIST_ROWS, SEC_ASSIGN, SEC_LEVELS,
                             SEC_ADJ_OFF = 0x100, SEC_ADJ_IDS = 0x200 };  // + level
struct SectionHeader {
    uint32_t tag, reserved;
    uint64_t bytes;
};
struct PinnedChunk {
    void *p = nullptr;
    PinnedChunk() {
        if (cudaMallocHost(&p, kIoChunk) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            throw Error(HB_ERR_OOM, "cudaMallocHost of the I/O staging chunk failed");
        }
    }
    ~PinnedChunk() {
        if (p) cudaFreeHost(p);
    }
};
struct File {
    FILE *f = nullptr;
    File(const char *path, const char *mode) : f(fopen(path, mode)) {
        if (!f) throw Error(HB_ERR_INVALID, std::string("cannot open ") + path + ": " + strerror(errno));
    }
    // flush + fsync + close with every result checked: only then may the temporary file replace the target
    void commit() {
        FILE *g = f;
        f = nullptr;
        const bool ok = fflush(g) == 0 && fsync(fileno(g)) == 0;
        const bool closed = fclose(g) == 0;
        if (!ok || !closed) throw Error(HB_ERR_INVALID, std::string("write failed: ") + strerror(errno));
    }
    ~File() {
        if (f) fclose(f);
    }
};
void write_section(FILE *f, uint32_t tag, const void *dev, size_t bytes, PinnedChunk &pin) {
    SectionHeader sh{tag, 0, (uint64_t)bytes};
    if (fwrite(&sh, sizeof(sh), 1, f) != 1) throw Error(HB_ERR_INVALID, "write failed");
    for (size_t o = 0; o < bytes; o += kIoChunk) {
        const size_t c = std::min(kIoChunk, bytes - o);
        HB_CUDA(cudaMemcpyAsync(pin.p, (const char *)dev + o, c, cudaMemcpyDeviceToHost, g_stream));
        HB_CUDA(cudaStreamSynchronize(g_stream));
        if (fwrite(pin.p, 1, c, f) != c) throw Error(HB_ERR_INVALID, "write failed (disk full?)");
    }
}
void read_section(FILE *f, void *dev, size_t bytes, PinnedChunk &pin) {
    for (size_t o = 0; o < bytes; o += kIoChunk) {
        const size_t c = std::min(kIoChunk, bytes - o);
        if (fread(pin.p, 1, c, f) != c) throw Error(HB_ERR_INVALID, "index file is truncated");
        HB_CUDA(cudaMemcpyAsync((char *)dev + o, pin.p, c, cudaMemcpyHostToDevice, g_stream));
        HB_CUDA(cudaStreamSynchronize(g_stream));
    }
}
}  // namespace

HB_API int hb_index_save(const hb_index *index, const char *path) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(index && path, "null argument");
        const hb_index &ix = *index;
        const int64_t n = ix.n;
        struct Sec {
            uint32_t tag;
            const void *p;
            size_t bytes;
        };
        std::vector<Sec> secs;
        secs.push_back({SEC_ROWS, ix.rows.p, (size_t)n * ix.d * dtype_size(ix.dtype)});
        secs.push_back({SEC_NORMS, ix.norms.p, (size_t)n * 8});
        if (ix.type == HB_INDEX_IVF_FLAT) {
            secs.push_back({SEC_CENTS, ix.cents.p, (size_t)ix.nlist * ix.d * 8});
            secs.push_back({SEC_CENT_NORM, ix.cent_norm.p, (size_t)ix.nlist * 8});
            secs.push_back({SEC_LIST_OFF, ix.list_off.p, ((size_t)ix.nlist + 1) * 8});
            secs.push_back({SEC_LIST_ROWS, ix.list_rows.p, (size_t)n * 8});
            secs.push_back({SEC_ASSIGN, ix.assign.p, (size_t)n * 4});
        } else if (ix.type == HB_INDEX_HNSW && n > 0) {
            secs.push_back({SEC_LEVELS, ix.levels.p, (size_t)n * 4});
            for (int l = 0; l <= ix.max_level; ++l) {
                int64_t tot = 0;
                HB_CUDA(cudaMemcpyAsync(&tot, (const int64_t *)ix.adj_off[l].p + n, 8, cudaMemcpyDeviceToHost, g_stream));
                sync_stream();
                secs.push_back({SEC_ADJ_OFF + (uint32_t)l, ix.adj_off[l].p, ((size_t)n + 1) * 8});
                secs.push_back({SEC_ADJ_IDS + (uint32_t)l, ix.adj_ids[l].p, (size_t)tot * 4});
            }
        }
        FileHeader h{};
        h.magic = kFileMagic;
        h.version = 1;
        h.type = ix.type, h.dtype = ix.dtype, h.metric = ix.metric, h.d = ix.d;
        h.nlist = ix.nlist, h.max_level = ix.max_level, h.entry = ix.entry;
        h.n = n, h.max_list = ix.max_list;
        h.nsections = (int32_t)secs.size();
        const std::string tmp = std::string(path) + ".tmp";
        {
            File f(tmp.c_str(), "wb");
            PinnedChunk pin;
            if (fwrite(&h, sizeof(h), 1, f.f) != 1) throw Error(HB_ERR_INVALID, "write failed");
            try {
                for (const Sec &s : secs) write_section(f.f, s.tag, s.p, s.bytes, pin);
                f.commit();
            } catch (...) {
                remove(tmp.c_str());
                throw;
            }
        }
        if (rename(tmp.c_str(), path) != 0) {
            remove(tmp.c_str());
            throw Error(HB_ERR_INVALID, std::string("cannot rename to ") + path);
        }
    });
}

HB_API int hb_index_load(const char *path, hb_index **out) {
    return guarded([&] {
        ensure_init();
        HB_REQUIRE(path && out, "null argument");
        File f(path, "rb");
        FileHeader h{};
        if (fread(&h, sizeof(h), 1, f.f) != 1 || h.magic != kFileMagic) throw Error(HB_ERR_INVALID, "not an hnswb200 index file");
        HB_REQUIRE(h.version == 1, "unsupported index file version");
        HB_REQUIRE(h.type >= HB_INDEX_FLAT && h.type <= HB_INDEX_HNSW && h.n >= 0 && h.d >= 1 && h.nlist >= 0 && h.max_level >= 0 &&
                       h.max_level < 64 && h.nsections >= 0 && h.nsections < 1024,
                   "corrupt index file header");
        check_dtype(h.dtype);
        HB_REQUIRE(h.metric == HB_COSINE || h.metric == HB_L2 || h.metric == HB_IP, "corrupt index file header (metric)");
        HB_REQUIRE(h.n < (1ll << 40) && h.d <= (1 << 20), "corrupt index file header (size)");
        hb_index *ix = new hb_index();
        try {
            ix->type = h.type, ix->dtype = h.dtype, ix->metric = h.metric, ix->d = h.d;
            ix->n = h.n, ix->nlist = h.nlist, ix->max_list = h.max_list, ix->max_level = h.max_level, ix->entry = h.entry;
            const int64_t n = h.n;
            if (ix->type == HB_INDEX_HNSW && n > 0) {
                ix->adj_off.resize((size_t)h.max_level + 1);
                ix->adj_ids.resize((size_t)h.max_level + 1);
            }
            PinnedChunk pin;
            uint64_t seen = 0;  // bit per fixed tag
            std::vector<char> seen_off((size_t)h.max_level + 1, 0), seen_ids((size_t)h.max_level + 1, 0);
            std::vector<int64_t> adj_count((size_t)h.max_level + 1, 0);
            for (int s = 0; s < h.nsections; ++s) {
                SectionHeader sh{};
                if (fread(&sh, sizeof(sh), 1, f.f) != 1) throw Error(HB_ERR_INVALID, "index file is truncated");
                DevBuf *dst = nullptr;
                size_t want = 0;
                switch (sh.tag) {
                    case SEC_ROWS: dst = &ix->rows, want = (size_t)n * h.d * dtype_size(h.dtype); break;
                    case SEC_NORMS: dst = &ix->norms, want = (size_t)n * 8; break;
                    case SEC_CENTS: dst = &ix->cents, want = (size_t)h.nlist * h.d * 8; break;
                    case SEC_CENT_NORM: dst = &ix->cent_norm, want = (size_t)h.nlist * 8; break;
                    case SEC_LIST_OFF: dst = &ix->list_off, want = ((size_t)h.nlist + 1) * 8; break;
                    case SEC_LIST_ROWS: dst = &ix->list_rows, want = (size_t)n * 8; break;
                    case SEC_ASSIGN: dst = &ix->assign, want = (size_t)n * 4; break;
                    case SEC_LEVELS: dst = &ix->levels, want = (size_t)n * 4; break;
                    default: {
                        const uint32_t l = sh.tag & 0xff;
                        HB_REQUIRE(ix->type == HB_INDEX_HNSW && n > 0 && l <= (uint32_t)h.max_level, "unknown section in index file");
                        if ((sh.tag & ~0xffu) == SEC_ADJ_OFF) dst = &ix->adj_off[l], want = ((size_t)n + 1) * 8, seen_off[l] = 1;
                        else if ((sh.tag & ~0xffu) == SEC_ADJ_IDS) {
                            HB_REQUIRE(sh.bytes % 4 == 0 && sh.bytes <= (uint64_t)n * 4096 * 4, "corrupt adjacency section");
                            dst = &ix->adj_ids[l], want = sh.bytes, seen_ids[l] = 1;
                            adj_count[l] = (int64_t)(sh.bytes / 4);
                        }
                        else throw Error(HB_ERR_INVALID, "unknown section in index file");
                    }
                }
                HB_REQUIRE(sh.bytes == want, "section size does not match the header");
                if (sh.tag < 64) seen |= 1ull << sh.tag;
                void *p = dst->get(std::max<size_t>(want, 16));
                read_section(f.f, p, want, pin);
            }
            auto need = [&](uint32_t tag) { HB_REQUIRE(seen >> tag & 1, "index file lacks a required section"); };
            need(SEC_ROWS), need(SEC_NORMS);
            if (ix->type == HB_INDEX_IVF_FLAT) {
                HB_REQUIRE(n >= 1 && h.nlist >= 1, "corrupt index file header");
                need(SEC_CENTS), need(SEC_CENT_NORM), need(SEC_LIST_OFF), need(SEC_LIST_ROWS), need(SEC_ASSIGN);
                // the contents drive device addressing (slab offsets, row ids, scratch and grid sizes): check them, and take
                // max_list from the offsets rather than from the header
                Validator val;
                val.offsets((const int64_t *)ix->list_off.p, (int64_t)h.nlist + 1, n);
                val.range_i64((const int64_t *)ix->list_rows.p, n, 0, n);
                val.range_i32((const int32_t *)ix->assign.p, n, 0, h.nlist);
                val.require("corrupt index file (list offsets / row ids / assignments out of range)");
                std::vector<int64_t> off((size_t)h.nlist + 1);
                HB_CUDA(cudaMemcpyAsync(off.data(), ix->list_off.p, off.size() * 8, cudaMemcpyDeviceToHost, g_stream));
                sync_stream();
                ix->max_list = 0;
                for (int l = 0; l < h.nlist; ++l) ix->max_list = std::max(ix->max_list, off[(size_t)l + 1] - off[(size_t)l]);
            }
            if (ix->type == HB_INDEX_HNSW && n > 0) {
                need(SEC_LEVELS);
                HB_REQUIRE(h.entry >= 0 && h.entry < n, "corrupt index file header");
                std::vector<const void *> po((size_t)h.max_level + 1), pi((size_t)h.max_level + 1);
                Validator val;
                val.range_i32((const int32_t *)ix->levels.p, n, 0, (int64_t)h.max_level + 1);
                for (int l = 0; l <= h.max_level; ++l) {
                    HB_REQUIRE(seen_off[l] && seen_ids[l], "index file lacks an adjacency level");
                    po[l] = ix->adj_off[l].p, pi[l] = ix->adj_ids[l].p;
                    // offsets must be a CSR over exactly the ids stored for the level, the ids must be nodes
                    val.offsets((const int64_t *)ix->adj_off[l].p, n + 1, adj_count[l]);
                    val.require("corrupt index file (adjacency offsets / levels)");
                    val.csr((const int64_t *)ix->adj_off[l].p, (const int32_t *)ix->adj_ids[l].p, n);
                }
                val.require("corrupt index file (neighbour id out of range)");
                HB_CUDA(cudaMemcpyAsync(ix->adj_off_ptrs.as<const void *>(po.size()), po.data(), po.size() * sizeof(void *),
                                        cudaMemcpyHostToDevice, g_stream));
                HB_CUDA(cudaMemcpyAsync(ix->adj_ids_ptrs.as<const void *>(pi.size()), pi.data(), pi.size() * sizeof(void *),
                                        cudaMemcpyHostToDevice, g_stream));
                sync_stream();
            }
            sync_stream();
        } catch (...) {
            ix->release();
            delete ix;
            throw;
        }
        *out = ix;
    });
}

HB_API int hb_index_info(const hb_index *index, hb_info *out) {
    return guarded([&] {
        HB_REQUIRE(index && out, "null argument");
        out->type = index->type;
        out->dtype = index->dtype;
        out->metric = index->metric;
        out->dim = index->d;
        out->n = index->n;
        out->nlist = index->nlist;
        out->max_level = index->max_level;
        out->device_bytes = index->device_bytes();
    });
}

HB_API int hb_index_free(hb_index *index) {
    return guarded([&] {
        if (!index) return;
        if (g_inited) cudaStreamSynchronize(g_stream);
        index->release();
        delete index;
    });
}

}  // extern "C"
