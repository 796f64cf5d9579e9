/*
 * hnswb200.h — C ABI of libhnswb200.so: the B200 (sm_100a) distance core behind hnsw-clj's
 * build-index / search-knn / search-batch* surface.
 *
 * The reference (pure Clojure) has no FFI of its own; its extension points are Clojure functions.
 * Each entry point below names the reference interface it stands in for (file:line relative to the
 * reference repo).  INTEGRATION.md shows the java.lang.foreign downcall stubs a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, a negative hb_status otherwise; hb_last_error() gives the
 *     calling thread's last message.  There is NO CPU fallback: without a usable sm_100 device every
 *     compute call fails with HB_ERR_NO_DEVICE.
 *   - `rows`, `queries` and output buffers may be HOST or DEVICE pointers (detected with
 *     cudaPointerGetAttributes).  Host inputs are borrowed for the duration of the call only.
 *   - ids are ROW INDICES (position in the `rows` passed at build); the host shim owns the String ids
 *     (reference: ["vec_0" double[]] pairs, test/data_generator.clj:84-87).
 *   - results are ascending by distance, ties in the reference's stable-sort order (row order; for IVF
 *     probe rank first — src/hnsw/ann/partition/ivf_flat.clj:281-294).  Unused slots (k > candidates,
 *     test/hnsw/core_test.clj:90-96) hold id -1 and distance +inf.
 *   - values must be fp32-representable (HB_F32/HB_BF16 storage) unless HB_F64 is used: the reference
 *     stores embeddings as double[] holding fp32 values.
 */
#ifndef HNSWB200_H
#define HNSWB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_API __attribute__((visibility("default")))

typedef struct hb_index hb_index;

enum hb_status {
    HB_OK = 0,
    HB_ERR_INVALID = -1,   /* IllegalArgumentException in the reference (src/hnsw/api/simple.clj:13-14) */
    HB_ERR_NO_DEVICE = -2, /* no CUDA device / not sm_100 */
    HB_ERR_CUDA = -3,
    HB_ERR_OOM = -4,
    HB_ERR_UNSUPPORTED = -5
};

enum hb_dtype { HB_F32 = 0, HB_BF16 = 1, HB_F64 = 2 };

/* :distance-fn of the reference mapped to an enum (src/hnsw/api.clj:16-19): cosine-distance-ultra,
 * euclidean-distance-ultra; HB_IP ranks by descending dot-product (src/hnsw/simd_optimized.clj:283-293),
 * an extension for BASELINE config 3. */
enum hb_metric { HB_COSINE = 0, HB_L2 = 1, HB_IP = 2 };

enum hb_index_type { HB_INDEX_FLAT = 0, HB_INDEX_IVF_FLAT = 1, HB_INDEX_HNSW = 2 };

/* search arithmetic: HB_MODE_EXACT restates the reference's fp64 sequential sums on the device for every
 * pair; HB_MODE_FAST selects candidates on the tensor cores (exact int8-digit dot products, tcgen05), re-scores
 * them in fp64 in the reference's summation order and accepts a query only when an a-priori error bound proves
 * the top-k complete — otherwise the query is recomputed by the exact path.  Both modes return identical ids
 * and distance bits.  FAST serves cosine / inner-product searches with k <= 112 and the k-means
 * assignment passes (hb_kmeans_assign, hb_kmeans, hb_ivf_build); everything else runs the exact kernels. */
enum hb_mode { HB_MODE_EXACT = 0, HB_MODE_FAST = 1 };

typedef struct hb_info {
    int32_t type;   /* hb_index_type */
    int32_t dtype;  /* storage dtype of the rows */
    int32_t metric;
    int32_t dim;
    int64_t n;      /* rows */
    int32_t nlist;  /* IVF: partitions; else 0   (index-info :partitions, ivf_flat.clj:319-327) */
    int32_t max_level; /* HNSW: top level; else 0 */
    int64_t device_bytes;
} hb_info;

/* ---- process ------------------------------------------------------------------------------ */
HB_API int hb_init(int device);           /* select the device for this process (one process per GPU) */
HB_API int hb_shutdown(void);
HB_API const char *hb_last_error(void);
HB_API int hb_version(void);
HB_API int hb_set_stream(void *cuda_stream); /* launch on this stream (default: legacy stream 0) */
HB_API int hb_set_mode(int mode);            /* hb_mode for subsequent searches and builds (default EXACT); per index: hb_index_set_mode */
/* knobs: "scratch_mb" = budget for the transient distance scratch (default 8192); "profile" = 1 records
 * CUDA events around the main kernels (and resets the counters); "fast_digits" = 2 | 3 signed 8-bit digits per
 * element in the HB_MODE_FAST candidate pass (16- or 24-bit block-fixed-point mantissas); "fast_sample_tiles" = row
 * tiles (128 rows) of each query's nearest list scored first to seed the IVF scan's thresholds (default 2);
 * "fast_prune" = 0 scans every probed list in HB_MODE_FAST (default 1: lists that provably hold no top-k row of a query are
 * dropped, same results); "rowstream" = 0 keeps small batches (<= 8 queries) on the register-buffered scan instead of
 * the bulk-copy ring, "stream_seg" (256 | 384 | 512 bytes per row and stage), "stream_stages" (2..4), "stream_warps"
 * (0 = as many as fit) shape that ring; "micro_batch" = n combines concurrent hb_search calls of <= 8 queries on host buffers
 * into batches of up to n queries (default 64, 0 = off: parallel-search-futures, src/hnsw/helper/parallel_search.clj:15-49,
 * without a thread per query); "comm_p2p" = 0 exchanges the local top-k of hb_sharded_search by ncclAllGather instead of
 * the peer-window kernel; "kpp_scale" = 0 runs k-means++ as the one-thread prefix walk with an exhaustive distance pass
 * (default 1: pruned distance pass + chunked ordered sum, hb_kpp.cu); "tc_narrow" = 0 keeps sparse IVF units on the
 * 128 x 128 candidate kernel; "fast_carry" = 0 discards the threshold sample's candidates instead of keeping them;
 * "fast_level_min" (default 33 row tiles) = the length from which a flat FAST scan runs in levels, "fast_level_dense"
 * (default 8, 0 = off) = the row tiles of its first level, scored into a matrix and selected per query,
 * "fast_level_ratio" (default 16) = growth of the later levels; "host_feed" = B copies a host query batch in blocks of B
 * queries on a copy stream under the coarse stage (default 0: one copy, which measures faster); "prof_coarse" = 1 points
 * the stage timers at the coarse job of an IVF search.  Results never depend on a knob. */
HB_API int hb_set_option(const char *name, int64_t value);
/* measurements: "scan_ms"/"scan_count" (list/flat scan kernel), "coarse_ms", "select_ms", "plan_ms", "assign_ms",
 * "tc_ms" (tensor-core candidate pass over all probed lists), "tc_sample_ms" (its threshold-seeding pass), "pack_ms",
 * "rescore_ms"
 * (device time between the events, accumulated since "profile" was set), "fast_queries" / "fast_fallbacks"
 * (queries served in FAST mode / of those recomputed by the exact path), "fp64_peak_tflops" (runs a DFMA
 * microbenchmark: the measured peak of the pipe the exact kernels are bound by) */
HB_API int hb_get_stat(const char *name, double *out);
/* number of kernels this library launched since hb_init / the last reset (bench `gpu_launches`) */
HB_API int64_t hb_launch_count(int reset);

/* ---- distance core ------------------------------------------------------------------------ */
/* sqrt(sum v^2) per row, sequential fp64.  Replaces the norm precompute at
 * src/hnsw/ann/partition/ivf_flat.clj:161-179 and precompute-norms, src/hnsw/simd_optimized.clj:206-216. */
HB_API int hb_row_norms(const void *rows, int64_t n, int32_t d, int dtype, double *out_norms);

/* out[i*nb + j] = distance-fn(a_i, b_j): the batched form of the (fn [^doubles a ^doubles b]) :distance-fn
 * extension point (src/hnsw/ultra_fast.clj:43-95; batch-distances-parallel, src/hnsw/simd_optimized.clj:164-179).
 * HB_COSINE applies the zero-norm guard of cosine-distance-ultra (:92-95); HB_IP returns the dot product. */
HB_API int hb_pairwise(const void *a, int64_t na, int adtype, const void *b, int64_t nb, int bdtype,
                       int32_t d, int metric, double *out);

/* ---- flat exact search ---------------------------------------------------------------------- */
/* Device-resident copy of the rows (+ norms).  Replaces the data seq handed to compute-exact-knn
 * (src/hnsw/bench.clj:72-84) / top-k-distances (src/hnsw/simd_optimized.clj:271-280). */
HB_API int hb_flat_create(const void *rows, int64_t n, int32_t d, int dtype, int metric, hb_index **out);

/* ---- IVF-FLAT -------------------------------------------------------------------------------- */
/* build-ivf-flat-index (src/hnsw/ann/partition/ivf_flat.clj:137-211): norms, k-means++ with
 * java.util.Random(seed) (:32-60), `iters` Lloyd rounds + final assignment (:92-131), list-major slabs.
 * `metric` is the :distance-fn used for clustering and coarse routing; the list scan itself is always
 * cosine, as in the reference (:217-234). */
HB_API int hb_ivf_build(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist,
                        int32_t iters, int64_t seed, hb_index **out);
/* build-lightning-index with :smart-partition? true (src/hnsw/ann/partition/lightning.clj:46-160): seeds by the
 * k-means++ walk weighted with d_i (not d_i^2, :86-109, java.util.Random(seed)), every row to its nearest seed
 * (assign-to-partition :31-44), routing centroids = partition means, zero vector for an empty partition
 * (:122-126); no Lloyd rounds.  The result is an IVF-FLAT-type index: search-lightning (:184-298) is hb_search with
 * nprobe = max(1, (int)(partitions * percent)) (:262) — centroid routing, per-list cosine scan, stable merge.
 * The reference's default (:smart-partition? false) is an unseeded shuffle split (:132-137): pass such a partition
 * through hb_ivf_import. */
HB_API int hb_lightning_build(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist,
                              int64_t seed, hb_index **out);
/* Same index from given centroids (fp64 [nlist x d]) and per-row assignments: parity mode for
 * oracle-built partitions, and the load path for a persisted index. */
HB_API int hb_ivf_import(const void *rows, int64_t n, int32_t d, int dtype, int metric,
                         const double *centroids, int32_t nlist, const int32_t *assignments, hb_index **out);
/* centroids fp64 [nlist x d] and assignments int32 [n] of a built index (either may be NULL) */
HB_API int hb_ivf_export(const hb_index *index, double *out_centroids, int32_t *out_assignments);

/* ---- search ---------------------------------------------------------------------------------- */
/* search-knn / search-batch* for every index type (src/hnsw/ann/partition/ivf_flat.clj:305-317,
 * src/hnsw/ultra_fast.clj:346-374, src/hnsw/api/protocol.clj:58-67; replaces the per-query fan-out of
 * src/hnsw/helper/parallel_search.clj:15-49).  `queries` is [nq x d] in `qdtype` (HB_F32 or HB_F64).
 * `param` is nprobe for IVF (:num-probes, ivf_flat.clj:243-251), ef for HNSW (0 = the reference's
 * max(k,50), ultra_fast.clj:355), ignored for flat.  out_ids [nq x k] int64, out_dist [nq x k] fp64. */
HB_API int hb_search(hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t k, int32_t param,
                     int64_t *out_ids, double *out_dist);
/* IVF only: also return the probed list ids [nq x nprobe] in probe-rank order (-1 padded) */
HB_API int hb_ivf_probes(hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t nprobe,
                         int32_t *out_probes);

/* ---- k-means steps (IVF build) ------------------------------------------------------------- */
/* kmeans-plus-plus-init (ivf_flat.clj:32-60): chosen data-row indices, int64 [nlist] */
HB_API int hb_kmeanspp_init(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist,
                            int64_t seed, int64_t *out_seed_rows);
/* assign-to-nearest-centroid over all rows (ivf_flat.clj:79-90): strict <, lowest index wins.  In HB_MODE_FAST
 * (cosine, fp32 / fp64 rows) the rows are the queries of a k = 1 candidate pass over the centroids; same result. */
HB_API int hb_kmeans_assign(const void *rows, int64_t n, int32_t d, int dtype, int metric,
                            const double *centroids, int32_t nlist, int32_t *out_assign);
/* compute-centroid per cluster in row order, empty cluster keeps its centroid (ivf_flat.clj:66-77,112-116).
 * `centroids` is updated in place.  If out_sums/out_counts are non-NULL the per-cluster fp64 sums
 * [nlist x d] and counts [nlist] are returned instead of dividing (for the multi-GPU all-reduce). */
HB_API int hb_kmeans_update(const void *rows, int64_t n, int32_t d, int dtype, const int32_t *assign,
                            int32_t nlist, double *centroids, double *out_sums, int64_t *out_counts);
/* partition-vectors-kmeans (ivf_flat.clj:92-131): seeds (k-means++ unless seed_rows given), iters Lloyd
 * rounds, final assignment */
HB_API int hb_kmeans(const void *rows, int64_t n, int32_t d, int dtype, int metric, int32_t nlist,
                     int32_t iters, int64_t seed, const int64_t *seed_rows, double *out_centroids,
                     int32_t *out_assign);

/* ---- HNSW neighbour-candidate scoring ------------------------------------------------------ */
/* Upload a graph built on the host (src/hnsw/ultra_fast.clj:216-344): per-node level, and per level a CSR
 * adjacency over all n nodes (offsets int64 [n+1], neighbour ids int32) in the iteration order the search
 * must follow.  level_offsets[l] / level_ids[l] for l = 0..max_level. */
HB_API int hb_hnsw_create(const void *rows, int64_t n, int32_t d, int dtype, int metric, const int32_t *levels,
                          int32_t max_level, int32_t entry_point, const int64_t *const *level_offsets,
                          const int32_t *const *level_ids, hb_index **out);
/* scores[p] = distance-fn(query[pair_query[p]], row[pair_row[p]]): the call at
 * src/hnsw/ultra_fast.clj:192 batched over (query, neighbour-id) pairs (cf. batch-distance-calculation,
 * src/hnsw/wip/parallel_build.clj:106-118).  Works on any index type (uses its rows). */
HB_API int hb_gather_score(hb_index *index, const void *queries, int qdtype, int64_t nq,
                           const int32_t *pair_query, const int32_t *pair_row, int64_t npairs,
                           double *out_scores);

/* ---- fp32 Vector-API variants + PCAF ---------------------------------------------------------------------------------
 * cosine-distance-simd-optimized / euclidean-distance-simd-optimized / dot-product-simd-optimized on float[]
 * (src/hnsw/simd.clj:18-115): per chunk of `lanes` floats (= FloatVector/SPECIES_PREFERRED .length of the reference's
 * machine: 4, 8 or 16) fp32 lane products reduced in fp32, chunk sums accumulated in fp64, scalar tail in fp64.  The JDK
 * leaves the lane order of reduceLanes(ADD) unspecified; the device uses left-to-right (its scalar fallback / HotSpot's
 * ordered reduction), so results are within the north star's 1e-5 relative bound of any JVM and bit-identical to the
 * oracle's restatement.  out[i*nb + j] = metric(a_i, b_j); HB_IP = the dot product. */
HB_API int hb_pairwise_f32lanes(const float *a, int64_t na, const float *b, int64_t nb, int32_t d, int metric, int32_t lanes,
                                double *out);
/* create-random-projection (src/hnsw/ann/dimreduct/pcaf.clj:33-46): [target_dim][original_dim] floats,
 * (float)(1/sqrt(target)) * (float) Random(seed).nextGaussian(), on the HOST */
HB_API int hb_pcaf_matrix(int32_t original_dim, int32_t target_dim, int64_t seed, float *out);
/* project-vector-simd (pcaf.clj:48-81) for n rows: out [n][target_dim] floats */
HB_API int hb_pcaf_project(const float *matrix, int32_t original_dim, int32_t target_dim, const float *rows, int64_t n,
                           int32_t lanes, float *out);
/* search-pcaf-parallel (pcaf.clj:195-253), batched: `high` / `low` are flat indexes (hb_flat_create, HB_F32) over the
 * float rows and their projections; phase 1 scans every projected row (cosine-distance-simd), keeps the
 * min(k_filter, 3k) best in stable order, phase 2 re-ranks them in the full dimension; ties keep phase-1 order. */
HB_API int hb_pcaf_search(hb_index *high, hb_index *low, const float *queries, const float *low_queries, int64_t nq, int32_t k,
                          int32_t k_filter, int32_t lanes, int64_t *out_ids, double *out_dist);

/* ---- LSH projections ------------------------------------------------------------------------------ */
/* generate-random-matrix x NUM-HASH-TABLES (src/hnsw/ann/hash/hybrid_lsh.clj:24-31, :77-81): ntables matrices of
 * proj_dim x d doubles drawn in order from ONE java.util.Random(seed).nextGaussian stream (polar method over
 * StrictMath.log / sqrt, restated bit for bit), out [ntables][proj_dim][d] on the HOST.  The hashing itself
 * (compute-hash-vector, :33-45) is hb_pairwise with HB_IP against these rows, the bucket scan
 * (search-bucket-brute-force, :147-193) is hb_gather_score, the final sort + take k is hb_topk_merge. */
HB_API int hb_lsh_matrices(int32_t d, int32_t ntables, int32_t proj_dim, int64_t seed, double *out);

/* ---- k-means++ diagnostics ----------------------------------------------------------------------------------------
 * One k-means++ pick on given non-negative weights: out_total = the weights added left to right in fp64
 * (ivf_flat.clj:51-52), out_pick = the first i whose running sum reaches u * total (:53-58).  The device computes the
 * ordered sum with integer adds per chunk (hb_kpp.cu) and must return the bits of the sequential loop; used by the
 * parity tests to pin that machinery on adversarial inputs (ties, power-of-two crossings, huge dynamic range). */
HB_API int hb_kpp_sum_pick(const double *weights, int64_t n, double u, double *out_total, int64_t *out_pick);

/* ---- FAST-mode diagnostics ----------------------------------------------------------------------- */
/* The candidate pass of HB_MODE_FAST on a flat index, unfiltered: for every (query, row) the tensor-core score
 * (exact integer dot product of the quantised digits times the row scale).  Batched form of
 * batch-distances-parallel (src/hnsw/simd_optimized.clj:164-179) at candidate precision; used by the parity
 * tests to pin the tcgen05 path.  out_scores is [ceil(nq/128)][ceil(n/128)][128 query slots][128 rows] fp32;
 * similarity (cosine) or dot (HB_IP) of query q and row r = out_scores[q/128][r/128][q%128][r%128] * out_scale[q],
 * and |that - exact| <= out_eps[q].  out_scale / out_eps may be NULL. */
HB_API int hb_fast_scores(hb_index *index, const void *queries, int qdtype, int64_t nq, float *out_scores,
                          double *out_scale, double *out_eps);

/* ---- multi-GPU merge --------------------------------------------------------------------------- */
/* Merge `nparts` per-shard top-k lists (dist [nparts x nq x k], ids likewise, already global row ids)
 * into the global top-k: the sort-by :distance + take k at ivf_flat.clj:291-294 /
 * partitioned_hnsw.clj:187-196, ties by (part, position). */
HB_API int hb_topk_merge(const double *dist, const int64_t *ids, int32_t nparts, int64_t nq, int32_t k,
                         int64_t *out_ids, double *out_dist);

/* ---- multi-GPU: one process per GPU, rows sharded ------------------------------------------------------------------
 * The reference scales out by building independent sub-indexes over row ranges, searching all of them in parallel and
 * merging: (apply concat) + (sort-by :distance) + (take k), src/hnsw/ann/partition/partitioned_hnsw.clj:86-196.  Here rank g
 * of G (one process per GPU of an NVSwitch box) holds the global rows [first_global_row, first_global_row + n_local) of ONE
 * index: an exact flat shard, or the rows of that range inside every list of a global IVF-FLAT index whose centroids are
 * replicated.  hb_sharded_search = local search -> exchange of the G local top-k lists -> merge by (distance, rank,
 * position), i.e. the stable sort of the concatenation in rank order; with contiguous row blocks that is the single-GPU
 * (distance, row) order.  The exchange is ONE kernel launch per search: every rank writes its block into its peers' windows
 * over NVLink (cudaIpc-mapped memory), signals, waits and merges; hb_set_option("comm_p2p", 0) — or a box without peer
 * access — uses ncclAllGather + the same merge instead.  NCCL (libnccl.so.2, resolved with dlopen at hb_comm_init)
 * carries the bulk collectives of the build.  All collectives run on the library's stream; no host synchronisation
 * between the local search and the merge.  Every rank must make the same calls in the same order. */
#define HB_COMM_ID_BYTES 128
/* rank 0: an opaque id to hand to every rank by any means (file, environment, socket, torch.distributed) */
HB_API int hb_comm_unique_id(void *out_id /* HB_COMM_ID_BYTES */);
/* collective; after hb_init(device).  nranks <= 8 */
HB_API int hb_comm_init(const void *id, int32_t nranks, int32_t rank);
/* out_p2p: 1 once the peer windows are mapped (the first hb_sharded_search decides) */
HB_API int hb_comm_info(int32_t *out_nranks, int32_t *out_rank, int32_t *out_p2p);
HB_API int hb_comm_shutdown(void);
/* plumbing for the host side (seeds, centroids, timings): in place, host or device buffers */
HB_API int hb_comm_broadcast(void *buf, int64_t bytes, int32_t root);
HB_API int hb_comm_allreduce_f64(double *buf, int64_t count, int32_t op /* 0 = sum, 1 = max */);
/* the global row of this index's local row 0; hb_sharded_search adds it to every id */
HB_API int hb_index_set_id_base(hb_index *index, int64_t first_global_row);
/* IVF-FLAT row shards: split the coarse routing of hb_sharded_search over the ranks as well (rank r ranks ALL centroids,
 * exactly, for the queries [r ceil(nq/G), (r+1) ceil(nq/G)); one all-gather replicates the probe lists).  On by default for
 * indexes built by hb_sharded_ivf_build; every rank must use the same setting.  Same results either way. */
HB_API int hb_index_set_coarse_sharded(hb_index *index, int on);
/* search mode of THIS index: HB_MODE_EXACT / HB_MODE_FAST, or -1 to follow hb_set_mode (threads that want different
 * modes on different indexes do not race on the process default) */
HB_API int hb_index_set_mode(hb_index *index, int mode);
/* search-knn over all shards: same arguments and result layout as hb_search, ids are GLOBAL rows, identical results on
 * every rank.  Collective. */
HB_API int hb_sharded_search(hb_index *index, const void *queries, int qdtype, int64_t nq, int32_t k, int32_t param,
                             int64_t *out_ids, double *out_dist);
/* partition-vectors-kmeans (ivf_flat.clj:92-131) data-parallel: every rank assigns its rows and sums its members per
 * cluster, one all-reduce(sum) of the fp64 sums + counts per Lloyd round, centroids replicated.  seed_rows = nlist GLOBAL
 * row ids (the k-means++ result, the same array on every rank).  out_assign: this rank's rows.  Collective. */
HB_API int hb_sharded_kmeans(const void *rows, int64_t n_local, int32_t d, int dtype, int metric, int32_t nlist, int32_t iters,
                             const int64_t *seed_rows, int64_t first_global_row, double *out_centroids, int32_t *out_assign);
/* build-ivf-flat-index (ivf_flat.clj:137-211) over the row shards: hb_sharded_kmeans, then this rank's rows laid out as the
 * list-major slabs of the global partitioning (id base = first_global_row).  Collective. */
HB_API int hb_sharded_ivf_build(const void *rows, int64_t n_local, int32_t d, int dtype, int metric, int32_t nlist, int32_t iters,
                                const int64_t *seed_rows, int64_t first_global_row, hb_index **out);

/* ---- persistence ------------------------------------------------------------------------------- */
/* save-index / load-index (src/hnsw/api.clj:40-50, src/hnsw/helper/index_io.clj:10-80: EDN text of the HNSW graph
 * only; IVF-FLAT has no persistence in the reference).  The file is the device layout verbatim (header + tagged
 * sections: list-major slab, fp64 norms, centroids, list offsets, row ids, assignments / HNSW levels + CSR
 * adjacency), written atomically (path.tmp + rename); loading streams it into HBM through pinned staging without
 * re-clustering or re-computing norms, and the loaded index returns the same bits as the saved one.  String ids
 * stay with the host shim.  A missing / truncated / foreign file is HB_ERR_INVALID (the reference prints and
 * returns nil, index_io.clj:78-80). */
HB_API int hb_index_save(const hb_index *index, const char *path);
HB_API int hb_index_load(const char *path, hb_index **out);

/* ---- bookkeeping -------------------------------------------------------------------------------- */
HB_API int hb_index_info(const hb_index *index, hb_info *out);
HB_API int hb_index_free(hb_index *index);

#ifdef __cplusplus
}
#endif
#endif /* HNSWB200_H */
