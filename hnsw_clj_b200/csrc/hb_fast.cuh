// hb_fast.cuh — HB_MODE_FAST: tensor-core candidate pass with a proof of exactness.
//
// The reference ranks by fp64 distances.  FAST mode reproduces those results bit for
// bit in three steps:
//   1. candidate pass (hb_tc.cu): every (query, row) score of the probed lists is computed on the 5th-gen
//      tensor cores as an EXACT integer dot product of block-fixed-point digits (tcgen05.mma kind::i8,
//      int32 accumulators in TMEM).  Rows and queries are quantised to NS signed 8-bit digits per element
//      with one scale per vector, so the only error is the quantisation itself, which is bounded a priori
//      per query (eps_q, below) — no assumption about the tensor core's floating-point accumulation.
//   2. the kk best candidates per query are re-scored with the reference's sequential fp64 arithmetic
//      (gather_score_kernel) and ordered by the reference's (distance, position) rule;
//   3. a query is accepted only if its exact k-th best similarity beats every rejected row's upper bound
//      (approximate score + eps_q); otherwise it is recomputed by the exact fp64 path.  Accepted results are
//      therefore identical (ids and distance bits) to the exact path and to the oracle.
//
// Layout of the digit images (what the kernel's bulk copies land in shared memory, already in the canonical
// K-major SWIZZLE_128B form of the UMMA shared-memory descriptor):
//   image[tile][kb][slice] = 128 vectors x 128 bytes; byte (r, c) at (r/8)*1024 + (r%8)*128 + (((c/16) ^ (r%8))*16) + c%16
// tile = 128 consecutive slab rows of one list (B side) or the 128 query slots of one unit (A side);
// kb = block of 128 dimensions; slice 0 is the most significant digit.
#pragma once
#include "hb_common.cuh"

namespace hb {

constexpr int kFastTile = 128;          // rows per B tile = query slots per unit = UMMA M = UMMA N
constexpr int kFastKB = 128;            // dimensions per k-block (one 128-byte swizzle row of int8)
constexpr int kFastImg = kFastTile * kFastKB;  // bytes of one (tile, kb, slice) image
constexpr int kNarrowSlots = 32;        // a unit with at most this many selections runs with the rows on the M side (tc_narrow_kernel)
                                        // (16 or 32.  Measured at configs[1] with 16: the query images halve, but 12 % of the units — 24 % of
                                        //  the items — fall to the dense kernel: step 1.628 vs 1.632 ms, 32 queries on flat 1M 673 vs 575 us)

// quantisation range: |m| <= qmax so that every digit fits int8 after the balanced split
__host__ __device__ constexpr double fast_qmax(int ns) { return ns == 2 ? 32000.0 : 8000000.0; }
// |x/u - m| <= kFastRound (division by reciprocal in fp64 + round to nearest)
constexpr double kFastRound = 0.5000001;

struct FastSide {           // quantised vectors of one side, device pointers
    int8_t *img = nullptr;  // B side: [ntiles][kbn][ns][kFastImg]
    float *rs = nullptr;    // per slab position (padded to tiles): score = S * rs + ro
    float *ro = nullptr;    // 0 for a real row, -inf for padding
};

// ---- preparation kernels (hb_fastprep.cu) ---------------------------------------------------------------
// B side: quantise the rows of every list into tile images.  tile_off[l] = first tile of list l.
// norm: fp64 norms per slab row (cosine) or NULL (inner product: rs = scale only).
// stats[0] = max over rows of u_r*w_r, stats[1] = max of ||r||_1*w_r  (w_r = 1/norm or 1): for eps_q;
// stats[2] > 0 if some row has a zero / non-finite norm (the index is then not eligible for FAST mode).
// The caller zeroes img / rs / stats and fills ro with -inf (padding) beforehand.
void launch_fill_f32(float *p, int64_t n, float v);
void launch_quant_rows(const void *rows, int dtype, int d, int kbn, int ns, int nlist, const int64_t *list_off,
                       const int64_t *tile_off, const double *norm, int8_t *img, float *rs, float *ro, float *stats);
// tile_off[l] = exclusive prefix of ceil(len_l / 128)   (device, [nlist+1]); single block
void launch_tile_offsets(const int64_t *list_off, int nlist, int64_t *tile_off);

// A side: digits of every query [nq][ns][kbn*128] + per-query scale u_q, L1 norm
void launch_quant_queries(const void *queries, int qdtype, int64_t nq, int d, int kbn, int ns, int8_t *dig, double *qu,
                          double *ql1);

// Units: unit = (list, block of <= 128 selections of that list).  From lq_off/unit_prefix (= exclusive prefix of
// ceil(lq_len/128) per list) derive per unit: list id, first selection, and the exclusive prefix of row tiles
// (items) with at most `tile_limit` leading tiles per unit (tile_limit <= 0: all).
struct UnitPlan {
    int32_t *unit_list = nullptr;   // [nunits]
    int32_t *unit_sel0 = nullptr;   // [nunits] first selection
    int32_t *unit_nsel = nullptr;   // [nunits] selections (<= 128)
    int32_t *unit_ntile = nullptr;  // [nunits+1] scratch: row tiles per unit
    int32_t *unit_item0 = nullptr;  // [nunits+1]
    int32_t *unit_tile0 = nullptr;  // [nunits] first row tile of the unit's items (units whose tiles [0, skip) were already scored start at skip)
    int32_t *unit_item0n = nullptr; // optional [nunits+1]: when set, the units with <= kNarrowSlots selections are counted
                                    // here (tc_narrow_kernel) and unit_item0 covers only the others
    int32_t *slot_query = nullptr;  // [nunits*128] query index or -1
    int32_t *slot_rel0 = nullptr;   // [nunits*128] position of the (query, list) segment inside the query's concatenation
};
// nunits = unit slots (>= nunits_real; a multiple of `interleave` when interleave > 0, see unit_plan_kernel);
// nunits_real < 0: read the real count from unit_prefix[nlist] on the device (nunits is then an upper bound)
// skip_tiles > 0 (IVF main pass after a sample of the first skip_tiles tiles of every query's nearest list, whose candidates
// were kept): a unit all of whose selections are probe rank 0 (pair % pair_div == 0) starts at tile skip_tiles.
void launch_unit_plan(int nlist, const int64_t *lq_off, const int64_t *unit_prefix, const int64_t *tile_off, int nunits,
                      int nunits_real, int interleave, int tile_limit, int tile_div, int tile_start, const int32_t *qsel,
                      const int64_t *pair_out, int pair_div, const int32_t *pair_query, UnitPlan U, int skip_tiles = 0,
                      unsigned long long *stats = nullptr /* optional [6] counters of what the pass covers */);
// gathers the digits of each unit's queries into A images [nunits][kbn][ns][kFastImg]
// unit_nsel (optional): units with <= 64 selections get only slots 0..63 packed (the M = 64 candidate pass reads no more)
// narrow: units with <= kNarrowSlots selections get only slots 0..31 (tc_narrow_kernel's B operand)
void launch_pack_units(const int8_t *dig, int kbn, int ns, int nunits, const int32_t *slot_query, int8_t *aimg,
                       const int32_t *unit_nsel = nullptr, bool narrow = false);

// ---- the tensor-core pass (hb_tc.cu) -------------------------------------------------------------------
enum FastMode { FAST_EMIT = 1, FAST_DUMP = 2 };
struct TcParams {
    const int8_t *aimg = nullptr;
    const int8_t *bimg = nullptr;
    int kbn = 0;
    int nunits = 0;
    const int32_t *unit_list = nullptr;
    const int32_t *unit_item0 = nullptr;
    const int32_t *unit_item0n = nullptr;  // non-NULL: item prefix of the narrow units (see UnitPlan); EMIT only
    const int32_t *unit_nsel_all = nullptr;  // [nunits] selections per unit (always set with unit_item0n)
    const int32_t *unit_tile0 = nullptr;     // [nunits] first row tile of each unit's items (NULL: tile_start)
    // > 0: the first skip_tiles tiles of every query's NEAREST list (slots with rel0 == 0) were scored by the sample pass and
    // their candidates kept: such slots emit nothing for tiles below skip_tiles
    int skip_tiles = 0;
    // upper bound of the pass's items (unit x row tile) when the host knows it (flat scans: units x tiles of the level); sizes
    // the grid.  0: units x 8 (IVF plans live on the device)
    int64_t items_hint = 0;
    const int64_t *tile_off = nullptr;  // first B tile of each list
    const int64_t *list_off = nullptr;  // first slab row of each list
    const int32_t *slot_query = nullptr;
    const int32_t *slot_rel0 = nullptr;
    // optional [nunits] selections per unit: EMIT passes run a unit of <= 64 selections with M = 64 (half the A operand)
    const int32_t *unit_nsel = nullptr;
    const float *rs = nullptr;
    const float *ro = nullptr;
    float *thr = nullptr;  // per query: rows scoring below it cannot matter (see the epilogue); raised with atomic max
    const float *margin = nullptr;  // per query, in score units: > 2 eps_q (query_bounds_kernel)
    int k = 0;                      // results wanted per query
    int kk = 64;                    // candidates re-scored per query (64 or 128): thr keeps >= kk emitted rows at or above it
    int tile_stride = 1;   // item j of a unit is row tile tile_start + j*tile_stride of its list (the sample pass of a flat
    int tile_start = 0;    // scan takes every 8th tile, a level of a long flat scan a range of tiles)
    long long *timing = nullptr;  // debug: per CTA {mma warp total, wait accumulators free, wait operands, items, epilogue wait, phase A} clocks
    // EMIT
    int cap = 0;
    int32_t *cnt = nullptr;      // per query
    double *cand_negv = nullptr; // [nq][cap]  -score (ascending = best first)
    int32_t *cand_rel = nullptr; // [nq][cap]  position in the query's concatenated probed lists
    int32_t *cand_pos = nullptr; // [nq][cap]  slab row
    // DUMP
    float *dump = nullptr;       // [items][128 slots][128 rows]
};
void launch_tc_pass(const TcParams &P, int ns, int mode);

// ---- final ordering + proof (hb_fastprep.cu) ------------------------------------------------------------
// Per query: the kk selected candidates (sel_pos = index into the query's cand arrays, -1 unused) with their
// exact fp64 distances -> top-k by (distance, rel) + the acceptance test.  One CTA per query.
struct FinalParams {
    int64_t nq = 0;
    int k = 0, kk = 0, cap = 0;
    const int64_t *sel_pos = nullptr;   // [nq][kk]
    const double *sel_negv = nullptr;   // [nq][kk] approximate -score of the selected
    const double *exact = nullptr;      // [nq][kk] exact distances
    const int32_t *pair_row = nullptr;  // [nq][kk] row that was re-scored, -1: not re-scored
    const int32_t *cand_rel = nullptr;
    const int32_t *cnt = nullptr;
    const float *thr = nullptr;
    const double *q_scale = nullptr;    // per query: similarity = score * q_scale
    const double *q_eps = nullptr;      // per query bound on |similarity - approximate similarity|
    int metric = HB_COSINE;
    int64_t *out_rel = nullptr;         // [nq][k] rel of the winners (-1 unused)
    double *out_dist = nullptr;         // [nq][k]
    int32_t *out_ok = nullptr;          // [nq] 1 = proven exact
    double *out_simub = nullptr;        // optional [nq][k]: upper bound of each winner's similarity (probe pruning)
    // non-NULL (IVF list scan over an approximately ordered probe list): first slab row of each list; a distance tie
    // between rows of different lists among the k + 1 best fails the query (the reference breaks it by probe rank)
    const int64_t *tie_list_off = nullptr;
    int tie_nlist = 0;
};
void launch_fast_final(const FinalParams &P);

// per query: q_scale = u_q / ||q|| (cosine) or u_q (ip); q_eps from the index stats
void launch_query_bounds(const double *qu, const double *ql1, const double *qnorm, int64_t nq, int ns, int d, int metric,
                         const float *stats, double *q_scale, double *q_eps, float *q_margin, float *thr_init = nullptr /* [nq] <- -inf */,
                         int32_t *cnt_init = nullptr /* [nq] <- 0 */);
// pairs for the exact re-score: per selected slot the row or -1 (slot_row), exact = +inf where not re-scored (set_only:
// slot_row -2 and exact -inf for a candidate that is in the top-k set without a re-score), and the
// dense list of wanted (query, row, slot) triples with its length in *total
void launch_rescore_pairs(const int64_t *sel_pos, const double *sel_negv, const int32_t *cand_pos, int64_t nq, int kk, int cap,
                          int k, const float *margin, int32_t *slot_row, double *exact, int32_t *total, int32_t *pair_query,
                          int32_t *pair_row, int32_t *pair_slot, bool set_only = false);
// kk best of the min(cnt, cap) candidates of every query (ascending -score); unused slots: pos -1, +inf
void launch_cand_select(const double *cand_negv, const int32_t *cnt, int64_t nq, int kk, int cap, double *sel_negv,
                        int64_t *sel_pos);
// short flat scans (<= 2048 rows): from the dumped score matrix [unit][tile][slot][row] of a FAST_DUMP pass, per query the
// rows within `margin` of its k-th best score become its candidate list (sorted; sel_* filled, cnt = their number, or
// cap + 1 if more than kk) and thr the best rejected score
void launch_dense_select(const float *dump, int ntiles, int64_t nq, int nrows, int k, int kk, int cap, const float *margin,
                         double *cand_negv, int32_t *cand_rel, int32_t *cand_pos, int32_t *cnt, float *thr, double *sel_negv,
                         int64_t *sel_pos);
// the same selection, then only the kk best stay in the query's candidate list (in place, best first, cnt = min(cnt, kk))
// and its threshold rises to the kk-th best / the k-th best - margin: the step between two levels of a long flat scan
void launch_cand_compact(double *cand_negv, int32_t *cand_rel, int32_t *cand_pos, int32_t *cnt, int64_t nq, int kk, int cap, int k,
                         const float *margin, float *thr, double *sel_negv, int64_t *sel_pos);
// exact fp64 distance of (query, row) pairs in the reference's summation order (cosine / -dot epilogues, no L2)
void launch_rescore(const void *rows, int rdtype, const double *row_norm, const double *queries64, bool q_f32_repr,
                    const double *q_norm, int d, const int32_t *pair_query, const int32_t *pair_row, const int32_t *pair_slot,
                    const int32_t *total, int64_t max_pairs, int epi, double *out, bool sparse = false);
// (sparse: few pairs per query and fp64 rows — every lane stages its own query chunk, see rescore_kernel)
// fp64 copy of the queries for the re-score (returns `queries` itself when they already are fp64)
const double *launch_widen_queries(const void *queries, int qdtype, int64_t count, double *buf);
// fallback plumbing: dst[i] = src[idx[i]] (rows of row_bytes, multiple of 4) and dst[idx[i]] = src[i] (rows of k 8-byte words)
void launch_gather_bytes(const void *src, const int32_t *idx, int64_t n, int64_t row_bytes, void *dst);
// idx[0 .. count) = the queries whose proof failed (ok[q] == 0), ascending; idx[nq] = count.  idx holds nq + 1 entries.
void launch_compact_failed(const int32_t *ok, int64_t nq, int32_t *idx);
void launch_scatter_rows64(const void *src, const int32_t *idx, int64_t n, int k, void *dst);
void launch_i32_to_i64(const int32_t *in, int64_t n, int64_t *out);
// first[q] = pos[q*stride]  (probe rank 0 of every query)
void launch_first_column(const int64_t *pos, int64_t nq, int stride, int64_t *first);
// thr[q] = max(thr[q], kk-th best sample candidate) where the sample pass collected kk..cap candidates
void launch_thr_from_sample(const double *sel_negv, const int32_t *cnt, int64_t nq, int kk, int cap, int k, const float *margin,
                            float *thr);
// Exact pruning of probed lists (triangle inequality on angles, see hb_fastprep.cu): radius[l] = largest angle between a
// row of list l and its centroid; probe_pos[q][p >= 1] = -1 where no row of the list can reach the query's threshold
void launch_list_radius(const void *slab, int dtype, const double *slab_norm, const double *cents, const double *cent_norm,
                        const int64_t *list_off, int nlist, int64_t n, int d, double *radius);
void launch_prune_probes(int64_t *probe_pos, const double *sim_ub, const double *radius, const float *thr, const double *q_scale,
                         const double *q_eps, int64_t nq, int np, const int64_t *list_off, unsigned long long *pruned /*[2]*/);
// profiling counters accumulated on the device (read by hb_get_stat), and a 4-word plan written without host memory
void launch_accumulate_u64(unsigned long long *acc, const unsigned long long *src, int n);
void launch_tc_cover(const int64_t *unit_prefix, const int64_t *tile_off, const int64_t *lq_off /*NULL: M = 64 units off*/,
                     int nlist, unsigned long long *acc /*[4]: units, items, row tiles, M = 64 units*/);
void launch_set_i64x4(int64_t *p, int64_t a, int64_t b, int64_t c, int64_t d);
// sim[i] = 1 - dist[i] (+ a rounding allowance): the similarity bound of an exactly ranked probe list
void launch_sim_from_dist(const double *dist, int64_t n, double *sim);
// the all-gathered probe lists of the query-split coarse routing (hb_api.cu: sharded_coarse) -> ppos, simub = 1 - dist (+ allowance)
void launch_unpack_probe_blocks(const void *gathered, int nranks, int64_t per, int64_t nq, int np, int64_t *ppos, double *simub);
// ok[q] &= other[q]
void launch_and_flags(int32_t *ok, const int32_t *other, int64_t nq);

}  // namespace hb
