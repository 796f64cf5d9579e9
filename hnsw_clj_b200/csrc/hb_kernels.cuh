// hb_kernels.cuh — launcher declarations for the sm_100a kernels of libhnswb200.
#pragma once
#include "hb_common.cuh"

namespace hb {

// -------------------------------------------------------------------------------------------
// Exact pair scan (hb_pairscan.cu): a tiled fp64 "GEMM" whose every (row, query) accumulator is
// the reference's sequential index-order sum.  Work is a set of lists; list l pairs the slab rows
// [list_off[l], list_off[l+1]) with the query selections qsel[lq_off[l] .. lq_off[l+1]).
// -------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;  // TR
constexpr int kTileQ = 64;      // TQ

struct ScanParams {
    const void *rows = nullptr;        // slab [*, d], row dtype
    const double *row_norm = nullptr;  // sqrt(sum v^2) per slab row (cosine epilogues)
    const void *queries = nullptr;     // [*, d], query dtype
    const double *q_norm = nullptr;    // per query
    int d = 0;
    int nlist = 0;
    const int64_t *list_off = nullptr;     // [nlist+1] slab rows of each list
    const int64_t *lq_off = nullptr;       // [nlist+1] selections of each list
    const int32_t *qsel = nullptr;         // selection -> pair id        (NULL: identity)
    const int32_t *pair_query = nullptr;   // pair id -> query index      (NULL: identity)
    int32_t pair_div = 0;                  // if > 0 and pair_query NULL: query = pair / pair_div
    const int64_t *pair_out = nullptr;     // pair id -> base offset in out (NULL: pair * out_stride)
    int64_t out_stride = 0;
    const int64_t *tile_prefix = nullptr;  // [nlist+1] exclusive prefix of tiles per list
    double *out = nullptr;                 // distances
    int epi = EPI_COS;
};
// max_sel: the caller's upper bound on the selections of any one list (0: unknown).  Up to kSmallScanQ the scan runs as the
// HBM-bound thread-per-row kernel (smallscan_kernel) instead of 128 x 64 fp64 tiles; the results are the same bits.
constexpr int kSmallScanQ = 8;
// batch_nq > 0 (with many lists): P.queries holds exactly that many queries and every selection refers to one of them
void launch_pairscan(const ScanParams &P, int rdtype, int qdtype, bool l2, int max_sel = 0, int batch_nq = 0);
// hb_rowstream.cu: the same small-batch scan for ONE list (flat search, coarse routing) fed by per-row bulk copies into
// shared-memory stages; launch_pairscan picks it when it applies.  false: does not fit, nothing was launched.
bool launch_rowstream(const ScanParams &P, int rdtype, int qdtype, bool l2, int max_sel);
// ... and for the probed lists of a small-batch IVF scan (many lists; the batch's nq <= kSmallScanQ queries are rows
// 0..nq-1 of P.queries, every selection refers to one of them)
bool launch_liststream(const ScanParams &P, int rdtype, int qdtype, bool l2, int nq);
void set_rowstream_option(const char *name, int value);  // "stream_seg" | "stream_stages" | "stream_warps"
extern int g_use_rowstream;                              // hb_set_option("rowstream", 0/1)

// k-means assignment (hb_pairscan.cu): argmin over centroids of the same exact distances, strict <,
// lowest index wins (ivf_flat.clj:79-90).
struct AssignParams {
    const void *rows = nullptr;  // points [n, d]
    const double *row_norm = nullptr;
    int64_t n = 0;
    int d = 0;
    const double *cents = nullptr;  // [nlist, d] fp64
    const double *cent_norm = nullptr;
    int nlist = 0;
    int epi = EPI_COS_GUARD;
    int32_t *out_assign = nullptr;
    double *out_best = nullptr;  // optional: the winning distance
};
void launch_assign(const AssignParams &P, int rdtype, bool l2);

// sqrt(sum v^2) per row, sequential fp64 (ivf_flat.clj:171-177)
void launch_row_norms(const void *rows, int dtype, int64_t n, int d, double *out);

// out[p] = epi(sum over the row pair_row[p] and query pair_query[p]) — one thread per pair,
// sequential (ultra_fast.clj:192).
void launch_gather_score(const void *rows, int rdtype, const double *row_norm, const void *queries, int qdtype,
                         const double *q_norm, int d, const int32_t *pair_query, const int32_t *pair_row,
                         int64_t npairs, bool l2, int epi, double *out);

// -------------------------------------------------------------------------------------------
// Segmented top-k selection (hb_select.cu).  Segment s covers vals[seg_begin(s) .. +seg_len(s));
// ordering is (value by Double/compare, position) — position order is the reference's stable-sort
// order for every caller.  Writes k (key, position-in-segment) pairs per segment; unused slots
// have pos -1 / +inf.
// -------------------------------------------------------------------------------------------
struct SelectParams {
    const double *vals = nullptr;
    int64_t nseg = 0;                  // segments (one per query)
    const int64_t *seg_off = nullptr;  // segment s spans [seg_off[s*seg_off_stride], seg_off[(s+1)*seg_off_stride])
    int64_t seg_off_stride = 1;
    int64_t seg_stride = 0;            // if seg_off == NULL: segment s spans [s*seg_stride, +seg_len_const)
    int64_t seg_len_const = 0;
    // optional split of every segment into nsub sub-ranges of sub_len candidates, one CTA each (few queries,
    // long segments); sub-range j of segment s writes slot s*out_seg_stride + out_slot_base + j
    int nsub = 1;
    int64_t sub_len = 0;
    int64_t out_seg_stride = 1;
    int64_t out_slot_base = 0;
    int k = 0;
    double *out_val = nullptr;   // [slots, k]
    int64_t *out_pos = nullptr;  // [slots, k] position within the segment (-1: unused)
};
void launch_select(const SelectParams &P);
int select_small_max();  // longest (sub-)range the shared-memory selection (select_small_kernel) takes


}  // namespace hb
