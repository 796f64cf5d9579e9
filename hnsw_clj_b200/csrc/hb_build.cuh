// hb_build.cuh — launchers for the IVF build / bookkeeping kernels (hb_build.cu).
#pragma once
#include "hb_common.cuh"

namespace hb {

// centroids[c][:] = (double) rows[seed_rows[c]][:]   (centroids start as data rows, ivf_flat.clj:40,58)
void launch_init_centroids(const void *rows, int dtype, int d, const int64_t *seed_rows, int nlist, double *cents);

// compute-centroid (ivf_flat.clj:66-77) per cluster over its members in row order; empty keeps the old
// centroid (:112-116).  list_off/list_rows: CSR of rows per cluster, rows ascending.
// If out_sums != NULL the raw sums/counts are stored instead (multi-GPU partials) and cents is untouched.
void launch_update_centroids(const void *rows, int dtype, int d, const int64_t *list_off, const int64_t *list_rows,
                             int nlist, double *cents, double *out_sums, int64_t *out_counts);

// k-means++ step (ivf_flat.clj:43-58).  *pick is the row chosen last; mind is updated with
// min(mind, dist(x_i, x_pick)); then S = sum mind^2 in row order, r = u*S, and *pick becomes the first i
// with cum_i + mind_i^2 >= r (clamped to n-1).
struct KppParams {
    const void *rows = nullptr;
    int dtype = HB_F32;
    int64_t n = 0;
    int d = 0;
    const double *row_norm = nullptr;
    bool l2 = false;
    bool linear = false;        // weights d_i instead of d_i^2 (Lightning, lightning.clj:100-106)
    double *mind = nullptr;     // [n]
    double *cum = nullptr;      // [n] scratch
    double *total = nullptr;    // [1] scratch
    int64_t *pick = nullptr;    // [1] in: newest centroid row, out: next
    const double *u = nullptr;  // [1] the nextDouble() of this step (device)
    int64_t *out_seed = nullptr;  // where to record the new pick
};
void launch_kpp_step(const KppParams &P);

// The same step at scale and bit-identical (hb_kpp.cu): rows the triangle inequality excludes are skipped, the ordered
// sum runs as integer adds per chunk of kKppChunk rows.  State per row: mind, near (pick order of the seed that gave mind),
// theta (the row's bound towards that seed); per chunk: the scratch of the ordered sum.
constexpr int kKppChunk = 512;
struct KppScaleParams {
    const void *rows = nullptr;
    int dtype = HB_F32;
    int64_t n = 0;
    int d = 0;
    const double *row_norm = nullptr;
    bool l2 = false, linear = false;
    double *mind = nullptr;    // [n]
    int32_t *near = nullptr;   // [n]
    float *theta = nullptr;    // [n]
    const int64_t *seeds = nullptr;  // [nlist] rows picked so far (device); seeds[t - 1] is the newest
    int64_t *seeds_rw = nullptr;     // the same array: the new pick goes to seeds_rw[t]
    int t = 0;                       // seeds picked so far
    double *bound = nullptr;         // [nlist] scratch
    double *c_approx = nullptr, *c_prefix = nullptr, *c_start = nullptr;  // [nchunks], [nchunks], [nchunks + 1]
    int *c_exp = nullptr;            // [nchunks]
    long long *c_q = nullptr;        // [nchunks]
    const double *u = nullptr;       // [1] the nextDouble() of this step (device)
    double *total = nullptr;         // [1]
    int64_t *pick = nullptr;         // [1]
    unsigned long long *n_scored = nullptr, *n_walked = nullptr;  // optional counters: rows scored / chunks walked with fp64 adds
};
void launch_kpp_init_state(int64_t n, double *mind, int32_t *near, float *theta);
void launch_kpp_scale_step(const KppScaleParams &P);
// S, r = u * S and the pick for the weights already in P.mind (the second half of a step on its own)
void launch_kpp_sum_pick(const KppScaleParams &P);

// slab[j][:] = rows[list_rows[j]][:], slab_norm[j] = norm[list_rows[j]]
void launch_gather_rows(const void *rows, int dtype, int d, const int64_t *list_rows, int64_t n, void *slab,
                        const double *norm, double *slab_norm);

// CSR of rows per cluster from assignments, rows ascending inside a cluster (ivf_flat.clj:126-129).
// Returns through device arrays; uses cub radix sort (stable).
struct DevBuf;
void build_lists(const int32_t *assign, int64_t n, int nlist, int64_t *list_off /*[nlist+1]*/,
                 int64_t *list_rows /*[n]*/, DevBuf &tmp);

// IVF search bookkeeping: from probes [nq x nprobe] (list ids, -1 = none) derive
//   pair_out [np+1] exclusive scan of probed list lengths (np = nq*nprobe),
//   qsel [np] pair ids grouped by list, lq_off [nlist+1], tile_prefix [nlist+1].
void ivf_plan(const int64_t *probe_pos /*[np] from select, -1 padded*/, int64_t np, int nlist,
              const int64_t *list_off, int32_t *probes, int64_t *pair_out, int32_t *qsel, int64_t *lq_off,
              int64_t *tile_prefix, int tile_rows, int tile_q, DevBuf &tmp);

// The FAST list scan's variant (three small kernels, no sort): pairs grouped by list in no particular order inside a list;
// pair_out (optional) [nq x npq]: offset of each probed list inside ITS query's concatenation; unit_prefix = exclusive
// prefix of ceil(pairs of the list / tile_q) over the lists that hold rows.
// probe_pos: [nq x pos_stride], the first npq entries of every row are planned (npq = 1: the nearest list of every query).
void ivf_plan_fast(const int64_t *probe_pos, int pos_stride, int64_t nq, int npq, int nlist, const int64_t *list_off, int32_t *probes,
                   int64_t *pair_out /*NULL: not wanted*/, int32_t *qsel, int64_t *lq_off, int64_t *unit_prefix, int tile_q,
                   DevBuf &tmp);

// pos (within the query's concatenated probed lists) -> row id
void launch_ivf_resolve(const int64_t *pos, int64_t nq, int k, int nprobe, const int32_t *probes,
                        const int64_t *pair_out, const int64_t *list_off, const int64_t *list_rows,
                        int64_t *out_ids, int32_t *ok = nullptr /* optional: ok[q] &= ok_other[q] */,
                        const int32_t *ok_other = nullptr);
// flat: row id = pos + base (pos -1 stays -1) for the `run` entries at ids[q*stride ..] of every query q
void launch_offset_ids(int64_t *ids, int64_t nq, int64_t run, int64_t stride, int64_t base);
// [nparts][nq][k] -> [nq][nparts][k]
void launch_parts_to_query_major(const double *dist, const int64_t *ids, int nparts, int64_t nq, int k, double *out_dist,
                                 int64_t *out_ids);
void launch_fill_f64(double *p, int64_t n, double v);
// merge: id = ids_in[q][pos]
void launch_lookup_ids(const int64_t *pos, int64_t nq, int k, const int64_t *ids_in, int64_t stride, int64_t *out_ids);
// probes int32 [nq x nprobe] from select positions
void launch_pos_to_i32(const int64_t *pos, int64_t n, int32_t *out);

}  // namespace hb
