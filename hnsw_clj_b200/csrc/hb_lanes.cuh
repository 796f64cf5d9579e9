// hb_lanes.cuh — launchers of the fp32-lane distance kernels (hb_lanes.cu; SURVEY §8 a4, src/hnsw/simd.clj:18-115).
#pragma once
#include "hb_common.cuh"

namespace hb {

// out[i * out_stride + j] = metric(a_i, b_j) in the reference's Vector-API arithmetic with `lanes` floats per chunk;
// HB_COSINE = cosine-distance-simd-optimized (simd.clj:73-115), HB_L2 = euclidean-distance-simd-optimized (:45-71),
// HB_IP = dot-product-simd-optimized (:18-43).  All pointers on the device.
void launch_lanes_pairwise(const float *a, int64_t na, const float *b, int64_t nb, int d, int metric, int lanes, double *out,
                           int64_t out_stride);
// project-vector-simd (pcaf.clj:48-81) for n rows: out[r * target_dim + t] = (float) dot(matrix_t, row_r)
void launch_lanes_project(const float *matrix, int64_t target_dim, const float *rows, int64_t n, int d, int lanes, float *out);
// out[q * c + s] = cosine(query q, row cand[q * c + s]); +inf for cand < 0  (phase 2 of search-pcaf-parallel, :236-243)
void launch_lanes_gather(const float *queries, const float *rows, int d, int lanes, const int64_t *cand, int64_t nq, int c,
                         double *out);

}  // namespace hb
