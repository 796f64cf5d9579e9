// hb_hnsw.cuh — launcher of the batched HNSW search kernel (hb_hnsw.cu).
#pragma once
#include "hb_common.cuh"

namespace hb {

struct HnswSearchParams {
    // graph + vectors (hb_hnsw_create)
    const void *rows = nullptr;
    const double *row_norm = nullptr;  // cosine: sqrt(sum v^2) per row
    int d = 0;
    int max_level = 0;
    int entry = 0;
    const int64_t *const *adj_off = nullptr;  // [max_level+1] device pointers: CSR offsets over all n nodes
    const int32_t *const *adj_ids = nullptr;  // [max_level+1] device pointers: neighbour ids in iteration order
    // queries
    const void *queries = nullptr;
    const double *q_norm = nullptr;
    int nwork = 0;                       // queries to run
    const int32_t *work_list = nullptr;  // work item -> query index (NULL: identity)
    int32_t *next_work = nullptr;        // work counter (zeroed by the caller)
    int ef = 0, k = 0, epi = EPI_COS_GUARD;
    int pf_window = 6;  // 128-byte lines of each gathered row prefetched to L2 ahead of the chunk being staged
    // per-warp state
    int ef_cap = 0;         // slots of the `nearest` queue and of the entry list (>= ef + 1)
    int cand_cap_smem = 0;  // slots of the candidate queue in shared memory
    size_t warp_smem = 0;   // hnsw_warp_smem(ef_cap, cand_cap_smem)
    double *g_cand_d = nullptr;  // non-NULL: candidate queues in global memory, [warp slots][g_cand_cap]
    int32_t *g_cand_i = nullptr;
    int g_cand_cap = 0;
    uint32_t *visited = nullptr;  // [warp slots][vwords] all zero on entry and on exit
    int64_t vwords = 0;
    int32_t *vlist = nullptr;  // [warp slots][vcap] ids marked in the current layer
    int vcap = 0;
    // results
    int64_t *out_ids = nullptr;  // [nq][k]
    double *out_dist = nullptr;
    int32_t *n_overflow = nullptr;     // queries whose candidate queue did not fit ...
    int32_t *overflow_list = nullptr;  // ... and which they are
    unsigned long long *n_scored = nullptr;  // (query, row) pairs scored: the gather-dot count
};

int hnsw_warps_per_cta();
size_t hnsw_warp_smem(int ef_cap, int cand_cap_smem);
// grid CTAs of hnsw_warps_per_cta() warps; the per-warp buffers must cover grid * warps slots
void launch_hnsw_search(const HnswSearchParams &P, int rdtype, int qdtype, bool l2, int grid);

}  // namespace hb
