// hb_comm.cuh — the multi-GPU data plane of libhnswb200: one process per GPU (hb_comm_init), NCCL (loaded at run time)
// for the bulk collectives of the k-means build, and a fused exchange + merge kernel over NVLink peer memory for the
// search: every rank pushes its local top-k straight into its peers' windows, signals, waits for theirs and merges —
// one launch, no host round trip.  Reference scale-out model: independent sub-indexes searched in parallel, results
// concatenated, (sort-by :distance) + (take k) — src/hnsw/ann/partition/partitioned_hnsw.clj:149-196.
#pragma once
#include "hb_common.cuh"

namespace hb {

constexpr int kCommIdBytes = 128;  // ncclUniqueId
constexpr int kCommMaxRanks = 8;   // GPUs of one NVSwitch box

struct CommInfo {
    int nranks = 1, rank = 0;
    bool inited = false;
    bool p2p = false;  // peer windows mapped (cudaIpc over NVLink): the fused exchange kernel is available
};
const CommInfo &comm_info();

void comm_unique_id(void *out128);
// collective over all ranks; `device` is this process's GPU (hb_init)
void comm_init(const void *id128, int nranks, int rank, int device);
void comm_shutdown();

// in-place collectives on device buffers, enqueued on g_stream (no host synchronisation)
void comm_allreduce_sum_f64(double *buf, int64_t count);
void comm_allreduce_sum_i64(int64_t *buf, int64_t count);
void comm_allreduce_max_f64(double *buf, int64_t count);
void comm_broadcast_bytes(void *buf, int64_t bytes, int root);
void comm_allgather_bytes(const void *send, void *recv, int64_t bytes_per_rank);
void comm_barrier();  // all ranks' g_streams have reached this point (host returns after the local stream has)

// Global top-k from every rank's local top-k (device, [nq x k], ascending by (distance, position), ids already local row
// indices; id_base is added to every id >= 0).  out_* on the device, identical on every rank.  Ties between ranks fall to the
// lower rank = the lower global row for contiguous row shards (the stable sort of the concatenation in rank order).
// use_p2p: 1 = the fused push + signal + wait + merge kernel over the peer windows when there are peers, 2 = that kernel even
// with one rank (tests), 0 = pack -> ncclAllGather -> merge.
// timing (optional, host): [0] exchange (pack + all-gather, or 0 for the fused kernel), [1] merge / fused kernel — CUDA events.
struct ExchangeStats {
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
};
void comm_topk_exchange_merge(const double *loc_dist, const int64_t *loc_ids, int64_t id_base, int64_t nq, int k, double *out_dist,
                              int64_t *out_ids, int use_p2p, ExchangeStats *st);
// after the stream has been synchronised: throws if the fused kernel timed out waiting for a peer
void comm_check_exchange();

// centroids[c][:] = sums[c][:] / counts[c] where counts[c] > 0 (an empty cluster keeps its centroid, ivf_flat.clj:112-116)
void launch_divide_centroids(const double *sums, const int64_t *counts, int nlist, int d, double *cents);
// cents[j][:] = rows[seed_rows[j] - first_row][:] if this shard owns the seed row, else 0 (all-reduce(sum) replicates them)
void launch_init_centroids_sharded(const void *rows, int dtype, int d, const int64_t *seed_rows, int nlist, int64_t first_row,
                                   int64_t n_local, double *cents);

}  // namespace hb
