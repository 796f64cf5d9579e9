"""hnsw_clj_b200 — B200-native distance core behind hnsw-clj's build-index / search-knn / search-batch* surface.

Python host side of the C ABI (libhnswb200.so); module names follow the reference's namespaces:
  simd_optimized  <- hnsw.simd-optimized          (pairwise distances, norms, top-k)
  simd            <- hnsw.simd                    (float[] Vector-API variants: fp32 lanes, fp64 accumulation)
  pcaf            <- hnsw.ann.dimreduct.pcaf      (random projection + low-dim scan + full-dim re-rank)
  hybrid_lsh      <- hnsw.ann.hash.hybrid-lsh     (hash tables on the host, bucket scans on the device)
  flat            <- hnsw.bench/compute-exact-knn (exact flat search)
  ivf_flat        <- hnsw.ann.partition.ivf-flat  (build-index / search-knn / index-info)
  lightning       <- hnsw.ann.partition.lightning (k-means++-seeded partitions, percentage probing; the IVF-FLAT scan)
  ultra_fast      <- hnsw.ultra-fast              (HNSW neighbour-candidate scoring on an uploaded graph)
  index_io        <- hnsw.helper.index-io         (save-index / load-index of the device layout)
  data_loader     <- hnsw.helper.data-loader      (embeddings JSON -> ids + one pinned fp32 / fp64 matrix)
  api             <- hnsw.api + hnsw.api.protocol (index / search, ANNIndex + BatchSearchIndex)
  parallel_search <- hnsw.helper.parallel-search  (batch fan-out = one device call; MicroBatcher for concurrent single-query callers)
  sharded         <- hnsw.ann.partition.partitioned-hnsw's scale-out model: row shards, one process per GPU, the
                     library's own data plane (hb_comm_* / hb_sharded_*: NVLink peer windows + NCCL)
"""
from . import _lib
from ._lib import HbError, HbInvalid, launch_count  # noqa: F401

__all__ = ["_lib", "HbError", "HbInvalid", "launch_count"]
