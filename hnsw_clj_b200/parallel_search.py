"""Mirror of hnsw.helper.parallel-search (src/hnsw/helper/parallel_search.clj:15-49): the reference fans N
single-query searches out on a fixed thread pool and returns results in query order; the device path
answers the whole batch in one call, so num_threads only survives as an accepted (ignored) argument."""
from __future__ import annotations

import time

from . import api


def parallel_search_futures(index, queries, k, search_fn=None, num_threads=None, **opts):
    """(parallel-search-futures index queries k search-fn num-threads) -> results in query order."""
    if search_fn is not None and not getattr(search_fn, "_hb_batched", False) and search_fn not in (
            api.search, api.search_knn_):
        # a caller-supplied per-query function keeps the reference semantics
        return [search_fn(index, q, k) for q in queries]
    return api.search_batch_(index, queries, k, **opts)


def benchmark_parallel_search(index, queries, k, search_fn=None, num_threads=None, **opts):
    """(benchmark-parallel-search ...), :51-95 -> {:total-ms :avg-ms :qps}."""
    t0 = time.perf_counter()
    res = parallel_search_futures(index, queries, k, search_fn, num_threads, **opts)
    dt = (time.perf_counter() - t0) * 1e3
    n = max(len(res), 1)
    return {"total-ms": dt, "avg-ms": dt / n, "qps": n / (dt / 1e3) if dt > 0 else float("inf"), "results": res}
