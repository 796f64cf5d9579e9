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


class MicroBatcher:
    """Turns concurrent single-query searches into device batches.

    The reference's callers issue one (search-knn index query k) per thread — up to 50 threads on one shared index
    (src/hnsw/wip/31k-multithread-sb.clj:127-134, test/hnsw/core_test.clj:112-121).  On the device a lone query costs
    the same launch chain as ten thousand, so the calls that arrive while a batch is in flight (or within `max_wait_s`
    of the first waiter) are answered together by ONE hb_search; every caller gets exactly the result a lone call would
    have returned (the batch is a stack of independent queries).  search() is a drop-in for the per-query search-fn of
    parallel-search-futures."""

    def __init__(self, index, k, max_batch=4096, max_wait_s=0.0002, **opts):
        import threading

        self.index, self.k, self.opts = index, k, opts
        self.max_batch, self.max_wait_s = int(max_batch), float(max_wait_s)
        self._cv = threading.Condition()
        self._pending = []  # [query, slot] pairs; slot = [done event, result or exception]
        self._closed = False
        self.batches = 0
        self.served = 0
        self._worker = threading.Thread(target=self._run, name="hb-microbatcher", daemon=True)
        self._worker.start()

    def search(self, query):
        import threading

        slot = [threading.Event(), None]
        with self._cv:
            if self._closed:
                raise RuntimeError("MicroBatcher is closed")
            self._pending.append((query, slot))
            self._cv.notify_all()
        slot[0].wait()
        if isinstance(slot[1], BaseException):
            raise slot[1]
        return slot[1]

    __call__ = search

    def _run(self):
        import numpy as np

        while True:
            with self._cv:
                while not self._pending and not self._closed:
                    self._cv.wait()
                if not self._pending and self._closed:
                    return
                if len(self._pending) < self.max_batch and self.max_wait_s > 0:
                    deadline = time.perf_counter() + self.max_wait_s
                    while len(self._pending) < self.max_batch and not self._closed:
                        left = deadline - time.perf_counter()
                        if left <= 0:
                            break
                        self._cv.wait(left)
                batch, self._pending = self._pending[:self.max_batch], self._pending[self.max_batch:]
            try:
                res = api.search_batch_(self.index, np.stack([np.asarray(q) for q, _ in batch]), self.k, **self.opts)
            except BaseException as e:  # every waiter sees the failure
                res = [e] * len(batch)
            self.batches += 1
            self.served += len(batch)
            for (_, slot), r in zip(batch, res):
                slot[1] = r
                slot[0].set()

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
