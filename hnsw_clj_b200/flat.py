"""Exact flat search: hnsw.bench/compute-exact-knn (src/hnsw/bench.clj:72-84) and calc-recall (:86-92)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as hb
from .index import DeviceIndex, metric_code, new_handle, results_to_maps, split_data


class FlatIndex(DeviceIndex):
    """Rows resident on the device + their norms; search = every distance, stable sort, take k."""

    def __init__(self, data, distance_fn="cosine"):
        ids, rows = split_data(data)
        self.metric = metric_code(distance_fn)
        h = new_handle()
        n, d = (rows.shape[0], rows.shape[1]) if rows.shape[0] else (0, max(int(rows.shape[1]) if rows.ndim == 2 else 1, 1))
        hb.check(hb.lib().hb_flat_create(hb.ptr(rows) if n else None, n, d, hb.dtype_code(rows) if n else hb.F32,
                                         self.metric, C.byref(h)))
        super().__init__(h.value, ids)

    def search_knn(self, query, k):
        ids, dist = self.search_raw(query, k)
        return results_to_maps(ids, dist, self.ids)[0]

    def search_batch(self, queries, k):
        ids, dist = self.search_raw(queries, k)
        return results_to_maps(ids, dist, self.ids)


def compute_exact_knn(vectors, query, k):
    """(compute-exact-knn vectors query k): vectors = seq of [id v]; -> [{:id :distance} ...]."""
    with FlatIndex(vectors, distance_fn="cosine") as ix:
        return ix.search_knn(query, k)


def calc_recall(approx, exact) -> float:
    """(calc-recall approx exact), src/hnsw/bench.clj:86-92: |approx ∩ exact| / |exact|."""
    a = {r["id"] for r in approx}
    e = {r["id"] for r in exact}
    return len(a & e) / len(e) if e else 1.0


def recall_at_k(approx_ids: np.ndarray, exact_ids: np.ndarray) -> float:
    """Mean calc-recall over a batch of raw id arrays (padding -1 ignored), cf. measure-recall (:124-132)."""
    tot = 0.0
    for a, e in zip(np.asarray(approx_ids), np.asarray(exact_ids)):
        es = set(int(x) for x in e if x >= 0)
        tot += (len(es & set(int(x) for x in a if x >= 0)) / len(es)) if es else 1.0
    return tot / max(len(approx_ids), 1)
