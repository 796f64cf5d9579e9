"""Mirror of the search side of hnsw.ultra-fast (src/hnsw/ultra_fast.clj) / hnsw.wip.ultra-optimized:
the graph is built on the host (graph mutation is out of scope) and uploaded; the device
does the neighbour-candidate scoring (:185-204) and the batched traversal."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as hb
from .index import DeviceIndex, metric_code, new_handle, results_to_maps, split_data


class HnswIndex(DeviceIndex):
    def __init__(self, data, levels, entry_point, adjacency, distance_fn="cosine"):
        """adjacency: list over levels of (offsets int64 [n+1], neighbour ids int32) in the iteration order the
        search follows (UltraNode neighbours, ultra_fast.clj:99-102)."""
        ids, rows = split_data(data)
        n, d = rows.shape
        lv = np.ascontiguousarray(levels, dtype=np.int32)
        max_level = len(adjacency) - 1
        offs = [np.ascontiguousarray(a[0], dtype=np.int64) for a in adjacency]
        nbrs = [np.ascontiguousarray(a[1], dtype=np.int32) if len(a[1]) else np.zeros(1, np.int32) for a in adjacency]
        po = (C.c_void_p * len(offs))(*[o.ctypes.data for o in offs])
        pi = (C.c_void_p * len(nbrs))(*[x.ctypes.data for x in nbrs])
        h = new_handle()
        hb.check(hb.lib().hb_hnsw_create(hb.ptr(rows), n, d, hb.dtype_code(rows), metric_code(distance_fn), hb.ptr(lv),
                                         max_level, int(entry_point), po, pi, C.byref(h)))
        super().__init__(h.value, ids)

    def gather_score(self, queries, pair_query, pair_row) -> np.ndarray:
        """scores[p] = distance-fn(query[pair_query[p]], vector[pair_row[p]]) — ultra_fast.clj:192 batched."""
        return gather_score(self, queries, pair_query, pair_row)


def gather_score(index: DeviceIndex, queries, pair_query, pair_row) -> np.ndarray:
    q = hb.as_matrix(queries, allow=(hb.F32, hb.F64))
    pq = np.ascontiguousarray(pair_query, dtype=np.int32)
    pr = np.ascontiguousarray(pair_row, dtype=np.int32)
    out = np.empty(pq.shape[0], dtype=np.float64)
    hb.check(hb.lib().hb_gather_score(index._h, hb.ptr(q), hb.dtype_code(q), q.shape[0], hb.ptr(pq), hb.ptr(pr),
                                      pq.shape[0], hb.ptr(out)))
    return out


def search_knn(index: HnswIndex, query, k, ef=0):
    """(search-knn graph query k), ultra_fast.clj:346-374; ef = 0 -> (max k 50) (:355)."""
    ids, dist = index.search_raw(query, k, ef)
    return results_to_maps(ids, dist, index.ids)[0]


def search_batch(index: HnswIndex, queries, k, ef=0):
    ids, dist = index.search_raw(queries, k, ef)
    return results_to_maps(ids, dist, index.ids)
