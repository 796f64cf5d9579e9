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


def bulk_knn_graph(data, M=16, distance_fn="cosine", level_seed=42, batch=16384):
    """A layered graph for the batched search, built in bulk on the device by the flat search itself.

    Levels are drawn like random-level (ultra_fast.clj:143-147: (long)(ml * -ln U), ml = 1/ln 2, :133), from a
    seeded generator.  The neighbour lists are NOT the product of the reference's incremental insert-single
    (:216-275, host-side graph mutation, out of scope): every member of level l links to its m nearest members of
    level l (m = 2M at level 0, M above, :131) in ascending (distance, row) order, found by the exact flat search —
    the list insert-single's (take m candidates) approaches with a large ef-construction.  Used to get a graph of
    BASELINE configs[4]'s size (1M nodes) in seconds; the traversal over it is the reference's.
    Returns (levels int32 [n], entry point, adjacency = per level (offsets int64 [n+1], ids int32))."""
    from .flat import FlatIndex

    _, rows = split_data(data)
    n = rows.shape[0]
    u = np.random.default_rng(level_seed).random(n)
    levels = np.floor(-np.log(np.maximum(u, 1e-300)) / np.log(2.0)).astype(np.int32)
    max_level = int(levels.max()) if n else 0
    entry = int(np.flatnonzero(levels == max_level)[0]) if n else -1
    adjacency = []
    for l in range(max_level + 1):
        members = np.flatnonzero(levels >= l)
        nm = len(members)
        m = 2 * M if l == 0 else M
        kk = min(m + 1, nm)
        off = np.zeros(n + 1, dtype=np.int64)
        if kk <= 1:
            adjacency.append((off, np.zeros(0, dtype=np.int32)))
            continue
        if l == 0:
            sub = rows
        elif hb._is_torch(rows):
            import torch

            sub = rows[torch.as_tensor(members, device=rows.device)]
        else:
            sub = rows[members]
        nbr = np.empty((nm, kk - 1), dtype=np.int32)
        with FlatIndex(sub, distance_fn=distance_fn) as fx:
            for b0 in range(0, nm, batch):
                b1 = min(b0 + batch, nm)
                ids, _ = fx.search_raw(sub[b0:b1], kk)
                keep = ids != np.arange(b0, b1)[:, None]  # drop the node itself (or, among duplicates, the last hit)
                order = np.argsort(~keep, axis=1, kind="stable")[:, :kk - 1]
                nbr[b0:b1] = members[np.take_along_axis(ids, order, axis=1)]
        counts = np.zeros(n, dtype=np.int64)
        counts[members] = kk - 1
        np.cumsum(counts, out=off[1:])
        adjacency.append((off, nbr.reshape(-1)))
    return levels, entry, adjacency


def build_index_bulk(data, M=16, distance_fn="cosine", level_seed=42):
    """bulk_knn_graph + upload: an HnswIndex whose `graph` attribute keeps (levels, entry, adjacency)."""
    levels, entry, adjacency = bulk_knn_graph(data, M, distance_fn, level_seed)
    ix = HnswIndex(data, levels, entry, adjacency, distance_fn=distance_fn)
    ix.graph = (levels, entry, adjacency)
    return ix


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
