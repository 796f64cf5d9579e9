"""Mirror of hnsw.ann.partition.ivf-flat (src/hnsw/ann/partition/ivf_flat.clj): build-index / search-knn /
index-info with the same options, on device-resident list-major slabs."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as hb
from .index import DeviceIndex, metric_code, new_handle, results_to_maps, split_data

# mode table of search-ivf-flat (ivf_flat.clj:243-247).  :turbo probes ONE RANDOM list in the reference
# (non-deterministic, :271-272); here it probes the nearest list, documented in DESIGN.md.
MODE_PROBES = {"turbo": 1, "fast": 2, "balanced": 4, "accurate": 8, "precise": 12}


class IVFFlatIndex(DeviceIndex):
    """->IVFFlatIndex (ivf_flat.clj:21-26): centroids, partitions, vector norms — all on the device."""

    def __init__(self, handle, ids, num_partitions, distance_fn):
        super().__init__(handle, ids)
        self.num_partitions = num_partitions
        self.distance_fn = distance_fn

    # -- parity / persistence helpers ---------------------------------------------------------------
    def export(self):
        """(centroids fp64 [nlist, d], assignments int32 [n])."""
        i = self.info()
        cents = np.empty((i["nlist"], i["dim"]), dtype=np.float64)
        asg = np.empty(i["n"], dtype=np.int32)
        hb.check(hb.lib().hb_ivf_export(self._h, hb.ptr(cents), hb.ptr(asg)))
        return cents, asg

    def probes(self, queries, num_probes):
        q = hb.as_matrix(queries, allow=(hb.F32, hb.F64))
        out = np.empty((q.shape[0], num_probes), dtype=np.int32)
        hb.check(hb.lib().hb_ivf_probes(self._h, hb.ptr(q), hb.dtype_code(q), q.shape[0], num_probes, hb.ptr(out)))
        return out


def _num_probes(mode, num_probes):
    if isinstance(mode, str):
        mode = mode.lstrip(":")
    if mode in MODE_PROBES:
        return MODE_PROBES[mode]
    return int(num_probes) if num_probes else 4  # (or num-probes 4), ivf_flat.clj:249-251


def build_index(data, num_partitions=24, distance_fn="cosine", max_iterations=10, show_progress=False, seed=42):
    """(build-index data & {:keys [num-partitions distance-fn max-iterations show-progress?]}),
    ivf_flat.clj:137-211,300-303.  Defaults as in the reference (:144-148); the k-means++ RNG is
    java.util.Random(42) (:37)."""
    ids, rows = split_data(data)
    if rows.shape[0] == 0:
        raise hb.HbInvalid(hb.ERR_INVALID, "cannot build an IVF-FLAT index from no vectors")
    metric = metric_code(distance_fn)
    h = new_handle()
    hb.check(hb.lib().hb_ivf_build(hb.ptr(rows), rows.shape[0], rows.shape[1], hb.dtype_code(rows), metric,
                                   int(num_partitions), int(max_iterations), int(seed), C.byref(h)))
    return IVFFlatIndex(h.value, ids, int(num_partitions), distance_fn)


def import_index(data, centroids, assignments, distance_fn="cosine"):
    """Same index from given centroids/assignments (oracle-built partitions, or a persisted index)."""
    ids, rows = split_data(data)
    cents = np.ascontiguousarray(centroids, dtype=np.float64)
    asg = np.ascontiguousarray(assignments, dtype=np.int32)
    h = new_handle()
    hb.check(hb.lib().hb_ivf_import(hb.ptr(rows), rows.shape[0], rows.shape[1], hb.dtype_code(rows), metric_code(distance_fn),
                                    hb.ptr(cents), cents.shape[0], hb.ptr(asg), C.byref(h)))
    return IVFFlatIndex(h.value, ids, cents.shape[0], distance_fn)


def search_knn(index: IVFFlatIndex, query, k, mode="balanced", num_probes=None):
    """(search-knn index query k) / (search-knn index query k mode), ivf_flat.clj:305-317."""
    ids, dist = index.search_raw(query, k, _num_probes(mode, num_probes))
    return results_to_maps(ids, dist, index.ids)[0]


def search_batch(index: IVFFlatIndex, queries, k, mode="balanced", num_probes=None):
    """BatchSearchIndex/search-batch* (src/hnsw/api/protocol.clj:58-67): one device call for all queries."""
    ids, dist = index.search_raw(queries, k, _num_probes(mode, num_probes))
    return results_to_maps(ids, dist, index.ids)


def index_info(index: IVFFlatIndex) -> dict:
    """(index-info index), ivf_flat.clj:319-327."""
    i = index.info()
    return {"type": "IVF-FLAT", "vectors": i["n"], "partitions": i["nlist"],
            "avg-partition-size": (i["n"] / i["nlist"]) if i["nlist"] else 0.0, "device-bytes": i["device_bytes"]}


# ---- k-means steps, exposed for the sharded build and for parity tests --------------------------------
def kmeanspp_init(rows, num_partitions, distance_fn="cosine", seed=42) -> np.ndarray:
    """kmeans-plus-plus-init (ivf_flat.clj:32-60) -> chosen row indices."""
    R = hb.as_matrix(rows)
    out = np.empty(num_partitions, dtype=np.int64)
    hb.check(hb.lib().hb_kmeanspp_init(hb.ptr(R), R.shape[0], R.shape[1], hb.dtype_code(R), metric_code(distance_fn),
                                       num_partitions, seed, hb.ptr(out)))
    return out


def assign_to_nearest_centroid(rows, centroids, distance_fn="cosine") -> np.ndarray:
    """assign-to-nearest-centroid (ivf_flat.clj:79-90) for all rows."""
    R = hb.as_matrix(rows)
    Cn = np.ascontiguousarray(centroids, dtype=np.float64)
    out = np.empty(R.shape[0], dtype=np.int32)
    hb.check(hb.lib().hb_kmeans_assign(hb.ptr(R), R.shape[0], R.shape[1], hb.dtype_code(R), metric_code(distance_fn),
                                       hb.ptr(Cn), Cn.shape[0], hb.ptr(out)))
    return out


def compute_centroids(rows, assignments, centroids) -> np.ndarray:
    """compute-centroid per cluster (ivf_flat.clj:66-77), empty keeps the previous centroid (:112-116)."""
    R = hb.as_matrix(rows)
    a = np.ascontiguousarray(assignments, dtype=np.int32)
    Cn = np.array(centroids, dtype=np.float64, order="C", copy=True)
    hb.check(hb.lib().hb_kmeans_update(hb.ptr(R), R.shape[0], R.shape[1], hb.dtype_code(R), hb.ptr(a), Cn.shape[0],
                                       hb.ptr(Cn), None, None))
    return Cn


def partial_sums(rows, assignments, num_partitions):
    """Per-cluster fp64 sums [nlist, d] and counts [nlist] of this shard's rows (multi-GPU Lloyd)."""
    R = hb.as_matrix(rows)
    a = np.ascontiguousarray(assignments, dtype=np.int32)
    sums = np.empty((num_partitions, R.shape[1]), dtype=np.float64)
    cnt = np.empty(num_partitions, dtype=np.int64)
    hb.check(hb.lib().hb_kmeans_update(hb.ptr(R), R.shape[0], R.shape[1], hb.dtype_code(R), hb.ptr(a), num_partitions,
                                       None, hb.ptr(sums), hb.ptr(cnt)))
    return sums, cnt


def partition_vectors_kmeans(rows, num_partitions, distance_fn="cosine", max_iterations=10, seed=42, seed_rows=None):
    """partition-vectors-kmeans (ivf_flat.clj:92-131) -> (centroids fp64, assignments int32)."""
    R = hb.as_matrix(rows)
    cents = np.empty((num_partitions, R.shape[1]), dtype=np.float64)
    asg = np.empty(R.shape[0], dtype=np.int32)
    sr = None if seed_rows is None else np.ascontiguousarray(seed_rows, dtype=np.int64)
    hb.check(hb.lib().hb_kmeans(hb.ptr(R), R.shape[0], R.shape[1], hb.dtype_code(R), metric_code(distance_fn),
                                num_partitions, max_iterations, seed, hb.ptr(sr), hb.ptr(cents), hb.ptr(asg)))
    return cents, asg
