"""Mirror of hnsw.ann.partition.lightning (src/hnsw/ann/partition/lightning.clj): build-index / search-knn /
index-info on the same device-resident list-major slabs as IVF-FLAT — Lightning is the IVF-FLAT scan behind a
different partitioning (k-means++ walk weighted by d_i, no Lloyd rounds) and a percentage-based probe count."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as hb
from .index import metric_code, new_handle, results_to_maps, split_data
from .ivf_flat import IVFFlatIndex, import_index

# mode tables of search-lightning (lightning.clj:193-226): fraction of the partitions searched.  :turbo / :fast pick
# RANDOM partitions in the reference (:use-centroids false, :276-279 — non-deterministic); here every mode routes by
# centroid distance, as :balanced / :accurate / :precise do.
_MODE_PERCENT = (
    (64, {"turbo": 0.03, "fast": 0.05, "balanced": 0.10, "accurate": 0.16, "precise": 0.25}),
    (32, {"turbo": 0.05, "fast": 0.08, "balanced": 0.15, "accurate": 0.25, "precise": 0.40}),
)
_MODE_PERCENT_24 = {"turbo": 0.08, "fast": 0.12, "balanced": 0.20, "accurate": 0.33, "precise": 0.50}
_MODE_PERCENT_SMALL = {"turbo": 0.10, "fast": 0.15, "balanced": 0.30, "accurate": 0.45, "precise": 0.60}


class LightningIndex(IVFFlatIndex):
    """->LightningIndex (lightning.clj:11-15): partitions, centroids, norms — on the device."""


def num_partitions_to_search(num_partitions: int, mode=None, search_percent=None) -> int:
    """(max 1 (int (* num-partitions final-percent))), lightning.clj:228-262."""
    percent = None
    if mode is not None:
        m = str(mode).lstrip(":")
        table = _MODE_PERCENT_SMALL
        for least, t in _MODE_PERCENT:
            if num_partitions >= least:
                table = t
                break
        else:
            if num_partitions == 24:
                table = _MODE_PERCENT_24
        percent = table.get(m)  # an unknown mode leaves percent nil -> the dynamic default below (:253-260)
    else:
        percent = search_percent
    if percent is None:
        percent = (0.30 if num_partitions <= 16 else 0.20 if num_partitions == 24 else 0.15 if num_partitions <= 32
                   else 0.10 if num_partitions <= 64 else 0.08)
    return max(1, int(num_partitions * percent))


def build_index(data, num_partitions=32, distance_fn="cosine", show_progress=False, smart_partition=False, seed=42):
    """(build-lightning-index data & {:keys [num-partitions distance-fn show-progress? smart-partition?]}),
    lightning.clj:46-160.  smart_partition=True is the k-means++ partitioning (:84-130), bit-compatible with the
    reference (java.util.Random(42)).  The default is the reference's shuffle split (:132-137): the rows are dealt
    into num-partitions chunks of ceil(n / num-partitions) after a shuffle — unseeded in the reference, seeded here."""
    ids, rows = split_data(data)
    n = rows.shape[0]
    if n == 0:
        raise hb.HbInvalid(hb.ERR_INVALID, "cannot build a Lightning index from no vectors")
    metric = metric_code(distance_fn)
    if smart_partition:
        h = new_handle()
        hb.check(hb.lib().hb_lightning_build(hb.ptr(rows), n, rows.shape[1], hb.dtype_code(rows), metric, int(num_partitions),
                                             int(seed), C.byref(h)))
        return LightningIndex(h.value, ids, int(num_partitions), distance_fn)
    size = -(-n // int(num_partitions))
    perm = np.random.default_rng(seed).permutation(n)
    asg = np.empty(n, dtype=np.int32)
    asg[perm] = np.arange(n) // size
    nparts = int(asg.max()) + 1  # partition-all may yield fewer chunks than requested
    cents = np.zeros((nparts, rows.shape[1]), dtype=np.float64)
    hb.check(hb.lib().hb_kmeans_update(hb.ptr(rows), n, rows.shape[1], hb.dtype_code(rows), hb.ptr(asg), nparts, hb.ptr(cents),
                                       None, None))  # compute-centroid per chunk (:136)
    ix = import_index(rows, cents, asg, distance_fn)
    return LightningIndex(ix.__dict__.pop("_h").value, ids, nparts, distance_fn)


def search_knn(index: LightningIndex, query, k, search_percent=None, parallel=False, mode=None):
    """(search-knn index query k) / (… k search-percent) / (… k :mode), lightning.clj:304-327; a keyword in the
    search-percent position is a mode (:318-320).  `parallel` only selects the host threading in the reference."""
    if isinstance(search_percent, str):
        mode, search_percent = search_percent, None
    nprobe = num_partitions_to_search(index.num_partitions, mode, search_percent)
    ids, dist = index.search_raw(query, k, nprobe)
    return results_to_maps(ids, dist, index.ids)[0]


def search_batch(index: LightningIndex, queries, k, search_percent=None, mode=None):
    if isinstance(search_percent, str):
        mode, search_percent = search_percent, None
    nprobe = num_partitions_to_search(index.num_partitions, mode, search_percent)
    ids, dist = index.search_raw(queries, k, nprobe)
    return results_to_maps(ids, dist, index.ids)


def index_info(index: LightningIndex) -> dict:
    """lightning.clj:329-336."""
    i = index.info()
    return {"type": "Lightning Index", "vectors": i["n"], "partitions": i["nlist"],
            "avg-partition-size": i["n"] / max(i["nlist"], 1)}
