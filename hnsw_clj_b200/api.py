"""Mirror of hnsw.api (src/hnsw/api.clj:6-33) and the ANNIndex / BatchSearchIndex protocols
(src/hnsw/api/protocol.clj:9-28,58-67) for the device-resident index types."""
from __future__ import annotations

from . import flat, hybrid_lsh, ivf_flat, lightning, pcaf, ultra_fast
from .index import DeviceIndex

INDEX_TYPES = {0: "FLAT", 1: "IVF-FLAT", 2: "HNSW"}


def index(data, index_type="ivf-flat", metric="cosine", **opts) -> DeviceIndex:
    """(hnsw.api/index {...}): metric -> distance-fn (:16-19).  index_type: 'flat' | 'ivf-flat'."""
    t = str(index_type).lstrip(":").lower()
    if t == "flat":
        return flat.FlatIndex(data, distance_fn=metric)
    if t in ("ivf-flat", "ivf_flat", "ivf"):
        return ivf_flat.build_index(data, distance_fn=metric, **opts)
    if t == "lightning":
        return lightning.build_index(data, distance_fn=metric, **opts)
    if t in ("hybrid-lsh", "lsh"):
        return hybrid_lsh.build_index(data, distance_fn=metric, **opts)
    if t in ("pcaf", "p-hnsw"):
        return pcaf.build_index(data, **opts)
    raise ValueError(f"unknown index type {index_type!r}")


def search(idx: DeviceIndex, query, k, mode="balanced", **opts):
    """(hnsw.api/search index query k)."""
    return search_knn_(idx, query, k, mode, **opts)


# ---- ANNIndex -----------------------------------------------------------------------------------------
def search_knn_(idx: DeviceIndex, query, k, mode="balanced", **opts):
    """ANNIndex/search-knn* [this query k mode]."""
    # Lightning first: it shares the IVF-FLAT device layout (subclass) but probes by its own percentage tables
    # (src/hnsw/ann/partition/lightning.clj:193-262), not the IVF mode table
    if isinstance(idx, lightning.LightningIndex):
        return lightning.search_knn(idx, query, k, opts.get("search_percent"), mode=mode)
    if isinstance(idx, hybrid_lsh.HybridIndex):
        return hybrid_lsh.search_knn(idx, query, k, mode)
    if isinstance(idx, pcaf.PCAFIndex):
        return pcaf.search_knn(idx, query, k, mode)
    if isinstance(idx, ivf_flat.IVFFlatIndex):
        return ivf_flat.search_knn(idx, query, k, mode, opts.get("num_probes"))
    if isinstance(idx, ultra_fast.HnswIndex):
        return ultra_fast.search_knn(idx, query, k, opts.get("ef", 0))
    return idx.search_knn(query, k)


def index_info_(idx: DeviceIndex) -> dict:
    if isinstance(idx, lightning.LightningIndex):
        return lightning.index_info(idx)
    if isinstance(idx, hybrid_lsh.HybridIndex):
        return hybrid_lsh.index_info(idx)
    if isinstance(idx, pcaf.PCAFIndex):
        return pcaf.index_info(idx)
    if isinstance(idx, ivf_flat.IVFFlatIndex):
        return ivf_flat.index_info(idx)
    i = idx.info()
    return {"type": INDEX_TYPES[i["type"]], "vectors": i["n"], "device-bytes": i["device_bytes"]}


def index_type_(idx: DeviceIndex) -> str:
    if isinstance(idx, lightning.LightningIndex):
        return "LIGHTNING"
    if isinstance(idx, hybrid_lsh.HybridIndex):
        return "HYBRID-LSH"
    if isinstance(idx, pcaf.PCAFIndex):
        return "PCAF"
    return INDEX_TYPES[idx.info()["type"]]


# ---- BatchSearchIndex ---------------------------------------------------------------------------------
def search_batch_(idx: DeviceIndex, queries, k, mode="balanced", **opts):
    """BatchSearchIndex/search-batch* [this queries k mode] -> vector of result vectors.  The reference's
    default is (mapv #(search-knn* ...)) (protocol.clj:92-95); here it is ONE batched device call."""
    if isinstance(idx, lightning.LightningIndex):
        return lightning.search_batch(idx, queries, k, opts.get("search_percent"), mode=mode)
    if isinstance(idx, hybrid_lsh.HybridIndex):
        return hybrid_lsh.search_batch(idx, queries, k, mode)
    if isinstance(idx, pcaf.PCAFIndex):
        return pcaf.search_batch(idx, queries, k, mode)
    if isinstance(idx, ivf_flat.IVFFlatIndex):
        return ivf_flat.search_batch(idx, queries, k, mode, opts.get("num_probes"))
    if isinstance(idx, ultra_fast.HnswIndex):
        return ultra_fast.search_batch(idx, queries, k, opts.get("ef", 0))
    return idx.search_batch(queries, k)
