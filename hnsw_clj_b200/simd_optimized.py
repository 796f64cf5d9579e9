"""Mirror of hnsw.simd-optimized (src/hnsw/simd_optimized.clj) and the pairwise kernels of hnsw.ultra-fast
(src/hnsw/ultra_fast.clj:43-95) on the device: same names, argument meaning and values (fp64, sequential)."""
from __future__ import annotations

import numpy as np

from . import _lib as hb


def _pairwise(a, b, metric):
    A = hb.as_matrix(a, allow=(hb.F32, hb.F64))
    B = hb.as_matrix(b)
    if A.shape[1] != B.shape[1]:
        raise hb.HbInvalid(hb.ERR_INVALID, "vectors must have the same dimension")
    out = np.empty((A.shape[0], B.shape[0]), dtype=np.float64)
    hb.check(hb.lib().hb_pairwise(hb.ptr(A), A.shape[0], hb.dtype_code(A), hb.ptr(B), B.shape[0], hb.dtype_code(B),
                                  A.shape[1], metric, hb.ptr(out)))
    return out


def cosine_distance(a, b) -> float:
    """simd-opt/cosine-distance (src/hnsw/simd_optimized.clj:145-153) == cosine-distance-ultra."""
    return float(_pairwise(a, b, hb.COSINE)[0, 0])


def euclidean_distance(a, b) -> float:
    """simd-opt/euclidean-distance (:155-160) == euclidean-distance-ultra (src/hnsw/ultra_fast.clj:43-51)."""
    return float(_pairwise(a, b, hb.L2)[0, 0])


def dot_product(a, b) -> float:
    """simd-opt/dot-product (:283-293)."""
    return float(_pairwise(a, b, hb.IP)[0, 0])


def batch_cosine_distances(query, vectors) -> np.ndarray:
    """batch-cosine-distances (:176-179): one query against a list of vectors."""
    return _pairwise(query, vectors, hb.COSINE)[0]


def batch_euclidean_distances(query, vectors) -> np.ndarray:
    return _pairwise(query, vectors, hb.L2)[0]


def batch_distances(queries, vectors, metric="cosine") -> np.ndarray:
    """All-pairs form [nq, n] of batch-distances-parallel (:164-174)."""
    from .index import metric_code

    return _pairwise(queries, vectors, metric_code(metric))


def precompute_norms(vectors) -> np.ndarray:
    """precompute-norms (:206-216) / the norm pass of build-ivf-flat-index (ivf_flat.clj:161-179)."""
    V = hb.as_matrix(vectors)
    out = np.empty(V.shape[0], dtype=np.float64)
    hb.check(hb.lib().hb_row_norms(hb.ptr(V), V.shape[0], V.shape[1], hb.dtype_code(V), hb.ptr(out)))
    return out


def top_k_distances(query, vectors, k, distance_fn="cosine"):
    """top-k-distances (:271-280): all distances, full (stable) sort, take k -> [[index distance] ...]."""
    from .flat import FlatIndex

    with FlatIndex(vectors, distance_fn=distance_fn) as ix:
        ids, dist = ix.search_raw(query, k)
    return [[int(i), float(d)] for i, d in zip(ids[0], dist[0]) if i >= 0]
