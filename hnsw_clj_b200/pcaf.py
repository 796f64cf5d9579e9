"""Mirror of hnsw.ann.dimreduct.pcaf (src/hnsw/ann/dimreduct/pcaf.clj): build-index / search-knn / index-info /
cleanup of the two-phase "P-HNSW" index — a Gaussian random projection to n-components dimensions, a brute-force scan
of the projected rows, and a full-dimension re-rank of the k-filter best — on the device.

Both phases use the float[] Vector-API cosine distance (src/hnsw/simd.clj:73-115; simd.py); the rows are stored as
floats like the reference's doubles-to-floats copies (pcaf.clj:83-90,170-176)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as hb
from .index import new_handle, results_to_maps, split_data
from .simd import SPECIES_LENGTH

# k-filter per mode (search-knn, pcaf.clj:272-281)
MODE_K_FILTER = {"turbo": 16, "fast": 24, "balanced": 32, "accurate": 48, "precise": 64}


def create_random_projection(original_dim: int, target_dim: int, seed: int = 42) -> np.ndarray:
    """create-random-projection (pcaf.clj:33-46): [target_dim, original_dim] floats from java.util.Random(42)."""
    out = np.empty((target_dim, original_dim), dtype=np.float32)
    hb.check(hb.lib().hb_pcaf_matrix(int(original_dim), int(target_dim), int(seed), hb.ptr(out)))
    return out


def project_vectors(matrix: np.ndarray, vectors, lanes: int = SPECIES_LENGTH) -> np.ndarray:
    """project-vector-simd (pcaf.clj:48-81) for a batch: [n, target_dim] floats."""
    V = np.ascontiguousarray(np.atleast_2d(np.asarray(vectors)), dtype=np.float32)
    out = np.empty((V.shape[0], matrix.shape[0]), dtype=np.float32)
    hb.check(hb.lib().hb_pcaf_project(hb.ptr(matrix), matrix.shape[1], matrix.shape[0], hb.ptr(V), V.shape[0], int(lanes), hb.ptr(out)))
    return out


def _flat_f32(rows: np.ndarray):
    h = new_handle()
    hb.check(hb.lib().hb_flat_create(hb.ptr(rows), rows.shape[0], rows.shape[1], hb.F32, hb.COSINE, C.byref(h)))
    return h


class PCAFIndex:
    """->PCAFIndex (pcaf.clj:96-102): projection, low-dim rows, high-dim float rows, k-filter — rows on the device."""

    def __init__(self, ids, matrix, high, low, n, original_dim, k_filter, lanes):
        self.ids, self.projection, self._high, self._low = ids, matrix, high, low
        self.n, self.original_dim, self.n_components = n, original_dim, int(matrix.shape[0])
        self.k_filter, self.lanes = int(k_filter), int(lanes)
        self.dimension_reduction = original_dim / matrix.shape[0]

    def search_raw(self, queries, k: int, k_filter: int | None = None):
        """(ids [nq, k] int64 row indices, -1 padded; distances [nq, k] fp64) — hb_pcaf_search."""
        if self._high is None:
            raise hb.HbInvalid(hb.ERR_INVALID, "index is closed")
        Q = np.ascontiguousarray(np.atleast_2d(np.asarray(queries)), dtype=np.float32)  # doubles-to-floats, :205-207
        if Q.shape[1] != self.original_dim:
            raise hb.HbInvalid(hb.ERR_INVALID, f"query dimension {Q.shape[1]} != index dimension {self.original_dim}")
        low_q = project_vectors(self.projection, Q, self.lanes)
        ids = np.empty((Q.shape[0], k), dtype=np.int64)
        dist = np.empty((Q.shape[0], k), dtype=np.float64)
        hb.check(hb.lib().hb_pcaf_search(self._high, self._low, hb.ptr(Q), hb.ptr(low_q), Q.shape[0], int(k),
                                         int(k_filter or self.k_filter), self.lanes, hb.ptr(ids), hb.ptr(dist)))
        return ids, dist

    def close(self):
        for name in ("_high", "_low"):
            h = getattr(self, name, None)
            if h is not None:
                hb.lib().hb_index_free(h)
                setattr(self, name, None)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_pcaf_index(data, n_components=100, k_filter=32, show_progress=False, num_threads=4, lanes=SPECIES_LENGTH) -> PCAFIndex:
    """(build-pcaf-index data & {:keys [n-components k-filter show-progress? num-threads]}), pcaf.clj:104-193."""
    ids, rows = split_data(data)
    if rows.shape[0] == 0:
        raise hb.HbInvalid(hb.ERR_INVALID, "cannot build a PCAF index from no vectors")
    if hb._is_torch(rows):
        rows = rows.detach().cpu().numpy()
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    matrix = create_random_projection(rows.shape[1], n_components)
    low = project_vectors(matrix, rows, lanes)
    return PCAFIndex(ids, matrix, _flat_f32(rows), _flat_f32(low), rows.shape[0], rows.shape[1], k_filter, lanes)


build_index = build_pcaf_index  # (build-index data & opts), pcaf.clj:259-269


def _k_filter(index: PCAFIndex, mode):
    if mode is None:
        return index.k_filter
    return MODE_K_FILTER.get(str(mode).lstrip(":"), index.k_filter)


def search_knn(index: PCAFIndex, query_vec, k, mode=None):
    """(search-knn index query-vec k) / (... k mode), pcaf.clj:271-284."""
    ids, dist = index.search_raw(query_vec, k, _k_filter(index, mode))
    return results_to_maps(ids, dist, index.ids)[0]


search_pcaf_parallel = search_knn


def search_batch(index: PCAFIndex, queries, k, mode=None):
    """BatchSearchIndex/search-batch* (src/hnsw/api/protocol.clj:58-67): one device call per phase for the whole batch."""
    ids, dist = index.search_raw(queries, k, _k_filter(index, mode))
    return results_to_maps(ids, dist, index.ids)


def index_info(index: PCAFIndex) -> dict:
    """pcaf.clj:286-295."""
    return {"type": "P-HNSW (SIMD Optimized)", "original-dim": index.original_dim, "reduced-dim": index.n_components,
            "reduction-ratio": index.dimension_reduction, "k-filter": index.k_filter, "vectors": index.n,
            "optimization": "B200: fp32-lane kernels (hb_lanes.cu)"}


def cleanup(index: PCAFIndex) -> None:
    """pcaf.clj:297-301 shuts a thread pool down; here it frees the device rows."""
    index.close()
