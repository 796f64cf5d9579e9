"""Device-resident index handles over the C ABI, shared by the namespace mirrors."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib as hb

METRICS = {"cosine": hb.COSINE, "euclidean": hb.L2, "l2": hb.L2, "ip": hb.IP, "dot": hb.IP,
           hb.COSINE: hb.COSINE, hb.L2: hb.L2, hb.IP: hb.IP}


def metric_code(distance_fn) -> int:
    """:distance-fn of the reference (src/hnsw/api.clj:16-19) as an enum; also accepts the mirror functions
    of hnsw_clj_b200.simd_optimized."""
    if callable(distance_fn):
        name = getattr(distance_fn, "__name__", "")
        distance_fn = {"cosine_distance": "cosine", "euclidean_distance": "euclidean", "dot_product": "ip"}.get(name)
    if isinstance(distance_fn, str):
        distance_fn = distance_fn.lstrip(":").lower()
    try:
        return METRICS[distance_fn]
    except KeyError:
        raise hb.HbInvalid(hb.ERR_INVALID, f"unknown distance-fn {distance_fn!r}") from None


def split_data(data):
    """The reference's `data` is a seq of [id vector] pairs (test/data_generator.clj:84-87); a bare matrix
    gets ids 0..n-1.  Returns (ids list | None, matrix)."""
    if hb._is_torch(data) or isinstance(data, np.ndarray):
        return None, hb.as_matrix(data)
    data = list(data)
    if not data:
        return [], np.zeros((0, 0), dtype=np.float32)
    first = data[0]
    if isinstance(first, (tuple, list)) and len(first) == 2 and not np.isscalar(first[1]):
        ids = [p[0] for p in data]
        rows = [p[1] for p in data]
        base = getattr(rows[0], "base", None)
        if (isinstance(base, np.ndarray) and base.ndim == 2 and base.shape[0] == len(rows) and base.flags.c_contiguous
                and all(getattr(r, "base", None) is base for r in rows)
                and all(r.ctypes.data == base.ctypes.data + i * base.strides[0] for i, r in enumerate(rows))):
            return ids, hb.as_matrix(base)  # the rows of one matrix, in order (data_loader.as_data): no copy
        return ids, hb.as_matrix(np.stack([np.asarray(r) for r in rows]))
    return None, hb.as_matrix(np.asarray(data))


def results_to_maps(ids_arr: np.ndarray, dist_arr: np.ndarray, ids: Sequence | None):
    """[{:id :distance} ...] per query, padding (k > n, test/hnsw/core_test.clj:90-96) dropped."""
    out = []
    for row_ids, row_d in zip(ids_arr.tolist(), dist_arr.tolist()):
        out.append([{"id": (ids[i] if ids is not None else i), "distance": d} for i, d in zip(row_ids, row_d) if i >= 0])
    return out


class DeviceIndex:
    """Owns an hb_index*; freed on close()/GC."""

    def __init__(self, handle: int, ids: Sequence | None):
        self._h = C.c_void_p(handle)
        self.ids = list(ids) if ids is not None else None

    # -- raw batched search: ids [nq, k] int64 (row index, -1 padded), distances [nq, k] fp64 ------------
    def search_raw(self, queries, k: int, param: int = 0, out_ids=None, out_dist=None):
        if self._h is None:
            raise hb.HbInvalid(hb.ERR_INVALID, "index is closed")
        q = hb.as_matrix(queries, allow=(hb.F32, hb.F64))
        nq = q.shape[0]
        info = self.info()
        if nq and q.shape[1] != info["dim"]:
            raise hb.HbInvalid(hb.ERR_INVALID, f"query dimension {q.shape[1]} != index dimension {info['dim']}")
        if k < 0:
            raise hb.HbInvalid(hb.ERR_INVALID, "k must be >= 0")
        if out_ids is None:
            out_ids = np.empty((nq, k), dtype=np.int64)
        if out_dist is None:
            out_dist = np.empty((nq, k), dtype=np.float64)
        hb.check(hb.lib().hb_search(self._h, hb.ptr(q), hb.dtype_code(q), nq, k, param, hb.ptr(out_ids), hb.ptr(out_dist)))
        return out_ids, out_dist

    def info(self) -> dict:
        i = hb.HbInfo()
        hb.check(hb.lib().hb_index_info(self._h, C.byref(i)))
        return {"type": i.type, "dtype": i.dtype, "metric": i.metric, "dim": i.dim, "n": i.n, "nlist": i.nlist,
                "max_level": i.max_level, "device_bytes": i.device_bytes}

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            hb.lib().hb_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def new_handle():
    return C.c_void_p()
