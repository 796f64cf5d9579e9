"""Mirror of hnsw.helper.data-loader (src/hnsw/helper/data_loader.clj:7-45) and hnsw.bench/load-bible-data-fn
(src/hnsw/bench.clj:24-44): the embeddings JSON ({"verses": [{"id", "text", "embedding": [768 doubles]}, ...]}) ->
ids + ONE contiguous matrix ready for the device.

The reference materialises a double[] per verse.  Here the values go into a single [n, d] buffer — fp32 when every
value is fp32-representable (sentence-transformers export fp32 embeddings, scripts/export_complete_bible.py:91, so the
widening the reference does is exact and HB_F32 storage loses nothing), else fp64 — in pinned host memory when torch
is available, so the build's host->device copy runs at PCIe rate without a staging pass."""
from __future__ import annotations

import json
import time

import numpy as np


def _pinned_like(a: np.ndarray) -> np.ndarray:
    """A page-locked copy of `a` (as a numpy view of a pinned torch tensor kept alive by the array's base)."""
    try:
        import torch

        if not torch.cuda.is_available():
            return a
        t = torch.empty(a.shape, dtype=torch.float32 if a.dtype == np.float32 else torch.float64).pin_memory()
        v = t.numpy()
        v[...] = a
        return v
    except Exception:
        return a


def to_matrix(embeddings, pinned=True):
    """list of per-row sequences of doubles -> ([n, d] float32 or float64 matrix, {"fp32-exact", "unit-norm"})."""
    a64 = np.asarray(embeddings, dtype=np.float64)
    if a64.ndim != 2:
        raise ValueError("embeddings must all have the same dimension")
    a32 = a64.astype(np.float32)
    fp32_exact = bool((a32.astype(np.float64) == a64).all())
    out = a32 if fp32_exact else a64
    norms = np.sqrt((a64 * a64).sum(axis=1)) if len(a64) else np.zeros(0)
    unit = bool(len(a64) and np.abs(norms - 1.0).max() < 1e-5)  # normalize_embeddings=True exports
    return (_pinned_like(out) if pinned else out), {"fp32-exact": fp32_exact, "unit-norm": unit}


def load_bible_vectors(filename, pinned=True):
    """(load-bible-vectors filename) -> {:vectors :text-map :metadata}; nil (None) on an unreadable file, as in the
    reference (:41-45).  :vectors is {"ids": [...], "matrix": [n, d]} instead of a seq of [id double[]] pairs — every
    build_index of this package takes it as (ids, matrix) via `as_data`."""
    t0 = time.perf_counter()
    try:
        with open(filename) as f:
            data = json.load(f)
        verses = data["verses"]
        ids = [v["id"] for v in verses]
        matrix, props = to_matrix([v["embedding"] for v in verses], pinned)
    except (OSError, ValueError, KeyError, TypeError) as e:
        print(f"Error loading file: {e}")
        return None
    return {"vectors": {"ids": ids, "matrix": matrix},
            "text-map": {v["id"]: v.get("text") for v in verses},
            "metadata": {"count": len(ids), "dimension": int(matrix.shape[1]) if len(ids) else 0,
                         "load-time": (time.perf_counter() - t0) * 1e3, "filename": filename, **props}}


def as_data(loaded):
    """[(id, row), ...] view accepted by every build_index (rows are views into the one matrix, no copies)."""
    v = loaded["vectors"]
    return list(zip(v["ids"], v["matrix"]))
