"""Mirror of hnsw.ann.hash.hybrid-lsh (src/hnsw/ann/hash/hybrid_lsh.clj): build-index / search-knn / search-hybrid /
search-hybrid-multiprobe / index-info.  SURVEY §8 f3: the LSH index is host bookkeeping (hash tables) around the same
scan arithmetic as search-partition — here the hashing is `hb_pairwise` (inner products against the projection rows),
the bucket scan is `hb_gather_score` over the probed buckets' members, the final sort + take k is `hb_topk_merge`.

What the reference computes, and what is kept:
* projections: NUM-HASH-TABLES = 8 matrices of PROJECTION-DIM = 64 x d doubles from one `java.util.Random(42)`
  nextGaussian stream (:12-14, :24-31, :77-81) — `hb_lsh_matrices` restates it bit for bit; only the first NUM-HASH-BITS
  = 12 rows of a table reach the bucket id (hash-to-bucket-id, :47-55);
* hash bit i = (sum_j v[j] * row_i[j] >= 0.0), a sequential fp64 sum (:33-45);
* a search probes, table by table, the query's bucket (and, multi-probe, the buckets one flipped bit away, :301-308),
  keeps the best k*2 / k*3 / k of each bucket (:147-193), concatenates, drops repeated ids (first kept), stable-sorts by
  distance and takes k (:244-259, :327-342).  The per-bucket cut never changes the result: a row among the k best of
  the deduplicated union is among the k best of the bucket that first holds it, so the device scores the whole
  buckets and selects once — the k smallest by (distance, first position in the concatenation).
* distance = 1 - dot / (qnorm * vnorm) with the precomputed norms (:159-167); zero-norm vectors give 1.0 here (the
  guard of cosine-distance-ultra) where the reference divides by zero.
The parallel branch of the multi-probe search lets its tasks append concurrently (:281-312), which can only reorder
rows of exactly equal distance; this mirror always concatenates table by table like the sequential branch."""
from __future__ import annotations

import numpy as np

from . import _lib as hb
from .flat import FlatIndex
from .index import metric_code, results_to_maps, split_data
from .simd_optimized import _pairwise

NUM_HASH_TABLES = 8   # :12
NUM_HASH_BITS = 12    # :13
PROJECTION_DIM = 64   # :14

# search-knn's modes (:350-364): (num-probes, probe-radius)
_MODES = {"turbo": (2, 1), "fast": (4, 1), "balanced": (6, 2), "accurate": (8, 3), "precise": (8, 4)}


def projection_matrices(d: int, seed: int = 42) -> np.ndarray:
    """[8, 64, d] fp64: generate-random-matrix for every table from one Random(seed) (:24-31, :77-81)."""
    out = np.empty((NUM_HASH_TABLES, PROJECTION_DIM, d), dtype=np.float64)
    hb.check(hb.lib().hb_lsh_matrices(d, NUM_HASH_TABLES, PROJECTION_DIM, seed, hb.ptr(out)))
    return out


def bucket_ids(vectors, matrices, chunk: int = 65536) -> np.ndarray:
    """compute-hash-vector + hash-to-bucket-id for every vector and table (:33-55): [n, 8] int32."""
    V = hb.as_matrix(vectors, allow=(hb.F32, hb.F64))
    d = V.shape[1]
    proj = np.ascontiguousarray(matrices[:, :NUM_HASH_BITS, :].reshape(NUM_HASH_TABLES * NUM_HASH_BITS, d))
    out = np.empty((V.shape[0], NUM_HASH_TABLES), dtype=np.int32)
    weights = (1 << np.arange(NUM_HASH_BITS, dtype=np.int64))
    for r0 in range(0, V.shape[0], chunk):
        part = V[r0:r0 + chunk]
        if hb._is_torch(part):
            part = part.contiguous()
        dots = _pairwise(part, proj, hb.IP)  # [rows, 96] sequential fp64 sums
        bits = (dots >= 0.0).reshape(dots.shape[0], NUM_HASH_TABLES, NUM_HASH_BITS)
        out[r0:r0 + dots.shape[0]] = (bits * weights).sum(axis=2)
    return out


class HybridIndex:
    """->HybridIndex (:19-22): hash tables (host: per table a CSR over 4096 buckets, members in insertion order),
    random matrices, data + norms (device, a flat index)."""

    def __init__(self, flat: FlatIndex, ids, matrices, buckets, distance_fn):
        self.flat, self.ids, self.matrices, self.distance_fn = flat, ids, matrices, distance_fn
        self.n = int(buckets.shape[0])
        nb = 1 << NUM_HASH_BITS
        self.bucket_off = np.zeros((NUM_HASH_TABLES, nb + 1), dtype=np.int64)
        self.bucket_members = np.empty((NUM_HASH_TABLES, self.n), dtype=np.int64)
        for t in range(NUM_HASH_TABLES):
            self.bucket_off[t, 1:] = np.cumsum(np.bincount(buckets[:, t], minlength=nb))
            self.bucket_members[t] = np.argsort(buckets[:, t], kind="stable")
        self.buckets = buckets

    def close(self):
        self.flat.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def build_index(data, distance_fn="cosine", show_progress=False, num_threads=8) -> HybridIndex:
    """(build-lsh-index data & {:keys [distance-fn show-progress? num-threads]}), :66-145, :345-348."""
    ids, rows = split_data(data)
    if rows.shape[0] == 0:
        raise hb.HbInvalid(hb.ERR_INVALID, "cannot build an LSH index from no vectors")
    # search-bucket-brute-force always scores 1 - dot / (qnorm * vnorm) whatever :distance-fn is (query-norm is always
    # truthy, hybrid_lsh.clj:159-167,199): the backing flat index is cosine, distance_fn stays as metadata only
    metric_code(distance_fn)  # still rejects an unknown :distance-fn like the reference's dispatch
    flat = FlatIndex(rows, "cosine")
    matrices = projection_matrices(rows.shape[1])
    return HybridIndex(flat, ids, matrices, bucket_ids(rows, matrices), distance_fn)


def _probe_lists(index: HybridIndex, qb: np.ndarray, num_probes: int, probe_radius: int, multiprobe: bool):
    """Per query the probed (table, bucket) pairs in the reference's order: [nq, P] flat bucket index t * 4096 + b."""
    probes = min(int(num_probes), NUM_HASH_TABLES)
    radius = min(int(probe_radius), NUM_HASH_BITS) if multiprobe else 0
    nb = 1 << NUM_HASH_BITS
    cols = []
    for t in range(probes):
        base = qb[:, t].astype(np.int64)
        cols.append(t * nb + base)
        for bit in range(radius):
            cols.append(t * nb + ((base ^ (1 << bit)) & (nb - 1)))
    return np.stack(cols, axis=1)


def _search(index: HybridIndex, queries, k: int, num_probes: int, probe_radius: int, multiprobe: bool):
    Q = hb.as_matrix(queries, allow=(hb.F32, hb.F64))
    if hb._is_torch(Q):
        Q = Q.cpu().numpy()
    nq, n, nb = Q.shape[0], index.n, 1 << NUM_HASH_BITS
    ids = np.full((nq, k), -1, dtype=np.int64)
    dist = np.full((nq, k), np.inf, dtype=np.float64)
    if nq == 0 or k == 0:
        return ids, dist
    qb = bucket_ids(Q, index.matrices)
    flat_b = _probe_lists(index, qb, num_probes, probe_radius, multiprobe)           # [nq, P]
    t_of, b_of = flat_b // nb, flat_b % nb
    start = index.bucket_off[t_of, b_of] + t_of * n                                   # into bucket_members.reshape(-1)
    length = index.bucket_off[t_of, b_of + 1] - index.bucket_off[t_of, b_of]
    seg_len = length.reshape(-1)
    total = int(seg_len.sum())
    if total == 0:
        return ids, dist
    seg_first = np.cumsum(seg_len) - seg_len                                          # first output slot of each segment
    src = np.repeat(start.reshape(-1) - seg_first, seg_len) + np.arange(total)
    cand_row = index.bucket_members.reshape(-1)[src]
    cand_q = np.repeat(np.repeat(np.arange(nq), flat_b.shape[1]), seg_len)
    # repeated ids: the first occurrence in the query's concatenation is kept (:244-249, :327-332)
    _, first = np.unique(cand_q * n + cand_row, return_index=True)
    first.sort()
    cand_row, cand_q = cand_row[first], cand_q[first]
    from .ultra_fast import gather_score

    d_pair = gather_score(index.flat, Q, cand_q.astype(np.int32), cand_row.astype(np.int32))
    # top-k per query by (distance, position in the concatenation): parts of k candidates for hb_topk_merge
    per_q = np.bincount(cand_q, minlength=nq)
    nparts = max(1, -(-int(per_q.max()) // k))
    pos = np.arange(cand_q.shape[0]) - np.repeat(np.cumsum(per_q) - per_q, per_q)
    part_d = np.full((nq, nparts * k), np.inf, dtype=np.float64)
    part_i = np.full((nq, nparts * k), -1, dtype=np.int64)
    part_d[cand_q, pos] = d_pair
    part_i[cand_q, pos] = cand_row
    part_d = np.ascontiguousarray(part_d.reshape(nq, nparts, k).transpose(1, 0, 2))
    part_i = np.ascontiguousarray(part_i.reshape(nq, nparts, k).transpose(1, 0, 2))
    hb.check(hb.lib().hb_topk_merge(hb.ptr(part_d), hb.ptr(part_i), nparts, nq, k, hb.ptr(ids), hb.ptr(dist)))
    return ids, dist


def search_hybrid_raw(index, queries, k, num_probes=2):
    """search-hybrid (:195-259) for a batch: (ids [nq, k] int64 row indices, -1 padded; distances)."""
    return _search(index, queries, k, num_probes, 0, False)


def search_hybrid_multiprobe_raw(index, queries, k, num_probes=6, probe_radius=2):
    """search-hybrid-multiprobe (:261-342) for a batch."""
    return _search(index, queries, k, num_probes, probe_radius, True)


def search_hybrid(index, query, k, num_probes=2, parallel=True):
    ids, dist = search_hybrid_raw(index, np.asarray(query)[None, :], k, num_probes)
    return results_to_maps(ids, dist, index.ids)[0]


def search_hybrid_multiprobe(index, query, k, num_probes=6, probe_radius=2, parallel=True):
    ids, dist = search_hybrid_multiprobe_raw(index, np.asarray(query)[None, :], k, num_probes, probe_radius)
    return results_to_maps(ids, dist, index.ids)[0]


def _mode(mode):
    return _MODES.get(str(mode).lstrip(":"), _MODES["balanced"]) if mode is not None else _MODES["balanced"]


def search_knn(index, query, k, mode=None):
    """(search-knn index query-vec k) / (search-knn index query-vec k mode), :350-364."""
    probes, radius = _mode(mode)
    return search_hybrid_multiprobe(index, query, k, probes, radius)


def search_batch(index, queries, k, mode=None):
    """search-batch* (src/hnsw/api/protocol.clj:58-67) over the LSH index: one device call per stage for the whole batch."""
    probes, radius = _mode(mode)
    ids, dist = search_hybrid_multiprobe_raw(index, queries, k, probes, radius)
    return results_to_maps(ids, dist, index.ids)


def index_info(index: HybridIndex):
    """index-info (:366-379)."""
    total_buckets = int((np.diff(index.bucket_off, axis=1) > 0).sum())
    return {"type": "Hybrid LSH Index", "vectors": index.n, "hash-tables": NUM_HASH_TABLES,
            "buckets-per-table": 1 << NUM_HASH_BITS, "total-buckets": total_buckets,
            "avg-bucket-size": (index.n / total_buckets) if total_buckets > 0 else 0}
