"""ctypes binding of oracle/libhnsw_oracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's arithmetic (see hnsw_oracle.c header).  It may
be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
and by nothing under hnsw_clj_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhnsw_oracle.so")

COSINE, L2, IP = 0, 1, 2
GAUSSIAN, UNIFORM, UNIT, CLUSTERED = 0, 1, 2, 3


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hnsw_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_p = C.c_void_p
_i64, _i32, _f64, _int = C.c_int64, C.c_int32, C.c_double, C.c_int


def _declare(L):
    for name in ("orc_cosine_distance_ultra", "orc_cosine_distance_direct", "orc_euclidean_distance", "orc_dot"):
        f = getattr(L, name)
        f.restype, f.argtypes = _f64, [_p, _p, _i64]
    L.orc_norm.restype, L.orc_norm.argtypes = _f64, [_p, _i64]
    L.orc_row_norms_f32.restype, L.orc_row_norms_f32.argtypes = None, [_p, _i64, _i64, _p]
    L.orc_rng_init.restype, L.orc_rng_init.argtypes = None, [_p, _i64]
    L.orc_rng_next_int.restype, L.orc_rng_next_int.argtypes = _i32, [_p]
    L.orc_rng_next_int_bound.restype, L.orc_rng_next_int_bound.argtypes = _i32, [_p, _i32]
    L.orc_rng_next_double.restype, L.orc_rng_next_double.argtypes = _f64, [_p]
    L.orc_rng_next_gaussian.restype, L.orc_rng_next_gaussian.argtypes = _f64, [_p]
    L.orc_rng_sizeof.restype, L.orc_rng_sizeof.argtypes = _i64, []
    L.orc_gen_dataset.restype, L.orc_gen_dataset.argtypes = None, [_i64, _i64, _int, _i32, _f64, _i64, _p]
    L.orc_exact_knn_f32.restype = None
    L.orc_exact_knn_f32.argtypes = [_p, _i64, _i64, _p, _i64, _i64, _int, _p, _p, _int]
    L.orc_recall.restype, L.orc_recall.argtypes = _f64, [_p, _p, _i64, _i64]
    L.orc_kmeanspp_init.restype = None
    L.orc_kmeanspp_init.argtypes = [_p, _i64, _i64, _i32, _int, _i64, _p, _int]
    L.orc_kmeanspp_init_literal.restype = None
    L.orc_kmeanspp_init_literal.argtypes = [_p, _i64, _i64, _i32, _int, _i64, _p]
    L.orc_assign.restype, L.orc_assign.argtypes = None, [_p, _i64, _i64, _p, _i32, _int, _p, _int]
    L.orc_update_centroids.restype, L.orc_update_centroids.argtypes = None, [_p, _i64, _i64, _p, _i32, _p]
    L.orc_kmeans.restype = None
    L.orc_kmeans.argtypes = [_p, _i64, _i64, _i32, _i32, _int, _i64, _p, _p, _p, _int]
    L.orc_build_lists.restype, L.orc_build_lists.argtypes = None, [_p, _i64, _i32, _p, _p]
    L.orc_strict_log.restype, L.orc_strict_log.argtypes = _f64, [_f64]
    L.orc_lsh_matrices.restype, L.orc_lsh_matrices.argtypes = None, [_i64, _i64, _p]
    L.orc_lsh_hash.restype, L.orc_lsh_hash.argtypes = None, [_p, _i64, _i64, _p, _p]
    L.orc_lsh_search.restype = None
    L.orc_lsh_search.argtypes = [_p, _i64, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _int, _i32, _p, _p]
    L.orc_ivf_search.restype = None
    L.orc_ivf_search.argtypes = [_p, _i64, _i64, _p, _i32, _p, _p, _p, _p, _i64, _i64, _i32, _int, _p, _p, _p, _int]
    L.orc_hnsw_create.restype, L.orc_hnsw_create.argtypes = _p, [_p, _i64, _i64, _int, _i32, _i32, _i64]
    L.orc_hnsw_free.restype, L.orc_hnsw_free.argtypes = None, [_p]
    L.orc_hnsw_build.restype, L.orc_hnsw_build.argtypes = None, [_p]
    L.orc_lightning_seeds.restype, L.orc_lightning_seeds.argtypes = None, [_p, _i64, _i64, _i32, _int, _i64, _p, _int]
    L.orc_lightning_build.restype, L.orc_lightning_build.argtypes = None, [_p, _i64, _i64, _i32, _int, _i64, _p, _p, _int]
    L.orc_hnsw_import.restype, L.orc_hnsw_import.argtypes = None, [_p, _p, _i32, _i32, _p, _p]
    L.orc_hnsw_entry.restype, L.orc_hnsw_entry.argtypes = _i32, [_p]
    L.orc_hnsw_max_level.restype, L.orc_hnsw_max_level.argtypes = _i32, [_p]
    L.orc_hnsw_levels.restype, L.orc_hnsw_levels.argtypes = None, [_p, _p]
    L.orc_hnsw_export_level.restype, L.orc_hnsw_export_level.argtypes = _i64, [_p, _i32, _p, _p]
    L.orc_hnsw_search.restype, L.orc_hnsw_search.argtypes = _i32, [_p, _p, _i32, _i32, _p, _p]
    L.orc_hnsw_dist_evals.restype, L.orc_hnsw_dist_evals.argtypes = _i64, [_p]
    L.orc_gather_score.restype = None
    L.orc_gather_score.argtypes = [_p, _i64, _p, _p, _p, _i64, _int, _p]
    for name in ("orc_simd_dot", "orc_simd_euclidean", "orc_simd_cosine", "orc_simd_cosine_tree"):
        f = getattr(L, name)
        f.restype, f.argtypes = _f64, [_p, _p, _i64, _i32]
    L.orc_pcaf_matrix.restype, L.orc_pcaf_matrix.argtypes = None, [_i64, _i64, _i64, _p]
    L.orc_pcaf_project.restype, L.orc_pcaf_project.argtypes = None, [_p, _i64, _i64, _p, _i64, _i32, _p]
    L.orc_pcaf_search.restype = None
    L.orc_pcaf_search.argtypes = [_p, _p, _i64, _i64, _i64, _p, _p, _i64, _i64, _i64, _i32, _p, _p]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _d64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def ncores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ---- pairwise ------------------------------------------------------------------------------
def cosine_distance(a, b) -> float:
    a, b = _d64(a), _d64(b)
    return lib().orc_cosine_distance_ultra(_ptr(a), _ptr(b), a.size)


def cosine_distance_direct(a, b) -> float:
    a, b = _d64(a), _d64(b)
    return lib().orc_cosine_distance_direct(_ptr(a), _ptr(b), a.size)


def euclidean_distance(a, b) -> float:
    a, b = _d64(a), _d64(b)
    return lib().orc_euclidean_distance(_ptr(a), _ptr(b), a.size)


def dot(a, b) -> float:
    a, b = _d64(a), _d64(b)
    return lib().orc_dot(_ptr(a), _ptr(b), a.size)


def norm(a) -> float:
    a = _d64(a)
    return lib().orc_norm(_ptr(a), a.size)


def row_norms(rows) -> np.ndarray:
    rows = _f32(rows)
    out = np.empty(rows.shape[0], dtype=np.float64)
    lib().orc_row_norms_f32(_ptr(rows), rows.shape[0], rows.shape[1], _ptr(out))
    return out


# ---- java.util.Random ---------------------------------------------------------------------
class JavaRandom:
    def __init__(self, seed: int):
        self._buf = C.create_string_buffer(int(lib().orc_rng_sizeof()))
        lib().orc_rng_init(self._buf, seed)

    def next_int(self, bound: int | None = None) -> int:
        if bound is None:
            return lib().orc_rng_next_int(self._buf)
        return lib().orc_rng_next_int_bound(self._buf, bound)

    def next_double(self) -> float:
        return lib().orc_rng_next_double(self._buf)

    def next_gaussian(self) -> float:
        return lib().orc_rng_next_gaussian(self._buf)


def gen_dataset(n, d, distribution=GAUSSIAN, num_clusters=10, noise=0.1, seed=42) -> np.ndarray:
    """test/data_generator.clj:50-87; returns fp64 [n, d]."""
    out = np.empty((n, d), dtype=np.float64)
    lib().orc_gen_dataset(n, d, distribution, num_clusters, noise, seed, _ptr(out))
    return out


# ---- flat ----------------------------------------------------------------------------------
def exact_knn(rows, queries, k, metric=COSINE, nthreads=None):
    rows, queries = _f32(rows), _f32(queries)
    nq = queries.shape[0]
    n, d = rows.shape if rows.ndim == 2 else (0, queries.shape[1])
    ids = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float64)
    lib().orc_exact_knn_f32(_ptr(rows), n, d, _ptr(queries), nq, k, metric, _ptr(ids), _ptr(dist),
                            nthreads or ncores())
    return ids, dist


def recall(approx_ids, exact_ids) -> float:
    a = np.ascontiguousarray(approx_ids, dtype=np.int64)
    e = np.ascontiguousarray(exact_ids, dtype=np.int64)
    return lib().orc_recall(_ptr(a), _ptr(e), a.shape[0], a.shape[1])


# ---- k-means / IVF ------------------------------------------------------------------------
def kmeanspp_init(rows, nlist, metric=COSINE, seed=42, nthreads=None, literal=False) -> np.ndarray:
    rows = _f32(rows)
    out = np.empty(nlist, dtype=np.int64)
    if literal:
        lib().orc_kmeanspp_init_literal(_ptr(rows), rows.shape[0], rows.shape[1], nlist, metric, seed, _ptr(out))
    else:
        lib().orc_kmeanspp_init(_ptr(rows), rows.shape[0], rows.shape[1], nlist, metric, seed, _ptr(out),
                                nthreads or ncores())
    return out


def lightning_seeds(rows, nlist, metric=COSINE, seed=42, nthreads=None) -> np.ndarray:
    """src/hnsw/ann/partition/lightning.clj:86-109: k-means++ walk weighted by d_i (not d_i^2)."""
    rows = _f32(rows)
    out = np.empty(nlist, dtype=np.int64)
    lib().orc_lightning_seeds(_ptr(rows), rows.shape[0], rows.shape[1], nlist, metric, seed, _ptr(out), nthreads or ncores())
    return out


def lightning_build(rows, nlist, metric=COSINE, seed=42, nthreads=None):
    """build-lightning-index :smart-partition? true (lightning.clj:84-130) -> (centroids fp64 [nlist, d], assignments)."""
    rows = _f32(rows)
    n, d = rows.shape
    cents = np.empty((nlist, d), dtype=np.float64)
    asg = np.empty(n, dtype=np.int32)
    lib().orc_lightning_build(_ptr(rows), n, d, nlist, metric, seed, _ptr(cents), _ptr(asg), nthreads or ncores())
    return cents, asg


def assign(rows, centroids, metric=COSINE, nthreads=None) -> np.ndarray:
    rows, centroids = _f32(rows), _d64(centroids)
    out = np.empty(rows.shape[0], dtype=np.int32)
    lib().orc_assign(_ptr(rows), rows.shape[0], rows.shape[1], _ptr(centroids), centroids.shape[0], metric,
                     _ptr(out), nthreads or ncores())
    return out


def update_centroids(rows, assign_, centroids) -> np.ndarray:
    rows = _f32(rows)
    a = np.ascontiguousarray(assign_, dtype=np.int32)
    c = _d64(centroids).copy()
    lib().orc_update_centroids(_ptr(rows), rows.shape[0], rows.shape[1], _ptr(a), c.shape[0], _ptr(c))
    return c


def kmeans(rows, nlist, iters=10, metric=COSINE, seed=42, seed_rows=None, nthreads=None):
    rows = _f32(rows)
    n, d = rows.shape
    cents = np.empty((nlist, d), dtype=np.float64)
    asg = np.empty(n, dtype=np.int32)
    sr = None if seed_rows is None else np.ascontiguousarray(seed_rows, dtype=np.int64)
    lib().orc_kmeans(_ptr(rows), n, d, nlist, iters, metric, seed, None if sr is None else _ptr(sr),
                     _ptr(cents), _ptr(asg), nthreads or ncores())
    return cents, asg


def build_lists(assign_, nlist):
    a = np.ascontiguousarray(assign_, dtype=np.int32)
    off = np.empty(nlist + 1, dtype=np.int64)
    rows = np.empty(a.shape[0], dtype=np.int64)
    lib().orc_build_lists(_ptr(a), a.shape[0], nlist, _ptr(off), _ptr(rows))
    return off, rows


def ivf_search(rows, centroids, assign_, queries, k, nprobe, coarse_metric=COSINE, nthreads=None,
               return_probes=False):
    rows, queries, centroids = _f32(rows), _f32(queries), _d64(centroids)
    n, d = rows.shape
    nlist = centroids.shape[0]
    off, lrows = build_lists(assign_, nlist)
    norms = row_norms(rows)
    nq = queries.shape[0]
    ids = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float64)
    probes = np.empty((nq, nprobe), dtype=np.int32) if return_probes else None
    lib().orc_ivf_search(_ptr(rows), n, d, _ptr(centroids), nlist, _ptr(off), _ptr(lrows), _ptr(norms),
                         _ptr(queries), nq, k, nprobe, coarse_metric, _ptr(ids), _ptr(dist),
                         None if probes is None else _ptr(probes), nthreads or ncores())
    return (ids, dist, probes) if return_probes else (ids, dist)


# ---- Hybrid LSH (src/hnsw/ann/hash/hybrid_lsh.clj) ---------------------------------------------
LSH_TABLES, LSH_BITS, LSH_PROJ = 8, 12, 64


def strict_log(x: float) -> float:
    return lib().orc_strict_log(float(x))


def lsh_matrices(d: int, seed: int = 42) -> np.ndarray:
    """The 8 projection matrices [8, 64, d] fp64 drawn from one java.util.Random(seed).nextGaussian stream (:24-31, :77-81)."""
    out = np.empty((LSH_TABLES, LSH_PROJ, d), dtype=np.float64)
    lib().orc_lsh_matrices(d, seed, _ptr(out))
    return out


def lsh_hash(rows, matrices) -> np.ndarray:
    """Bucket id of every row in every table, [n, 8] int32 (:33-55, :107-111)."""
    rows, matrices = _f32(rows), _d64(matrices)
    out = np.empty((rows.shape[0], LSH_TABLES), dtype=np.int32)
    lib().orc_lsh_hash(_ptr(rows), rows.shape[0], rows.shape[1], _ptr(matrices), _ptr(out))
    return out


def lsh_buckets(bucket_ids):
    """Per table the buckets as CSR over 4096 ids, members in insertion (data) order: (off [8, 4097], members [8, n])."""
    b = np.asarray(bucket_ids)
    n = b.shape[0]
    off = np.zeros((LSH_TABLES, (1 << LSH_BITS) + 1), dtype=np.int64)
    mem = np.empty((LSH_TABLES, n), dtype=np.int64)
    for t in range(LSH_TABLES):
        off[t, 1:] = np.cumsum(np.bincount(b[:, t], minlength=1 << LSH_BITS))
        mem[t] = np.argsort(b[:, t], kind="stable")
    return off, mem


def lsh_search(rows, matrices, bucket_ids, queries, k, num_probes=6, probe_radius=2, multiprobe=True, main_mult=3):
    """search-hybrid-multiprobe (:261-342) / search-hybrid (:195-259, multiprobe=False; main_mult = 3 for its parallel
    branch, 2 for the sequential one)."""
    rows, queries, matrices = _f32(rows), _f32(queries), _d64(matrices)
    n, d = rows.shape
    off, mem = lsh_buckets(bucket_ids)
    norms = row_norms(rows)
    nq = queries.shape[0]
    ids = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float64)
    lib().orc_lsh_search(_ptr(rows), n, d, _ptr(norms), _ptr(matrices), _ptr(off), _ptr(mem), _ptr(queries), nq, k,
                         num_probes, probe_radius, 1 if multiprobe else 0, main_mult, _ptr(ids), _ptr(dist))
    return ids, dist


# ---- HNSW ---------------------------------------------------------------------------------
class Hnsw:
    """src/hnsw/ultra_fast.clj graph, seeded levels, insertion-ordered neighbour sets."""

    def __init__(self, rows, metric=COSINE, M=16, ef_construction=200, level_seed=42):
        self.rows = _f32(rows)
        self.n, self.d = self.rows.shape
        self._h = lib().orc_hnsw_create(_ptr(self.rows), self.n, self.d, metric, M, ef_construction, level_seed)
        lib().orc_hnsw_build(self._h)

    @classmethod
    def from_graph(cls, rows, levels, entry, adjacency, metric=COSINE, M=16):
        """The oracle traversal over a graph built elsewhere: adjacency = per level (offsets int64 [n+1], ids int32)."""
        self = cls.__new__(cls)
        self.rows = _f32(rows)
        self.n, self.d = self.rows.shape
        self._h = lib().orc_hnsw_create(_ptr(self.rows), self.n, self.d, metric, M, 200, 42)
        lv = np.ascontiguousarray(levels, dtype=np.int32)
        offs = [np.ascontiguousarray(a[0], dtype=np.int64) for a in adjacency]
        ids = [np.ascontiguousarray(a[1], dtype=np.int32) if len(a[1]) else np.zeros(1, np.int32) for a in adjacency]
        po = (C.c_void_p * len(offs))(*[o.ctypes.data for o in offs])
        pi = (C.c_void_p * len(ids))(*[x.ctypes.data for x in ids])
        lib().orc_hnsw_import(self._h, _ptr(lv), int(entry), len(adjacency) - 1, po, pi)
        return self

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_hnsw_free(self._h)
            self._h = None

    @property
    def entry(self) -> int:
        return lib().orc_hnsw_entry(self._h)

    @property
    def max_level(self) -> int:
        return lib().orc_hnsw_max_level(self._h)

    def levels(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.int32)
        lib().orc_hnsw_levels(self._h, _ptr(out))
        return out

    def export_level(self, level: int):
        off = np.empty(self.n + 1, dtype=np.int64)
        tot = lib().orc_hnsw_export_level(self._h, level, _ptr(off), None)
        ids = np.empty(max(tot, 1), dtype=np.int32)
        lib().orc_hnsw_export_level(self._h, level, _ptr(off), _ptr(ids))
        return off, ids[:tot]

    def search(self, queries, k, ef=0):
        queries = _f32(queries)
        nq = queries.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        for i in range(nq):
            lib().orc_hnsw_search(self._h, _ptr(queries[i]), k, ef, _ptr(ids[i]), _ptr(dist[i]))
        return ids, dist

    def dist_evals(self) -> int:
        return lib().orc_hnsw_dist_evals(self._h)


def gather_score(rows, queries, pair_query, pair_row, metric=COSINE) -> np.ndarray:
    rows, queries = _f32(rows), _f32(queries)
    pq = np.ascontiguousarray(pair_query, dtype=np.int32)
    pr = np.ascontiguousarray(pair_row, dtype=np.int32)
    out = np.empty(pq.shape[0], dtype=np.float64)
    lib().orc_gather_score(_ptr(rows), rows.shape[1], _ptr(queries), _ptr(pq), _ptr(pr), pq.shape[0], metric,
                           _ptr(out))
    return out


# ---- a4: float[] Vector-API variants (src/hnsw/simd.clj:18-115) and PCAF (src/hnsw/ann/dimreduct/pcaf.clj) ----------------
def simd_cosine(a, b, lanes=8, tree=False) -> float:
    a, b = _f32(a), _f32(b)
    f = lib().orc_simd_cosine_tree if tree else lib().orc_simd_cosine
    return float(f(_ptr(a), _ptr(b), a.shape[0], lanes))


def simd_dot(a, b, lanes=8) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().orc_simd_dot(_ptr(a), _ptr(b), a.shape[0], lanes))


def simd_euclidean(a, b, lanes=8) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().orc_simd_euclidean(_ptr(a), _ptr(b), a.shape[0], lanes))


def simd_pairwise(A, B, metric=COSINE, lanes=8) -> np.ndarray:
    A, B = _f32(A), _f32(B)
    fn = {COSINE: simd_cosine, L2: simd_euclidean, IP: simd_dot}[metric]
    return np.array([[fn(a, b, lanes) for b in B] for a in A], dtype=np.float64)


def pcaf_matrix(original_dim, target_dim=100, seed=42) -> np.ndarray:
    out = np.empty((target_dim, original_dim), dtype=np.float32)
    lib().orc_pcaf_matrix(original_dim, target_dim, seed, _ptr(out))
    return out


def pcaf_project(matrix, rows, lanes=8) -> np.ndarray:
    matrix, rows = _f32(matrix), _f32(np.atleast_2d(rows))
    out = np.empty((rows.shape[0], matrix.shape[0]), dtype=np.float32)
    lib().orc_pcaf_project(_ptr(matrix), matrix.shape[1], matrix.shape[0], _ptr(rows), rows.shape[0], lanes, _ptr(out))
    return out


def pcaf_search(rows, queries, k, n_components=100, k_filter=32, lanes=8, seed=42):
    rows, queries = _f32(rows), _f32(np.atleast_2d(queries))
    m = pcaf_matrix(rows.shape[1], n_components, seed)
    low = pcaf_project(m, rows, lanes)
    ids = np.empty((queries.shape[0], k), dtype=np.int64)
    dist = np.empty((queries.shape[0], k), dtype=np.float64)
    lib().orc_pcaf_search(_ptr(rows), _ptr(low), rows.shape[0], rows.shape[1], n_components, _ptr(m), _ptr(queries),
                          queries.shape[0], k, k_filter, lanes, _ptr(ids), _ptr(dist))
    return ids, dist
