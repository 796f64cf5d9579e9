"""GPU parity: the CUDA path (through the C ABI, via the host mirror) against the CPU oracle on the same
seeded inputs.  Bit-exact ids AND bit-exact fp64 distances (the exact path restates the reference's
sequential sums on the device), so the north-star tolerance (1e-5 relative) holds with margin 0."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hnsw_clj_b200 as pkg
    from hnsw_clj_b200 import _lib

    _lib.check(_lib.lib().hb_init(0))
    return pkg


def rng_rows(n, d, seed, unit=False, clustered=0):
    r = np.random.default_rng(seed)
    if clustered:
        c = r.standard_normal((clustered, d))
        x = c[r.integers(0, clustered, n)] + 0.1 * r.standard_normal((n, d))
    else:
        x = r.standard_normal((n, d))
    if unit:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def same_bits(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool((a.view(np.int64) == b.view(np.int64)).all())


# ---- pairwise: reference KATs through the device (test/hnsw/core_test.clj:9-31) ---------------------
def test_pairwise_kats(hb):
    from hnsw_clj_b200 import simd_optimized as so

    assert so.euclidean_distance([1, 2, 3], [1, 2, 3]) == 0.0
    assert so.euclidean_distance([0, 0], [3, 4]) == 5.0
    assert so.euclidean_distance([1, 2, 3], [4, 5, 6]) == 5.196152422706632
    assert so.cosine_distance([1, 2, 3], [1, 2, 3]) < 0.001
    assert so.cosine_distance([1, 0], [-1, 0]) == 2.0
    assert so.cosine_distance([1, 2, 3], [4, 5, 6]) == 0.025368153802923787
    assert so.cosine_distance([0, 0, 0], [1, 2, 3]) == 1.0  # zero-norm guard, ultra_fast.clj:92-95
    assert so.dot_product([1, 2, 3], [4, 5, 6]) == 32.0


@pytest.mark.parametrize("d", [3, 7, 64, 100, 768])
@pytest.mark.parametrize("qd", [np.float32, np.float64])
def test_pairwise_matrix_matches_oracle(hb, d, qd):
    from hnsw_clj_b200 import simd_optimized as so

    a = np.random.default_rng(d).standard_normal((5, d)).astype(qd)
    b = rng_rows(9, d, d + 1)
    for metric, fn in (("cosine", orc.cosine_distance), ("euclidean", orc.euclidean_distance), ("ip", orc.dot)):
        got = so.batch_distances(a, b, metric)
        want = np.array([[fn(x.astype(np.float64), y.astype(np.float64)) for y in b] for x in a])
        assert same_bits(got, want), (metric, d)


def test_row_norms(hb):
    from hnsw_clj_b200 import simd_optimized as so

    for d in (5, 768):
        rows = rng_rows(301, d, 3)
        assert same_bits(so.precompute_norms(rows), orc.row_norms(rows))
        r64 = np.random.default_rng(1).standard_normal((17, d))
        assert same_bits(so.precompute_norms(r64), np.array([orc.norm(x) for x in r64]))


# ---- exact flat search (src/hnsw/bench.clj:72-84) ----------------------------------------------------
@pytest.mark.parametrize("n,d,nq,k", [(3000, 96, 70, 10), (2000, 768, 33, 10), (257, 30, 3, 100), (129, 5, 1, 1),
                                      (5000, 128, 200, 64)])
@pytest.mark.parametrize("metric", ["cosine", "euclidean", "ip"])
def test_flat_search_parity(hb, n, d, nq, k, metric):
    from hnsw_clj_b200.flat import FlatIndex

    rows, q = rng_rows(n, d, 10), rng_rows(nq, d, 11)
    code = {"cosine": orc.COSINE, "euclidean": orc.L2, "ip": orc.IP}[metric]
    want_ids, want_d = orc.exact_knn(rows, q, k, code)
    with FlatIndex(rows, distance_fn=metric) as ix:
        ids, dist = ix.search_raw(q, k)
    assert ids.tolist() == want_ids.tolist()
    assert same_bits(dist, want_d)


def test_flat_edge_cases(hb):
    from hnsw_clj_b200.flat import FlatIndex, compute_exact_knn

    rows = np.eye(4, dtype=np.float32)
    with FlatIndex(rows) as ix:
        ids, dist = ix.search_raw(rows[:2], 6)  # k > n -> n results (test/hnsw/core_test.clj:90-96)
        assert ids[0].tolist() == [0, 1, 2, 3, -1, -1] and np.isinf(dist[:, 4:]).all()
        assert len(ix.search_knn(rows[0], 6)) == 4
        assert ix.search_raw(np.zeros((0, 4), np.float32), 3)[0].shape == (0, 3)
    with FlatIndex(np.zeros((0, 4), np.float32)) as ix:  # empty index -> [] (ultra_fast.clj:349-351)
        assert ix.search_knn(rows[0], 3) == []
    res = compute_exact_knn([(f"vec_{i}", r.astype(np.float64)) for i, r in enumerate(rows)], rows[2], 2)
    assert res[0] == {"id": "vec_2", "distance": 0.0} and res[1]["id"] == "vec_0"
    with FlatIndex(rows[:1]) as ix:  # single vector (core_test.clj:70-78)
        assert ix.search_knn(rows[0], 5) == [{"id": 0, "distance": 0.0}]
    # duplicates: ties keep row order
    dup = np.repeat(rng_rows(5, 16, 2), 4, axis=0)
    with FlatIndex(dup) as ix:
        ids, _ = ix.search_raw(dup[:1], 8)
    assert ids[0].tolist() == orc.exact_knn(dup, dup[:1], 8)[0][0].tolist()
    assert ids[0, :4].tolist() == [0, 1, 2, 3]


def test_flat_bf16_rows_and_f64_queries(hb):
    import torch

    from hnsw_clj_b200.flat import FlatIndex

    rows = torch.from_numpy(rng_rows(4000, 64, 5)).to(torch.bfloat16)
    q = torch.from_numpy(rng_rows(50, 64, 6)).to(torch.bfloat16).float()
    rows_f = rows.float().numpy()
    want_ids, want_d = orc.exact_knn(rows_f, q.numpy(), 100, orc.IP)
    with FlatIndex(rows.cuda(), distance_fn="ip") as ix:
        ids, dist = ix.search_raw(q.cuda(), 100)  # device-resident inputs
        ids2, dist2 = ix.search_raw(q.numpy().astype(np.float64), 100)  # fp64 queries (the reference's double[])
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)
    assert ids2.tolist() == want_ids.tolist() and same_bits(dist2, want_d)


def test_flat_multipass_and_split_select(hb):
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    rows, q = rng_rows(70000, 32, 20), rng_rows(5, 32, 21)
    want_ids, want_d = orc.exact_knn(rows, q, 20)
    with FlatIndex(rows) as ix:
        ids, dist = ix.search_raw(q, 20)  # few queries, long rows -> split select + merge
        assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)
        _lib.set_option("scratch_mb", 1)  # forces several row passes
        try:
            ids, dist = ix.search_raw(q, 20)
        finally:
            _lib.set_option("scratch_mb", 8192)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)


# ---- k-means (ivf_flat.clj:32-131) -------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_kmeans_steps_parity(hb, metric):
    from hnsw_clj_b200 import ivf_flat

    code = orc.COSINE if metric == "cosine" else orc.L2
    rows = rng_rows(3000, 40, 30, clustered=20)
    seeds = ivf_flat.kmeanspp_init(rows, 24, metric, seed=42)
    assert seeds.tolist() == orc.kmeanspp_init(rows, 24, code, seed=42).tolist()
    cents = rows[seeds].astype(np.float64)
    asg = ivf_flat.assign_to_nearest_centroid(rows, cents, metric)
    want = orc.assign(rows, cents, code)
    assert asg.tolist() == want.tolist()
    new = ivf_flat.compute_centroids(rows, asg, cents)
    assert same_bits(new, orc.update_centroids(rows, want, cents))
    sums, cnt = ivf_flat.partial_sums(rows, asg, 24)
    assert cnt.tolist() == np.bincount(want, minlength=24).tolist()
    c2, a2 = ivf_flat.partition_vectors_kmeans(rows, 24, metric, max_iterations=4, seed=42)
    oc, oa = orc.kmeans(rows, 24, iters=4, metric=code, seed=42)
    assert a2.tolist() == oa.tolist() and same_bits(c2, oc)


def test_kmeans_empty_cluster_keeps_centroid(hb):
    from hnsw_clj_b200 import ivf_flat

    rows = rng_rows(50, 8, 31)
    cents = np.vstack([rows[:3].astype(np.float64), 100.0 * np.ones((1, 8))])  # 4th centroid attracts nothing (cosine: maybe)
    asg = orc.assign(rows, cents)
    got = ivf_flat.compute_centroids(rows, asg, cents)
    assert same_bits(got, orc.update_centroids(rows, asg, cents))


# ---- IVF-FLAT (ivf_flat.clj:137-294) ------------------------------------------------------------------
@pytest.fixture(scope="module")
def ivf_case():
    rows = rng_rows(6000, 64, 40, clustered=40)
    q = rng_rows(120, 64, 41, clustered=40)
    cents, asg = orc.kmeans(rows, 32, iters=3, seed=42)
    return rows, q, cents, asg


@pytest.mark.parametrize("nprobe", [1, 4, 8, 32, 40])
def test_ivf_search_parity_imported_partitions(hb, ivf_case, nprobe):
    from hnsw_clj_b200 import ivf_flat

    rows, q, cents, asg = ivf_case
    want_ids, want_d, want_p = orc.ivf_search(rows, cents, asg, q, 10, nprobe, return_probes=True)
    with ivf_flat.import_index(rows, cents, asg) as ix:
        ids, dist = ix.search_raw(q, 10, nprobe)
        probes = ix.probes(q, nprobe)
    assert probes.tolist() == want_p.tolist()
    assert ids.tolist() == want_ids.tolist()
    assert same_bits(dist, want_d)


def test_ivf_build_end_to_end_equals_oracle_build(hb):
    from hnsw_clj_b200 import ivf_flat

    rows = rng_rows(4000, 48, 50, clustered=30)
    q = rng_rows(64, 48, 51, clustered=30)
    data = [(f"vec_{i}", r) for i, r in enumerate(rows)]
    ix = ivf_flat.build_index(data, num_partitions=24, max_iterations=10)  # the reference defaults (:144-148)
    cents, asg = ix.export()
    oc, oa = orc.kmeans(rows, 24, iters=10, seed=42)
    assert asg.tolist() == oa.tolist() and same_bits(cents, oc)
    for mode, nprobe in (("fast", 2), ("balanced", 4), ("accurate", 8), ("precise", 12)):
        res = ivf_flat.search_batch(ix, q, 10, mode)
        want_ids, want_d = orc.ivf_search(rows, oc, oa, q, 10, nprobe)
        assert [[r["id"] for r in rr] for rr in res] == [[f"vec_{i}" for i in row] for row in want_ids.tolist()]
        assert same_bits(np.array([[r["distance"] for r in rr] for rr in res]), want_d)
    one = ivf_flat.search_knn(ix, q[0], 10, "custom", num_probes=5)
    assert [r["id"] for r in one] == [f"vec_{i}" for i in orc.ivf_search(rows, oc, oa, q[:1], 10, 5)[0][0]]
    info = ivf_flat.index_info(ix)
    assert info["type"] == "IVF-FLAT" and info["vectors"] == 4000 and info["partitions"] == 24
    ix.close()


def test_ivf_probe_all_lists_equals_flat(hb, ivf_case):
    from hnsw_clj_b200 import ivf_flat
    from hnsw_clj_b200.flat import FlatIndex

    rows, q, cents, asg = ivf_case
    with ivf_flat.import_index(rows, cents, asg) as ix, FlatIndex(rows) as fx:
        a = ix.search_raw(q, 10, 32)
        b = fx.search_raw(q, 10)
    assert a[0].tolist() == b[0].tolist() and same_bits(a[1], b[1])


# ---- HNSW neighbour-candidate scoring (ultra_fast.clj:185-204) ---------------------------------------
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_gather_score_parity(hb, metric):
    from hnsw_clj_b200.flat import FlatIndex
    from hnsw_clj_b200.ultra_fast import gather_score

    rows, q = rng_rows(2000, 768, 60, unit=True), rng_rows(40, 768, 61, unit=True)
    r = np.random.default_rng(0)
    pq = r.integers(0, 40, 5000).astype(np.int32)
    pr = r.integers(0, 2000, 5000).astype(np.int32)
    with FlatIndex(rows, distance_fn=metric) as ix:
        got = gather_score(ix, q, pq, pr)
    assert same_bits(got, orc.gather_score(rows, q, pq, pr, orc.COSINE if metric == "cosine" else orc.L2))


# ---- top-k merge (ivf_flat.clj:291-294; partitioned_hnsw.clj:187-196) ----------------------------------
def test_topk_merge(hb):
    from hnsw_clj_b200 import _lib

    r = np.random.default_rng(5)
    nparts, nq, k = 4, 37, 10
    dist = np.sort(r.integers(0, 30, (nparts, nq, k)).astype(np.float64), axis=2)  # many ties
    ids = r.integers(0, 10**9, (nparts, nq, k)).astype(np.int64)
    dist[1, :, 7:] = np.inf
    ids[1, :, 7:] = -1
    out_i = np.empty((nq, k), np.int64)
    out_d = np.empty((nq, k), np.float64)
    _lib.check(_lib.lib().hb_topk_merge(dist.ctypes.data, ids.ctypes.data, nparts, nq, k, out_i.ctypes.data, out_d.ctypes.data))
    for qi in range(nq):
        cd = dist[:, qi, :].reshape(-1)
        ci = ids[:, qi, :].reshape(-1)
        order = np.argsort(cd, kind="stable")[:k]  # stable: ties by (part, position)
        assert out_d[qi].tolist() == cd[order].tolist()
        assert out_i[qi].tolist() == ci[order].tolist()


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_ivf_bf16_slabs_equal_oracle_on_the_rounded_values(mode):
    """north_star: "contiguous per-list fp32/bf16 slabs".  bf16 storage holds fp32-representable values, so the reference's
    arithmetic on the bf16-rounded rows (widened to double[]) is what the device must return: build (k-means++ seeds, 3 Lloyd
    rounds, final assignment) and search, ids and fp64 distance bits, in both modes."""
    import torch

    from hnsw_clj_b200 import _lib, ivf_flat

    _lib.check(_lib.lib().hb_init(0))
    r = np.random.default_rng(23)
    c = r.standard_normal((40, 128))
    rows32 = (c[r.integers(0, 40, 6000)] + 0.2 * r.standard_normal((6000, 128))).astype(np.float32)
    rows_bf = torch.from_numpy(rows32).to(torch.bfloat16)
    rows = rows_bf.to(torch.float32).numpy()  # the values the slab holds
    q = (c[r.integers(0, 40, 300)] + 0.2 * r.standard_normal((300, 128))).astype(np.float32)
    oc, oa = orc.kmeans(rows, 32, iters=3, seed=42)
    want_i, want_d = orc.ivf_search(rows, oc, oa, q, 10, 6)
    _lib.set_mode(_lib.MODE_FAST if mode == "fast" else _lib.MODE_EXACT)
    try:
        for data in (rows_bf, rows_bf.cuda()):  # host and device bf16 buffers
            with ivf_flat.build_index(data, num_partitions=32, max_iterations=3) as ix:
                assert ix.info()["dtype"] == _lib.BF16
                cents, asg = ix.export()
                assert (asg == oa).all() and (cents.view(np.int64) == oc.view(np.int64)).all()
                ids, d = ix.search_raw(q, 10, 6)
                assert ids.tolist() == want_i.tolist()
                assert (d.view(np.int64) == want_d.view(np.int64)).all()
                ids1, d1 = ix.search_raw(q[:1], 10, 6)  # small-batch path on a bf16 slab
                assert ids1.tolist() == want_i[:1].tolist() and (d1.view(np.int64) == want_d[:1].view(np.int64)).all()
    finally:
        _lib.set_mode(_lib.MODE_EXACT)
