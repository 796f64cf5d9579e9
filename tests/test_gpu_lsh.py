"""Hybrid LSH on the device (src/hnsw/ann/hash/hybrid_lsh.clj): hashing = hb_pairwise against the
java.util.Random(42) projections, bucket scan = hb_gather_score over the probed buckets' members, final sort + take k =
hb_topk_merge.  Bucket ids, result ids and distance bits equal the oracle's restatement of the Clojure."""
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


def same_bits(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool((a.view(np.int64) == b.view(np.int64)).all())


@pytest.fixture(scope="module")
def case():
    r = np.random.default_rng(11)
    c = r.standard_normal((12, 96))
    rows = (c[r.integers(0, 12, 9000)] + 0.35 * r.standard_normal((9000, 96))).astype(np.float32)
    rows[4000:4020] = rows[100:120]  # repeated vectors: equal distances from different ids, same buckets
    q = (c[r.integers(0, 12, 40)] + 0.35 * r.standard_normal((40, 96))).astype(np.float32)
    q[:5] = rows[100:105]
    return rows, q


def test_bucket_ids_equal_oracle(hb, case):
    from hnsw_clj_b200 import hybrid_lsh

    rows, q = case
    M = hybrid_lsh.projection_matrices(rows.shape[1])
    assert (hybrid_lsh.bucket_ids(rows, M) == orc.lsh_hash(rows, orc.lsh_matrices(rows.shape[1]))).all()
    assert (hybrid_lsh.bucket_ids(q.astype(np.float64), M) == orc.lsh_hash(q, M)).all()  # the reference's double[] queries


@pytest.mark.parametrize("probes,radius", [(2, 1), (4, 1), (6, 2), (8, 3), (8, 4), (9, 13)])
def test_multiprobe_search_equals_oracle(hb, case, probes, radius):
    from hnsw_clj_b200 import hybrid_lsh

    rows, q = case
    with hybrid_lsh.build_index(rows) as ix:
        ids, dist = hybrid_lsh.search_hybrid_multiprobe_raw(ix, q, 10, probes, radius)
        buckets = ix.buckets
    want_ids, want_d = orc.lsh_search(rows, orc.lsh_matrices(rows.shape[1]), buckets, q, 10, probes, radius, True, 2)
    assert ids.tolist() == want_ids.tolist()
    assert same_bits(dist, want_d)


@pytest.mark.parametrize("probes,mult,k", [(2, 3, 10), (2, 2, 10), (8, 3, 1), (1, 3, 50)])
def test_hybrid_search_equals_oracle(hb, case, probes, mult, k):
    """search-hybrid: the parallel branch keeps k*3 per bucket, the sequential one k*2 — neither cut reaches the result."""
    from hnsw_clj_b200 import hybrid_lsh

    rows, q = case
    with hybrid_lsh.build_index(rows) as ix:
        ids, dist = hybrid_lsh.search_hybrid_raw(ix, q, k, probes)
        buckets = ix.buckets
    want_ids, want_d = orc.lsh_search(rows, orc.lsh_matrices(rows.shape[1]), buckets, q, k, probes, 0, False, mult)
    assert ids.tolist() == want_ids.tolist()
    assert same_bits(dist, want_d)


def test_api_shapes_and_modes(hb, case):
    from hnsw_clj_b200 import hybrid_lsh

    rows, q = case
    data = [(f"vec_{i}", r.astype(np.float64)) for i, r in enumerate(rows[:3000])]
    with hybrid_lsh.build_index(data) as ix:
        info = hybrid_lsh.index_info(ix)
        assert info["type"] == "Hybrid LSH Index" and info["vectors"] == 3000 and info["hash-tables"] == 8
        assert info["buckets-per-table"] == 4096 and 0 < info["total-buckets"] <= 8 * 3000
        res = hybrid_lsh.search_knn(ix, rows[100].astype(np.float64), 5)
        assert res[0]["id"] == "vec_100" and abs(res[0]["distance"]) < 1e-12 and len(res) <= 5
        assert [r["distance"] for r in res] == sorted(r["distance"] for r in res)
        want_ids, want_d = orc.lsh_search(rows[:3000], orc.lsh_matrices(96), ix.buckets, rows[100:101], 5, 6, 2, True, 2)
        assert [r["id"] for r in res] == [f"vec_{i}" for i in want_ids[0] if i >= 0]
        for mode, (p, rad) in {"turbo": (2, 1), "fast": (4, 1), "balanced": (6, 2), "accurate": (8, 3), "precise": (8, 4)}.items():
            got = hybrid_lsh.search_batch(ix, q[:6], 7, mode)
            w_ids, w_d = orc.lsh_search(rows[:3000], orc.lsh_matrices(96), ix.buckets, q[:6], 7, p, rad, True, 2)
            for qi in range(6):
                assert [g["id"] for g in got[qi]] == [f"vec_{i}" for i in w_ids[qi] if i >= 0]
                assert [g["distance"] for g in got[qi]] == w_d[qi][: len(got[qi])].tolist()
        # a query whose buckets are empty in every probed table returns [] (the `when bucket` of :227-231)
        far = hybrid_lsh.search_hybrid(ix, -1000.0 * rows[100].astype(np.float64) + 3.0, 5, num_probes=1)
        assert isinstance(far, list)
