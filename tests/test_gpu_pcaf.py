"""SURVEY §8 a4 + f3: the float[] Vector-API distance functions (src/hnsw/simd.clj:18-115) and PCAF
(src/hnsw/ann/dimreduct/pcaf.clj) on the device vs the oracle's restatement.

Tolerance.  The reference's fp32 lane sums go through jdk.incubator.vector reduceLanes(ADD), whose lane order the JDK
leaves unspecified and whose width depends on the CPU (4 / 8 / 16 floats).  The north star allows 1e-5 relative for fp32:
the device is asserted (1) bit-identical to the oracle restated with the same lane width and left-to-right order, and
(2) within 1e-5 relative of the oracle restated with ANOTHER order (pairwise tree) and OTHER widths — i.e. of whatever a
JVM may compute."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
REL = 1e-5  # BASELINE.json north_star: "distances within 1e-5 relative for fp32"


@pytest.fixture(scope="module")
def hb():
    from hnsw_clj_b200 import _lib

    _lib.check(_lib.lib().hb_init(0))
    return _lib


@pytest.mark.parametrize("d", [768, 100, 37, 5])
@pytest.mark.parametrize("lanes", [4, 8, 16])
def test_lane_distances_equal_oracle_bits(hb, d, lanes):
    from hnsw_clj_b200 import simd

    r = np.random.default_rng(d * 31 + lanes)
    A = r.standard_normal((9, d)).astype(np.float32)
    B = r.standard_normal((131, d)).astype(np.float32)
    B[3] = 0.0  # (zero? magnitude) -> 1.0
    for metric in (hb.COSINE, hb.L2, hb.IP):
        got = simd.pairwise(A, B, metric, lanes)
        want = orc.simd_pairwise(A, B, metric, lanes)
        assert (got.view(np.int64) == want.view(np.int64)).all(), (metric, d, lanes)
    assert simd.pairwise(A, B, hb.COSINE, lanes)[0, 3] == 1.0


def test_lane_order_and_width_stay_within_the_fp32_bound(hb):
    from hnsw_clj_b200 import simd

    r = np.random.default_rng(5)
    A = r.standard_normal((4, 768)).astype(np.float32)
    B = (A[r.integers(0, 4, 64)] + 0.3 * r.standard_normal((64, 768))).astype(np.float32)  # distances from ~0.04 up
    got = simd.pairwise(A, B, hb.COSINE, 8)
    for lanes in (4, 8, 16):
        tree = np.array([[orc.simd_cosine(a, b, lanes, tree=True) for b in B] for a in A])
        assert np.all(np.abs(got - tree) <= REL * np.abs(tree))
    # and of the double[] path the float[] functions stand in for (cosine-distance-direct, simd.clj:129-147)
    direct = np.array([[orc.cosine_distance_direct(a.astype(np.float64), b.astype(np.float64)) for b in B] for a in A])
    assert np.all(np.abs(got - direct) <= REL * np.abs(direct))
    # scalar mirrors
    assert simd.cosine_distance(A[0], B[0]) == got[0, 0]
    assert simd.dot_product(A[0], B[0]) == orc.simd_dot(A[0], B[0], 8)
    assert simd.euclidean_distance(A[1], B[2]) == orc.simd_euclidean(A[1], B[2], 8)
    assert simd.cosine_distance_direct(A[0], B[0]) == direct[0, 0]


def test_known_answers_of_the_reference_tests(hb):
    """test/hnsw/core_test.clj:9-31 through the float[] functions."""
    from hnsw_clj_b200 import simd

    assert simd.euclidean_distance([0, 0], [3, 4]) == 5.0
    assert abs(simd.euclidean_distance([1, 2, 3], [4, 5, 6]) - 5.196152422706632) < 1e-5
    assert abs(simd.cosine_distance([1, 2, 3], [4, 5, 6]) - 0.0253) < 1e-2
    assert abs(simd.cosine_distance([1, 0], [-1, 0]) - 2.0) < 1e-3
    assert simd.cosine_distance([1, 2, 3], [1, 2, 3]) < 1e-3


def _pcaf_data(n=4000, d=768, nq=40, seed=3):
    r = np.random.default_rng(seed)
    c = r.standard_normal((40, d))
    rows = (c[r.integers(0, 40, n)] + 0.3 * r.standard_normal((n, d))).astype(np.float32)
    q = (rows[r.integers(0, n, nq)] + 0.05 * r.standard_normal((nq, d))).astype(np.float32)
    return rows, q


def test_pcaf_projection_and_search_equal_oracle(hb):
    from hnsw_clj_b200 import pcaf

    rows, q = _pcaf_data()
    with pcaf.build_index(rows, n_components=100, k_filter=32) as ix:
        m = orc.pcaf_matrix(768, 100)
        assert (ix.projection.view(np.int32) == m.view(np.int32)).all()
        low = pcaf.project_vectors(ix.projection, rows[:300])
        assert (low.view(np.int32) == orc.pcaf_project(m, rows[:300]).view(np.int32)).all()
        for k, mode, kf in ((10, None, 32), (10, "precise", 64), (5, "turbo", 16), (20, "balanced", 32)):
            ids, dist = ix.search_raw(q, k, pcaf._k_filter(ix, mode))
            want_i, want_d = orc.pcaf_search(rows, q, k, 100, kf)
            assert ids.tolist() == want_i.tolist(), (k, mode)
            assert (dist.view(np.int64) == want_d.view(np.int64)).all()
        one = pcaf.search_knn(ix, q[0].astype(np.float64), 10, ":balanced")  # double[] query, keyword mode
        want_i, want_d = orc.pcaf_search(rows, q[:1], 10, 100, 32)
        assert [r["id"] for r in one] == want_i[0].tolist() and [r["distance"] for r in one] == want_d[0].tolist()
        info = pcaf.index_info(ix)
        assert info["reduced-dim"] == 100 and info["vectors"] == len(rows) and info["k-filter"] == 32
        # recall vs the exact flat search is what the reference's algorithm gives (only min(k-filter, 3k) = 30 candidates of a
        # 100-dimensional projection are re-ranked, pcaf.clj:229-230): the same number as the oracle's, and well above chance
        exact_i, _ = orc.exact_knn(rows, q, 10)
        ids, _ = ix.search_raw(q, 10, 64)
        want_i, _ = orc.pcaf_search(rows, q, 10, 100, 64)
        assert orc.recall(ids, exact_i) == orc.recall(want_i, exact_i) >= 0.4
        assert (ids[:, 0] == exact_i[:, 0]).mean() >= 0.9  # the nearest neighbour survives the projection
        # k-filter below k pads (the reference returns fewer than k results, pcaf.clj:229-253)
        ids, dist = ix.search_raw(q[:2], 10, 4)
        assert (ids[:, 4:] == -1).all() and np.isinf(dist[:, 4:]).all() and (ids[:, :4] >= 0).all()


def test_pcaf_through_the_api(hb):
    from hnsw_clj_b200 import api

    rows, q = _pcaf_data(n=1500, nq=6)
    data = [(f"vec_{i}", rows[i].astype(np.float64)) for i in range(len(rows))]  # ["vec_0" double[]] pairs
    ix = api.index(data, index_type="pcaf", n_components=64, k_filter=24)
    try:
        res = api.search_batch_(ix, q, 5, mode="accurate")
        want_i, want_d = orc.pcaf_search(rows, q, 5, 64, 48)
        assert [[r["id"] for r in one] for one in res] == [[f"vec_{i}" for i in row] for row in want_i.tolist()]
        assert api.index_type_(ix) == "PCAF" and api.index_info_(ix)["type"].startswith("P-HNSW")
        assert api.search(ix, q[0], 5, mode="accurate") == res[0]
    finally:
        ix.close()
