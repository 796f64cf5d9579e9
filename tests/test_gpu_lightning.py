"""GPU parity of the Lightning consumer of the scan kernels (src/hnsw/ann/partition/lightning.clj):
hb_lightning_build (d_i-weighted k-means++ walk, nearest-seed partitions, mean centroids, zero vector for an empty
partition) and search-lightning's percentage-based probing, against the oracle restatement."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hnsw_clj_b200 as pkg
    from hnsw_clj_b200 import _lib

    _lib.check(_lib.lib().hb_init(0))
    yield pkg
    _lib.set_mode(_lib.MODE_EXACT)


def clustered(n, d, seed, nc=20):
    r = np.random.default_rng(seed)
    c = r.standard_normal((nc, d))
    return (c[r.integers(0, nc, n)] + 0.1 * r.standard_normal((n, d))).astype(np.float32)


def same(a, b):
    return a[0].tolist() == b[0].tolist() and bool((a[1].view(np.int64) == b[1].view(np.int64)).all())


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("n,d,parts,metric", [(3000, 48, 24, "cosine"), (5000, 96, 40, "cosine"), (2000, 32, 16, "euclidean"),
                                              (9000, 64, 300, "cosine")])
def test_lightning_build_and_search_match_oracle(hb, mode, n, d, parts, metric):
    from hnsw_clj_b200 import _lib, lightning

    rows, q = clustered(n, d, n + parts), clustered(150, d, n + parts + 1)
    om = orc.COSINE if metric == "cosine" else orc.L2
    want_c, want_a = orc.lightning_build(rows, parts, metric=om, seed=42)
    _lib.set_mode(_lib.MODE_FAST if mode == "fast" else _lib.MODE_EXACT)
    try:
        ix = lightning.build_index(rows, num_partitions=parts, distance_fn=metric, smart_partition=True)
        cents, asg = ix.export()
        assert asg.tolist() == want_a.tolist()
        assert (cents.view(np.int64) == want_c.view(np.int64)).all()
        for m in ("balanced", "accurate", "precise"):
            nprobe = lightning.num_partitions_to_search(parts, mode=m)
            got = ix.search_raw(q, 10, nprobe)
            want = orc.ivf_search(rows, want_c, want_a, q, 10, nprobe, coarse_metric=om)
            assert same(got, want), m
        maps = lightning.search_knn(ix, q[0], 5, "accurate")
        assert [r["id"] for r in maps] == orc.ivf_search(rows, want_c, want_a, q[:1], 5,
                                                         lightning.num_partitions_to_search(parts, "accurate"),
                                                         coarse_metric=om)[0][0].tolist()
        assert lightning.index_info(ix)["partitions"] == parts
        ix.close()
    finally:
        _lib.set_mode(_lib.MODE_EXACT)


def test_lightning_empty_partition_gets_zero_centroid(hb):
    """More partitions than distinct rows: the walk re-picks duplicates, the later twin's partition stays empty and its
    routing centroid is the zero vector (lightning.clj:122-126; cosine distance to it is 1.0 by the guard)."""
    from hnsw_clj_b200 import lightning

    base = clustered(6, 16, 3)
    rows = np.concatenate([base] * 5)
    want_c, want_a = orc.lightning_build(rows, 12)
    ix = lightning.build_index(rows, num_partitions=12, smart_partition=True)
    cents, asg = ix.export()
    assert asg.tolist() == want_a.tolist() and (cents.view(np.int64) == want_c.view(np.int64)).all()
    assert (np.bincount(asg, minlength=12) == 0).any() and (cents[np.bincount(asg, minlength=12) == 0] == 0).all()
    q = clustered(20, 16, 4)
    assert same(ix.search_raw(q, 4, 6), orc.ivf_search(rows, want_c, want_a, q, 4, 6))
    ix.close()


def test_lightning_shuffle_partition_searches_like_its_partition(hb):
    """Default build (:smart-partition? false, lightning.clj:132-137): a seeded shuffle split; the search over it equals
    the oracle's search over the same partition."""
    from hnsw_clj_b200 import lightning

    rows, q = clustered(2500, 40, 9), clustered(60, 40, 10)
    ix = lightning.build_index(rows, num_partitions=32)
    cents, asg = ix.export()
    sizes = np.bincount(asg)
    assert sizes.max() == -(-2500 // 32) and len(sizes) == ix.num_partitions
    nprobe = lightning.num_partitions_to_search(ix.num_partitions, search_percent=0.5)
    assert same(ix.search_raw(q, 10, nprobe), orc.ivf_search(rows, cents, asg, q, 10, nprobe))
    assert len(lightning.search_knn(ix, q[0], 3, 0.2)) == 3
    ix.close()
