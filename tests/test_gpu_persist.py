"""save-index / load-index of the device layout (hb_index_save / hb_index_load; SURVEY §8 f2, the counterpart of
src/hnsw/helper/index_io.clj:10-80): a loaded index answers with the same ids and fp64 distance bits as the one
saved, in both search modes, without re-clustering; damaged files are rejected."""
import os

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


def clustered(n, d, seed, nc=12):
    r = np.random.default_rng(seed)
    c = r.standard_normal((nc, d))
    return (c[r.integers(0, nc, n)] + 0.1 * r.standard_normal((n, d))).astype(np.float32)


def same(a, b):
    return a[0].tolist() == b[0].tolist() and bool((a[1].view(np.int64) == b[1].view(np.int64)).all())


def test_ivf_round_trip_matches_oracle(hb, tmp_path):
    from hnsw_clj_b200 import _lib, index_io, ivf_flat

    rows, q = clustered(5000, 96, 1), clustered(300, 96, 2)
    data = [(f"vec_{i}", rows[i]) for i in range(len(rows))]  # String ids, test/data_generator.clj:84-87
    path = str(tmp_path / "ivf.hbix")
    ix = ivf_flat.build_index(data, num_partitions=20, max_iterations=4)
    before = ix.search_raw(q, 10, 6)
    maps_before = ivf_flat.search_knn(ix, q[0], 5, "accurate")
    cents, asg = ix.export()
    assert index_io.save_index(ix, path) is ix
    ix.close()
    ix2 = index_io.load_index(path)
    assert isinstance(ix2, ivf_flat.IVFFlatIndex) and ix2.num_partitions == 20 and ix2.ids[3] == "vec_3"
    info = ix2.info()
    assert (info["n"], info["dim"], info["nlist"], info["type"]) == (5000, 96, 20, _lib.INDEX_IVF_FLAT)
    c2, a2 = ix2.export()
    assert (c2.view(np.int64) == cents.view(np.int64)).all() and a2.tolist() == asg.tolist()
    assert same(ix2.search_raw(q, 10, 6), before)
    assert ivf_flat.search_knn(ix2, q[0], 5, "accurate") == maps_before
    want = orc.ivf_search(rows, cents, asg, q, 10, 6)
    assert same(ix2.search_raw(q, 10, 6), want)
    _lib.set_mode(_lib.MODE_FAST)  # digit images are rebuilt from the loaded slab
    try:
        assert same(ix2.search_raw(q, 10, 6), want)
    finally:
        _lib.set_mode(_lib.MODE_EXACT)
    ix2.close()


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "ip"])
def test_flat_round_trip(hb, tmp_path, metric):
    from hnsw_clj_b200 import index_io
    from hnsw_clj_b200.flat import FlatIndex

    rows, q = clustered(3000, 64, 3), clustered(100, 64, 4)
    path = str(tmp_path / "flat.hbix")
    with FlatIndex(rows, distance_fn=metric) as fx:
        before = fx.search_raw(q, 7)
        index_io.save_index(fx, path)
    with index_io.load_index(path) as fx2:
        assert isinstance(fx2, FlatIndex) and fx2.ids is None
        assert same(fx2.search_raw(q, 7), before)
    assert not os.path.exists(path + ".ids.json")


def test_hnsw_round_trip(hb, tmp_path):
    from hnsw_clj_b200 import index_io
    from hnsw_clj_b200.ultra_fast import HnswIndex

    r = np.random.default_rng(5)
    rows = r.standard_normal((2000, 48)).astype(np.float32)
    g = orc.Hnsw(rows, M=8, ef_construction=50, level_seed=42)
    adjacency = [g.export_level(l) for l in range(g.max_level + 1)]
    q = rows[:64] + 0.05 * r.standard_normal((64, 48)).astype(np.float32)
    want = g.search(q, 10, 64)
    path = str(tmp_path / "hnsw.hbix")
    with HnswIndex(rows, g.levels(), g.entry, adjacency) as ix:
        assert same(ix.search_raw(q, 10, 64), want)
        index_io.save_index(ix, path)
    with index_io.load_index(path) as ix2:
        assert isinstance(ix2, HnswIndex) and ix2.info()["max_level"] == g.max_level
        assert same(ix2.search_raw(q, 10, 64), want)


def test_empty_flat_round_trip(hb, tmp_path):
    from hnsw_clj_b200 import index_io
    from hnsw_clj_b200.flat import FlatIndex

    path = str(tmp_path / "empty.hbix")
    with FlatIndex(np.zeros((0, 16), dtype=np.float32)) as fx:
        index_io.save_index(fx, path)
    with index_io.load_index(path) as fx2:
        assert fx2.info()["n"] == 0
        assert fx2.search_knn(np.ones(16, dtype=np.float32), 3) == []  # empty index -> [], ultra_fast.clj:349-351


def test_missing_truncated_and_foreign_files(hb, tmp_path):
    from hnsw_clj_b200 import HbInvalid, index_io
    from hnsw_clj_b200.flat import FlatIndex

    assert index_io.load_index(str(tmp_path / "nope.hbix")) is None  # nil, index_io.clj:78-80
    path = str(tmp_path / "flat.hbix")
    with FlatIndex(clustered(500, 32, 6)) as fx:
        index_io.save_index(fx, path)
    blob = open(path, "rb").read()
    cut = str(tmp_path / "cut.hbix")
    open(cut, "wb").write(blob[: len(blob) // 2])
    with pytest.raises(HbInvalid):
        index_io.load_index(cut)
    junk = str(tmp_path / "junk.hbix")
    open(junk, "wb").write(b"{:nodes {}}" * 20)
    with pytest.raises(HbInvalid):
        index_io.load_index(junk)
    with pytest.raises(HbInvalid):  # unwritable destination
        with FlatIndex(clustered(10, 8, 7)) as fx:
            index_io.save_index(fx, str(tmp_path / "no_such_dir" / "x.hbix"))
