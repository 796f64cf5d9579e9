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


def test_corrupt_contents_are_rejected_not_dereferenced(hb, tmp_path):
    """hb_index_load must validate what later drives device addressing (list offsets, row ids, assignments,
    adjacency ids) and answer HB_ERR_INVALID — never an out-of-bounds read on the device.  Lightning flavour survives a
    save / load round trip (index_io keeps the host-side class)."""
    import struct

    from hnsw_clj_b200 import _lib, index_io, ivf_flat, lightning

    rows = clustered(3000, 32, 4)
    n = len(rows)
    path = str(tmp_path / "ivf.hbix")
    with ivf_flat.build_index(rows, num_partitions=8, max_iterations=2) as ix:
        index_io.save_index(ix, path)
    blob = bytearray(open(path, "rb").read())
    # layout tail of an IVF-FLAT file: ... [hdr16][list_rows n x int64][hdr16][assign n x int32]
    off_assign = len(blob) - n * 4
    off_rows = off_assign - 16 - n * 8
    for name, (off, fmt, bad) in {"assignment": (off_assign + 40, "<i", 10 ** 6), "row id": (off_rows + 80, "<q", -5),
                                  "row id high": (off_rows + 8, "<q", n)}.items():
        b2 = bytearray(blob)
        struct.pack_into(fmt, b2, off, bad)
        p2 = str(tmp_path / "bad.hbix")
        open(p2, "wb").write(b2)
        with pytest.raises(ValueError):
            index_io.load_index(p2)
    ok = index_io.load_index(path)  # the untouched file still loads, and the device is healthy after the rejections
    assert ok.search_raw(rows[:4], 3, 2)[0][:, 0].tolist() == [0, 1, 2, 3]
    ok.close()
    with lightning.build_index(rows, num_partitions=16, smart_partition=True) as lx:
        want = lightning.search_knn(lx, rows[5], 5, mode="balanced")
        index_io.save_index(lx, path)
    back = index_io.load_index(path)
    assert isinstance(back, lightning.LightningIndex) and lightning.search_knn(back, rows[5], 5, mode="balanced") == want
    back.close()


def test_out_of_range_indices_are_invalid_arguments(hb):
    """caller-supplied indices are range-checked (the reference throws on an unknown id)."""
    import ctypes as C

    from hnsw_clj_b200 import _lib, ivf_flat, ultra_fast
    from hnsw_clj_b200.flat import FlatIndex

    rows = clustered(500, 16, 5)
    with FlatIndex(rows) as fx:
        good = ultra_fast.gather_score(fx, rows[:2], np.array([0, 1], np.int32), np.array([3, 4], np.int32))
        assert good.shape == (2,)
        for pq, pr in (([0, 2], [3, 4]), ([0, 1], [3, 500]), ([0, -1], [3, 4])):
            with pytest.raises(ValueError):
                ultra_fast.gather_score(fx, rows[:2], np.array(pq, np.int32), np.array(pr, np.int32))
        assert ultra_fast.gather_score(fx, rows[:2], np.array([0, 1], np.int32), np.array([3, 4], np.int32)).tolist() == good.tolist()
    cents = np.zeros((4, 16))
    asg = np.zeros(500, np.int32)
    asg[7] = 4
    with pytest.raises(ValueError):
        ivf_flat.import_index(rows, cents, asg)
    with pytest.raises(ValueError):
        ivf_flat.partition_vectors_kmeans(rows, 4, seed_rows=[0, 1, 2, 500])
    levels = np.zeros(500, np.int32)
    off = np.arange(501, dtype=np.int64)
    nbr = np.arange(500, dtype=np.int32)[::-1].copy()
    nbr[9] = 777
    with pytest.raises(ValueError):
        ultra_fast.HnswIndex(rows, levels, 0, [(off, nbr)])
