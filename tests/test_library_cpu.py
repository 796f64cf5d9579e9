"""CPU-side checks of the boundary: the C-ABI library builds/loads, exports every symbol include/hnswb200.h
declares, refuses to compute without a device (no CPU fallback), and the host mirror's pure logic."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hnsw_clj_b200.build import build_library

    build_library()
    from hnsw_clj_b200 import _lib

    return _lib


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "hnswb200.h")).read()
    declared = set(re.findall(r"HB_API\s+[\w\s\*]+?\b(hb_\w+)\s*\(", header))
    assert len(declared) >= 24
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert L.hb_version() >= 100


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from hnsw_clj_b200 import simd_optimized as so
    from hnsw_clj_b200.flat import FlatIndex

    with pytest.raises(lib.HbError) as e:
        so.cosine_distance([1, 2, 3], [4, 5, 6])
    assert e.value.status == lib.ERR_NO_DEVICE
    with pytest.raises(lib.HbError):
        FlatIndex(np.eye(4, dtype=np.float32))


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "hnsw_clj_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # comments may mention "oracle-built partitions"; nothing may import / load / call it
                for needle in ("import oracle", "from oracle", "libhnsw_oracle", "orc_", "oracle/"):
                    assert needle not in text, (f, needle)


def test_host_helpers(lib):
    from hnsw_clj_b200 import flat, index, ivf_flat

    assert index.metric_code("cosine") == lib.COSINE and index.metric_code(":euclidean") == lib.L2
    with pytest.raises(ValueError):
        index.metric_code("manhattan")
    ids, rows = index.split_data([("a", [1.0, 2.0]), ("b", [3.0, 4.0])])
    assert ids == ["a", "b"] and rows.shape == (2, 2) and rows.dtype == np.float64
    ids, rows = index.split_data(np.ones((3, 2), np.float32))
    assert ids is None and rows.dtype == np.float32
    maps = index.results_to_maps(np.array([[1, 0, -1]]), np.array([[0.1, 0.2, np.inf]]), ["x", "y"])
    assert maps == [[{"id": "y", "distance": 0.1}, {"id": "x", "distance": 0.2}]]
    assert ivf_flat._num_probes("balanced", None) == 4 and ivf_flat._num_probes(":precise", None) == 12
    assert ivf_flat._num_probes("custom", 32) == 32 and ivf_flat._num_probes("custom", None) == 4
    assert flat.calc_recall([{"id": 1}, {"id": 2}], [{"id": 2}, {"id": 3}]) == 0.5
    assert flat.recall_at_k(np.array([[1, 2, -1]]), np.array([[2, 3, -1]])) == 0.5


def test_lightning_probe_counts_follow_the_reference_tables():
    """(max 1 (int (* num-partitions percent))) with the per-size mode tables of src/hnsw/ann/partition/lightning.clj:193-262."""
    from hnsw_clj_b200.lightning import num_partitions_to_search as nps

    assert nps(64, "balanced") == 6 and nps(100, "precise") == 25 and nps(64, "turbo") == 1
    assert nps(32, "balanced") == 4 and nps(48, "accurate") == 12
    assert nps(24, "balanced") == 4 and nps(24, "precise") == 12 and nps(24, ":fast") == 2
    assert nps(16, "balanced") == 4 and nps(8, "turbo") == 1 and nps(20, "precise") == 12
    assert nps(24) == 4 and nps(16) == 4 and nps(32) == 4 and nps(64) == 6 and nps(200) == 16  # dynamic default (:253-260)
    assert nps(40, search_percent=0.01) == 1 and nps(40, search_percent=0.5) == 20


def test_data_loader_reads_the_reference_json_layout(tmp_path):
    """src/hnsw/helper/data_loader.clj:7-45: {"verses": [{"id", "text", "embedding"}]} -> ids + one matrix; fp32-exact
    and unit-norm detection; an unreadable file gives None (nil)."""
    import json

    import numpy as np

    from hnsw_clj_b200 import data_loader
    from hnsw_clj_b200.index import split_data

    r = np.random.default_rng(0)
    emb = r.standard_normal((7, 12)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    verses = [{"id": f"Gen_1:{i + 1}", "text": f"verse {i}", "embedding": [float(x) for x in emb[i].astype(np.float32)]}
              for i in range(7)]
    path = tmp_path / "bible_embeddings.json"
    path.write_text(json.dumps({"verses": verses}))
    got = data_loader.load_bible_vectors(str(path))
    m = got["vectors"]["matrix"]
    assert m.dtype == np.float32 and (m == emb.astype(np.float32)).all()
    assert got["metadata"]["fp32-exact"] and got["metadata"]["unit-norm"]
    assert got["metadata"]["count"] == 7 and got["metadata"]["dimension"] == 12
    assert got["text-map"]["Gen_1:3"] == "verse 2"
    ids, rows = split_data(data_loader.as_data(got))
    assert ids[0] == "Gen_1:1" and rows.ctypes.data == m.ctypes.data  # zero-copy hand-over to build_index
    verses[2]["embedding"][0] = 0.1  # not fp32-representable, not unit norm
    path.write_text(json.dumps({"verses": verses}))
    got = data_loader.load_bible_vectors(str(path))
    assert got["vectors"]["matrix"].dtype == np.float64 and not got["metadata"]["fp32-exact"]
    assert data_loader.load_bible_vectors(str(tmp_path / "missing.json")) is None
