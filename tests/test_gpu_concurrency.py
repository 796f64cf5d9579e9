"""Concurrent callers on one shared index (test/hnsw/core_test.clj:112-121, src/hnsw/wip/31k-multithread-sb.clj:127-134):
hb_search is safe to call from many threads, and the MicroBatcher answers concurrent single-query calls in device
batches with the results lone calls return."""
import threading

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case():
    from hnsw_clj_b200 import _lib, ivf_flat

    _lib.check(_lib.lib().hb_init(0))
    r = np.random.default_rng(21)
    c = r.standard_normal((30, 64))
    rows = (c[r.integers(0, 30, 6000)] + 0.1 * r.standard_normal((6000, 64))).astype(np.float32)
    q = (c[r.integers(0, 30, 400)] + 0.1 * r.standard_normal((400, 64))).astype(np.float32)
    ix = ivf_flat.build_index(rows, num_partitions=24, max_iterations=3)
    cents, asg = ix.export()
    want = orc.ivf_search(rows, cents, asg, q, 10, 4)  # :balanced = 4 probes
    yield ix, q, want
    ix.close()


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_fifty_threads_share_one_index(case, mode):
    from hnsw_clj_b200 import _lib

    ix, q, (want_ids, want_d) = case
    _lib.set_mode(_lib.MODE_FAST if mode == "fast" else _lib.MODE_EXACT)
    out, errs = {}, []

    def worker(t):
        try:
            for j in range(t, len(q), 50):
                ids, d = ix.search_raw(q[j], 10, 4)  # one query per call, like the reference's callers
                out[j] = (ids[0].tolist(), d[0].view(np.int64).tolist())
        except Exception as e:  # pragma: no cover
            errs.append(e)

    try:
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(50)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    finally:
        _lib.set_mode(_lib.MODE_EXACT)
    assert not errs and len(out) == len(q)
    for j in range(len(q)):
        assert out[j] == (want_ids[j].tolist(), want_d[j].view(np.int64).tolist())


def test_microbatcher_matches_lone_calls(case):
    from hnsw_clj_b200 import ivf_flat
    from hnsw_clj_b200.parallel_search import MicroBatcher, parallel_search_futures

    ix, q, (want_ids, want_d) = case
    lone = [ivf_flat.search_knn(ix, q[j], 10) for j in range(40)]
    assert [r["id"] for r in lone[3]] == want_ids[3].tolist()
    got = [None] * len(q)
    with MicroBatcher(ix, 10, max_batch=64, max_wait_s=0.002) as mb:
        def worker(t):
            for j in range(t, len(q), 50):
                got[j] = mb.search(q[j])

        ts = [threading.Thread(target=worker, args=(t,)) for t in range(50)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        # as the per-query search-fn of parallel-search-futures (helper/parallel_search.clj:15-49)
        again = parallel_search_futures(ix, q[:5], 10, search_fn=lambda index, query, k: mb.search(query))
        assert mb.served == len(q) + 5 and mb.batches < mb.served  # concurrent calls were coalesced
    assert got[:40] == lone and again == lone[:5]
    for j in range(len(q)):
        assert [r["id"] for r in got[j]] == want_ids[j].tolist()
        assert [r["distance"] for r in got[j]] == want_d[j].tolist()
