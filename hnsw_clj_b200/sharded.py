"""Row-sharded multi-GPU search: one process per GPU (torch.distributed), each rank holds the rows of the
IVF lists it owns (list l lives on rank l mod G) next to the replicated centroids, produces a local top-k,
and a small all-gather plus the merge kernel (hb_topk_merge) gives the global top-k on every rank.

This is the reference's own scale-out model — independent sub-indexes searched in parallel and merged by
(sort-by :distance) + (take k) (src/hnsw/ann/partition/partitioned_hnsw.clj:149-196) — applied to the lists
of ONE global IVF-FLAT index, so results are identical to the single-GPU results.
torch.distributed is plumbing only: the collective carries G*nq*k (distance, id) pairs.
"""
from __future__ import annotations

import numpy as np

from . import _lib as hb
from . import ivf_flat


def list_owner(nlist: int, world: int) -> np.ndarray:
    return (np.arange(nlist) % world).astype(np.int32)


class ShardedIVFFlat:
    """The shard of a global IVF-FLAT index owned by this rank."""

    def __init__(self, rows, centroids, assignments, rank: int, world: int, distance_fn="cosine", group=None):
        self.rank, self.world, self.group = rank, world, group
        asg = np.asarray(assignments, dtype=np.int32)
        cents = np.ascontiguousarray(centroids, dtype=np.float64)
        mine = np.nonzero(list_owner(cents.shape[0], world)[asg] == rank)[0]
        self.global_ids = mine.astype(np.int64)  # local row -> global row, ascending (keeps list order)
        if hb._is_torch(rows):
            import torch

            local_rows = rows[torch.from_numpy(mine).to(rows.device)]
        else:
            local_rows = np.ascontiguousarray(np.asarray(rows)[mine])
        self.n_local = int(mine.shape[0])
        # lists owned by other ranks are simply empty here; the coarse quantiser still ranks ALL centroids,
        # so every rank derives the same probe lists
        self.index = ivf_flat.import_index(local_rows, cents, asg[mine], distance_fn) if self.n_local else None
        self._gid_dev = None

    def local_search(self, queries, k: int, num_probes: int):
        """Local top-k with GLOBAL row ids: (ids [nq,k] int64, dist [nq,k] fp64), numpy or torch like `queries`."""
        if hb._is_torch(queries) and queries.is_cuda:
            import torch

            nq = queries.shape[0]
            ids = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
            dist = torch.empty((nq, k), dtype=torch.float64, device=queries.device)
            if self.index is None:
                ids.fill_(-1)
                dist.fill_(float("inf"))
                return ids, dist
            self.index.search_raw(queries, k, num_probes, out_ids=ids, out_dist=dist)
            if self._gid_dev is None:
                self._gid_dev = torch.from_numpy(self.global_ids).to(queries.device)
            valid = ids >= 0
            ids = torch.where(valid, self._gid_dev[ids.clamp_min(0)], ids)
            return ids, dist
        q = hb.as_matrix(queries, allow=(hb.F32, hb.F64))
        if self.index is None:
            return (np.full((q.shape[0], k), -1, np.int64), np.full((q.shape[0], k), np.inf))
        ids, dist = self.index.search_raw(q, k, num_probes)
        ids = np.where(ids >= 0, self.global_ids[np.maximum(ids, 0)], -1)
        return ids, dist

    def search(self, queries, k: int, num_probes: int):
        """Global top-k on every rank: local search -> all-gather -> merge kernel."""
        ids, dist = self.local_search(queries, k, num_probes)
        return all_gather_merge(ids, dist, self.world, self.group)

    def close(self):
        if self.index is not None:
            self.index.close()
            self.index = None


def row_range(n: int, rank: int, world: int):
    """Contiguous row block of `rank`: [g*N/G, (g+1)*N/G), so ascending global ids = (rank, local id) order."""
    return (n * rank) // world, (n * (rank + 1)) // world


class ShardedFlat:
    """Row-sharded exact flat search (BASELINE configs[2]): rank g holds rows [g*N/G, (g+1)*N/G) (+ norms), queries are
    replicated, every rank produces a local top-k with global row ids, one all-gather of G*nq*k (distance, id) pairs and
    the merge kernel give the global top-k.  Ties across shards resolve by rank = by global row index, like the
    single-GPU stable sort (src/hnsw/bench.clj:72-84)."""

    def __init__(self, local_rows, first_row: int, rank: int, world: int, distance_fn="cosine", group=None):
        from .flat import FlatIndex

        self.rank, self.world, self.group, self.first_row = rank, world, group, int(first_row)
        self.index = FlatIndex(local_rows, distance_fn) if len(local_rows) else None

    def local_search(self, queries, k: int):
        if hb._is_torch(queries) and queries.is_cuda:
            import torch

            nq = queries.shape[0]
            ids = torch.full((nq, k), -1, dtype=torch.int64, device=queries.device)
            dist = torch.full((nq, k), float("inf"), dtype=torch.float64, device=queries.device)
            if self.index is not None:
                self.index.search_raw(queries, k, out_ids=ids, out_dist=dist)
                ids = torch.where(ids >= 0, ids + self.first_row, ids)
            return ids, dist
        q = hb.as_matrix(queries, allow=(hb.F32, hb.F64))
        if self.index is None:
            return np.full((q.shape[0], k), -1, np.int64), np.full((q.shape[0], k), np.inf)
        ids, dist = self.index.search_raw(q, k)
        return np.where(ids >= 0, ids + self.first_row, -1), dist

    def search(self, queries, k: int):
        ids, dist = self.local_search(queries, k)
        return all_gather_merge(ids, dist, self.world, self.group)

    def close(self):
        if self.index is not None:
            self.index.close()
            self.index = None


def all_gather_merge(ids, dist, world: int, group=None):
    """All-gather the per-rank (dist, id) blocks and merge them: ties by (rank, position) as in the stable
    sort of the concatenation, src/hnsw/ann/partition/ivf_flat.clj:291-294."""
    import torch
    import torch.distributed as dist_

    if world == 1:
        return ids, dist
    was_numpy = not hb._is_torch(ids)
    if was_numpy:
        ids, dist = torch.from_numpy(np.ascontiguousarray(ids)), torch.from_numpy(np.ascontiguousarray(dist))
    nq, k = ids.shape
    all_ids = torch.empty((world * nq, k), dtype=torch.int64, device=ids.device)
    all_dist = torch.empty((world * nq, k), dtype=torch.float64, device=ids.device)
    dist_.all_gather_into_tensor(all_ids, ids.contiguous(), group=group)
    dist_.all_gather_into_tensor(all_dist, dist.contiguous(), group=group)
    all_ids, all_dist = all_ids.view(world, nq, k), all_dist.view(world, nq, k)
    out_ids = torch.empty_like(ids)
    out_dist = torch.empty_like(dist)
    if ids.is_cuda:
        torch.cuda.current_stream().synchronize()  # NCCL ran on torch's stream; the library launches on its own
        hb.check(hb.lib().hb_topk_merge(all_dist.data_ptr(), all_ids.data_ptr(), world, nq, k, out_ids.data_ptr(),
                                        out_dist.data_ptr()))
    else:
        out_ids, out_dist = merge_host(all_ids.numpy(), all_dist.numpy())
        out_ids, out_dist = torch.from_numpy(out_ids), torch.from_numpy(out_dist)
    if was_numpy:
        return out_ids.numpy(), out_dist.numpy()
    return out_ids, out_dist


def merge_host(all_ids: np.ndarray, all_dist: np.ndarray):
    """Host statement of the merge rule (used on the gloo/CPU path of the tests and as documentation of
    hb_topk_merge): stable sort of the concatenation in rank order, take k."""
    world, nq, k = all_ids.shape
    cd = np.transpose(all_dist, (1, 0, 2)).reshape(nq, world * k)
    ci = np.transpose(all_ids, (1, 0, 2)).reshape(nq, world * k)
    order = np.argsort(cd, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(ci, order, 1), np.take_along_axis(cd, order, 1)


def sharded_kmeans_update(rows, assignments, centroids, group=None):
    """One Lloyd update with rows sharded across ranks: per-rank fp64 partial sums + counts (device kernel),
    all-reduce(sum), divide; an empty cluster keeps its centroid (ivf_flat.clj:66-77,112-116).  The all-reduce
    regroups the row-order sum, so centroids can differ from the single-GPU ones in the last ulp."""
    import torch
    import torch.distributed as dist_

    nlist = centroids.shape[0]
    sums, cnt = ivf_flat.partial_sums(rows, assignments, nlist)
    dev = rows.device if hb._is_torch(rows) and rows.is_cuda else "cpu"
    ts, tc = torch.from_numpy(sums).to(dev), torch.from_numpy(cnt).to(dev)
    dist_.all_reduce(ts, group=group)
    dist_.all_reduce(tc, group=group)
    ts, tc = ts.cpu().numpy(), tc.cpu().numpy()
    out = np.array(centroids, dtype=np.float64, copy=True)
    nz = tc > 0
    out[nz] = ts[nz] / tc[nz, None].astype(np.float64)
    return out


def lloyd_round_device(rows, centroids, assign_out, group=None, world: int = 1):
    """One Lloyd round of partition-vectors-kmeans (ivf_flat.clj:100-118) on device tensors, rows sharded across ranks:
    assign this rank's rows (hb_kmeans_assign; FAST mode runs the tensor-core candidate pass), per-cluster fp64 sums and
    counts of the shard (hb_kmeans_update), all-reduce(sum) over the ranks, divide; an empty cluster keeps its centroid
    (:112-116).  `centroids` [nlist, d] fp64 is updated in place and identical on every rank afterwards.
    Returns (assign_ms, update_ms, allreduce_ms) measured with CUDA events."""
    import torch
    import torch.distributed as dist_

    L = hb.lib()
    n, d = rows.shape
    nlist = centroids.shape[0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    sums = torch.empty((nlist, d), dtype=torch.float64, device=rows.device)
    cnt = torch.empty(nlist, dtype=torch.int64, device=rows.device)
    ev[0].record()
    hb.check(L.hb_kmeans_assign(rows.data_ptr(), n, d, hb.dtype_code(rows), hb.COSINE, centroids.data_ptr(), nlist,
                                assign_out.data_ptr()))
    ev[1].record()
    hb.check(L.hb_kmeans_update(rows.data_ptr(), n, d, hb.dtype_code(rows), assign_out.data_ptr(), nlist, None,
                                sums.data_ptr(), cnt.data_ptr()))
    ev[2].record()
    if world > 1:
        dist_.all_reduce(sums, group=group)
        dist_.all_reduce(cnt, group=group)
    nz = cnt > 0
    centroids[nz] = sums[nz] / cnt[nz].to(torch.float64).unsqueeze(1)
    ev[3].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
