"""Row-sharded multi-GPU search: one process per GPU, every rank holds a contiguous block of the global rows.

This is the reference's own scale-out model — independent sub-indexes over row ranges searched in parallel and merged
by (apply concat) + (sort-by :distance) + (take k) (src/hnsw/ann/partition/partitioned_hnsw.clj:86-196) — applied to
ONE global index, so results are identical to the single-GPU results.

The data plane lives in the library (include/hnswb200.h, "multi-GPU"): `comm_init` hands every rank the communicator
id, `RowShardedIVFFlat` / `RowShardedFlat` build the local shard (data-parallel k-means with one all-reduce per Lloyd
round) and `search_raw` is ONE C-ABI call per rank: local search -> exchange of the local top-k over NVLink peer
windows (or ncclAllGather) -> merge kernel, no host synchronisation in between.  torch.distributed only carries the
128-byte id.

`ShardedIVFFlat` / `ShardedFlat` / `all_gather_merge` below are the earlier host-orchestrated form (torch collectives +
hb_topk_merge); the gloo tests use them to cover the host logic on CPU.
"""
from __future__ import annotations

import numpy as np

import ctypes as C

from . import _lib as hb
from . import ivf_flat
from .index import DeviceIndex, metric_code, new_handle


# ---- the library's data plane ---------------------------------------------------------------------------------------
def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    hb.check(hb.lib().hb_comm_unique_id(buf))
    return buf.raw


def comm_init(rank: int | None = None, world: int | None = None, id_bytes: bytes | None = None, group=None):
    """hb_comm_init for this process (after hb_init(LOCAL_RANK)).  Without `id_bytes` rank 0 draws the id and
    torch.distributed (any backend) broadcasts it."""
    if id_bytes is None:
        import torch.distributed as dist_

        rank = dist_.get_rank(group) if rank is None else rank
        world = dist_.get_world_size(group) if world is None else world
        box = [comm_unique_id() if rank == 0 else None]
        dist_.broadcast_object_list(box, src=0, group=group)
        id_bytes = box[0]
    hb.check(hb.lib().hb_comm_init(id_bytes, int(world), int(rank)))
    return comm_info()


def comm_info() -> dict:
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    hb.check(hb.lib().hb_comm_info(C.byref(a), C.byref(b), C.byref(c)))
    return {"nranks": a.value, "rank": b.value, "p2p": bool(c.value)}


def comm_shutdown():
    hb.check(hb.lib().hb_comm_shutdown())


def comm_allreduce(values, op="sum") -> np.ndarray:
    """Small host-side reduction through the library's communicator (timings, counters)."""
    x = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64).copy()
    hb.check(hb.lib().hb_comm_allreduce_f64(hb.ptr(x), x.size, 0 if op == "sum" else 1))
    return x


def comm_broadcast(arr: np.ndarray, root: int = 0) -> np.ndarray:
    hb.check(hb.lib().hb_comm_broadcast(hb.ptr(arr), arr.nbytes, root))
    return arr


class _ShardedSearch(DeviceIndex):
    def search_raw(self, queries, k: int, param: int = 0, out_ids=None, out_dist=None):
        """Global top-k on every rank (ids = global rows): one hb_sharded_search call.  Collective."""
        q = hb.as_matrix(queries, allow=(hb.F32, hb.F64))
        nq = q.shape[0]
        if out_ids is None:
            out_ids = np.empty((nq, k), dtype=np.int64)
        if out_dist is None:
            out_dist = np.empty((nq, k), dtype=np.float64)
        hb.check(hb.lib().hb_sharded_search(self._h, hb.ptr(q), hb.dtype_code(q), nq, k, param, hb.ptr(out_ids), hb.ptr(out_dist)))
        return out_ids, out_dist

    def local_search_raw(self, queries, k: int, param: int = 0, out_ids=None, out_dist=None):
        """This shard only (local row ids): plain hb_search."""
        return DeviceIndex.search_raw(self, queries, k, param, out_ids, out_dist)


class RowShardedIVFFlat(_ShardedSearch):
    """This rank's rows [first_row, first_row + n_local) of a global IVF-FLAT index (hb_sharded_ivf_build): the k-means
    runs over all ranks' rows, the centroids are replicated, the local slabs hold this rank's part of every list."""

    def __init__(self, local_rows, first_row: int, num_partitions: int, seed_rows, max_iterations=10, distance_fn="cosine"):
        rows = hb.as_matrix(local_rows)
        sr = np.ascontiguousarray(seed_rows, dtype=np.int64)
        if sr.shape != (num_partitions,):
            raise hb.HbInvalid(hb.ERR_INVALID, "seed_rows must hold num_partitions global row ids")
        h = new_handle()
        hb.check(hb.lib().hb_sharded_ivf_build(hb.ptr(rows), rows.shape[0], rows.shape[1], hb.dtype_code(rows),
                                               metric_code(distance_fn), int(num_partitions), int(max_iterations), hb.ptr(sr),
                                               int(first_row), C.byref(h)))
        super().__init__(h.value, None)
        self.first_row, self.num_partitions = int(first_row), int(num_partitions)

    def export(self):
        i = self.info()
        cents = np.empty((i["nlist"], i["dim"]), dtype=np.float64)
        asg = np.empty(i["n"], dtype=np.int32)
        hb.check(hb.lib().hb_ivf_export(self._h, hb.ptr(cents), hb.ptr(asg)))
        return cents, asg


class RowShardedFlat(_ShardedSearch):
    """This rank's rows [first_row, first_row + n_local) of an exact flat index (BASELINE configs[2])."""

    def __init__(self, local_rows, first_row: int, distance_fn="cosine"):
        rows = hb.as_matrix(local_rows)
        h = new_handle()
        hb.check(hb.lib().hb_flat_create(hb.ptr(rows), rows.shape[0], rows.shape[1], hb.dtype_code(rows), metric_code(distance_fn),
                                         C.byref(h)))
        super().__init__(h.value, None)
        hb.check(hb.lib().hb_index_set_id_base(self._h, int(first_row)))
        self.first_row = int(first_row)


def import_row_shard(local_rows, first_row: int, centroids, local_assignments, distance_fn="cosine") -> _ShardedSearch:
    """A row shard of a global IVF-FLAT index from given centroids + this shard's assignments (parity: oracle-built
    partitions split by rows)."""
    ix = ivf_flat.import_index(local_rows, centroids, local_assignments, distance_fn)
    sh = _ShardedSearch(ix.__dict__.pop("_h").value, None)
    ix._h = None
    hb.check(hb.lib().hb_index_set_id_base(sh._h, int(first_row)))
    return sh


def sharded_kmeans(local_rows, first_row: int, num_partitions: int, seed_rows, max_iterations=10, distance_fn="cosine"):
    """hb_sharded_kmeans -> (centroids fp64 [nlist, d] — identical on every rank, assignments int32 of this shard)."""
    rows = hb.as_matrix(local_rows)
    sr = np.ascontiguousarray(seed_rows, dtype=np.int64)
    cents = np.empty((num_partitions, rows.shape[1]), dtype=np.float64)
    asg = np.empty(rows.shape[0], dtype=np.int32)
    hb.check(hb.lib().hb_sharded_kmeans(hb.ptr(rows), rows.shape[0], rows.shape[1], hb.dtype_code(rows), metric_code(distance_fn),
                                        int(num_partitions), int(max_iterations), hb.ptr(sr), int(first_row), hb.ptr(cents),
                                        hb.ptr(asg)))
    return cents, asg


# ---- the earlier host-orchestrated form (torch collectives + hb_topk_merge) ------------------------------------------
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
