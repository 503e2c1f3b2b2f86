"""Sharded multi-GPU search: one process per GPU (torch.distributed).  Exact / INT8: base rows partitioned
contiguously across ranks; IVF: inverted lists partitioned across ranks.  Queries are replicated, every rank computes
its local top-k, and there is ONE exchange step — an all-gather of the [nq x k] (id, key) candidates over NCCL /
NVLink — followed by the merge kernel (vs_merge_topk_dev) on every rank.

The reference has no multi-device code; this is the sharding BASELINE.json's north_star asks for.
The canonical (key, id) order of every local result makes the merged answer independent of the number of shards.
torch is plumbing here (process group, device tensors); all compute is libvsb200.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [r0, r1) of rank `rank`: contiguous, sizes differ by at most one, cover [0, n) exactly."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    return (n * rank) // world, (n * (rank + 1)) // world


def allgather_topk(ids_loc: torch.Tensor, keys_loc: torch.Tensor, group=None):
    """[nq, k] per rank -> ([G, nq, k] ids, [G, nq, k] keys), identical on every rank (shard g at index g: the layout
    vs_merge_topk_dev expects). Works for CUDA tensors over NCCL and for CPU tensors over gloo."""
    world = dist.get_world_size(group)
    nq, k = ids_loc.shape
    # the output is the concatenation of the ranks' inputs along dim 0 (the form every backend accepts)
    ids_all = torch.empty((world * nq, k), dtype=ids_loc.dtype, device=ids_loc.device)
    keys_all = torch.empty((world * nq, k), dtype=keys_loc.dtype, device=keys_loc.device)
    dist.all_gather_into_tensor(ids_all, ids_loc.contiguous(), group=group)
    dist.all_gather_into_tensor(keys_all, keys_loc.contiguous(), group=group)
    return ids_all.view(world, nq, k), keys_all.view(world, nq, k)


class ShardedExact:
    """Exact L2 kNN over a base sharded by rows across the ranks of a process group.

    index      this rank's vsb200.ExactIndex, created with id_base = first row of the shard
    search()   device query pointer -> (ids, dists) tensors on this rank's device holding the GLOBAL top-k
    """

    def __init__(self, vsb, index, nq_max: int, k: int, device: torch.device, group=None):
        self.vsb, self.index, self.k, self.group = vsb, index, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.ids_loc = torch.empty((nq_max, k), dtype=torch.int32, device=device)
        self.d_loc = torch.empty((nq_max, k), dtype=torch.float32, device=device)
        if self.world > 1:
            self.ids_all = torch.empty((self.world * nq_max, k), dtype=torch.int32, device=device)  # [G][nq][k]
            self.d_all = torch.empty((self.world * nq_max, k), dtype=torch.float32, device=device)
            self.ids_out = torch.empty((nq_max, k), dtype=torch.int32, device=device)
            self.d_out = torch.empty((nq_max, k), dtype=torch.float32, device=device)
        self.nq_max = nq_max

    def search(self, q_ptr: int, nq: int, precision: int, stream: int):
        """Enqueues on `stream` (torch's current stream must be that stream: NCCL orders against it)."""
        if nq != self.nq_max:
            raise ValueError("ShardedExact buffers are sized for nq_max queries per call")
        if self.world == 1:
            self.index.search_dev(q_ptr, nq, self.k, precision, self.ids_loc.data_ptr(), self.d_loc.data_ptr(), stream)
            return self.ids_loc, self.d_loc
        # enqueue the local search, the exchange and the merge back to back; only then wait for the certification count
        # of the local search (the launch latency of the collectives hides behind the fused kernel)
        self.index.search_dev_begin(q_ptr, nq, self.k, precision, self.ids_loc.data_ptr(), self.d_loc.data_ptr(), stream)
        self._exchange(nq, stream)
        redone = torch.tensor([self.index.search_dev_finish()], dtype=torch.int32, device=self.ids_loc.device)
        dist.all_reduce(redone, op=dist.ReduceOp.MAX, group=self.group)
        if int(redone.item()) > 0:  # some rank rewrote result rows after the exchange had been enqueued: exchange again
            self._exchange(nq, stream)
        return self.ids_out, self.d_out

    def _exchange(self, nq: int, stream: int):
        dist.all_gather_into_tensor(self.ids_all, self.ids_loc, group=self.group)
        dist.all_gather_into_tensor(self.d_all, self.d_loc, group=self.group)
        self.vsb.merge_topk_dev(self.ids_all.data_ptr(), self.d_all.data_ptr(), self.world, nq, self.k, True,
                                self.ids_out.data_ptr(), self.d_out.data_ptr(), stream)


# ------------------------------------------------------------------------------------------------------------------
# IVF: inverted LISTS are partitioned across the ranks: centroids and the coarse stage are replicated,
# every rank scans the probed lists it owns, local top-k (score desc, id asc) -> all-gather -> merge (largest first).
# A rank's index is an ordinary IVF index in which the lists it does not own are empty, so the single-GPU library
# path (vs_ivf_create / vs_ivf_search_dev) is used unchanged and the probe sets are identical on every rank.
# ------------------------------------------------------------------------------------------------------------------
def assign_lists(offsets, world: int):
    """Owner rank of every list: longest-first greedy onto the least loaded rank (balanced rows per rank, deterministic)."""
    import numpy as np

    sizes = np.diff(np.asarray(offsets, dtype=np.int64))
    owner = np.zeros(sizes.shape[0], dtype=np.int32)
    load = np.zeros(world, dtype=np.int64)
    for c in np.argsort(-sizes, kind="stable"):
        r = int(np.argmin(load))  # first minimum: deterministic
        owner[c] = r
        load[r] += sizes[c]
    return owner


def local_ivf_arrays(vectors_list_order, offsets, id_map, owner, rank: int):
    """(vectors, offsets, id_map) of rank `rank`: its own lists in list order, the others empty."""
    import numpy as np

    offsets = np.asarray(offsets, dtype=np.int64)
    sizes = np.diff(offsets)
    mine = owner == rank
    loc_sizes = np.where(mine, sizes, 0)
    loc_off = np.concatenate([[0], np.cumsum(loc_sizes)]).astype(np.int32)
    pos = np.concatenate([np.arange(offsets[c], offsets[c + 1]) for c in np.nonzero(mine)[0]] or [np.zeros(0, np.int64)])
    pos = pos.astype(np.int64)
    return (np.ascontiguousarray(np.asarray(vectors_list_order)[pos]), loc_off,
            np.ascontiguousarray(np.asarray(id_map)[pos]).astype(np.int32))


class ShardedIvf:
    """IVF search over lists partitioned across the ranks.  index = this rank's vsb200.IvfIndex built from
    local_ivf_arrays(); search() -> (ids, scores, counts) tensors holding the GLOBAL answer on every rank."""

    def __init__(self, vsb, index, nq_max: int, k: int, device: torch.device, group=None):
        self.vsb, self.index, self.k, self.group = vsb, index, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.ids_loc = torch.empty((nq_max, k), dtype=torch.int32, device=device)
        self.sc_loc = torch.empty((nq_max, k), dtype=torch.float32, device=device)
        self.cnt_loc = torch.empty((nq_max,), dtype=torch.int32, device=device)
        if self.world > 1:
            self.ids_all = torch.empty((self.world * nq_max, k), dtype=torch.int32, device=device)
            self.sc_all = torch.empty((self.world * nq_max, k), dtype=torch.float32, device=device)
            self.ids_out = torch.empty((nq_max, k), dtype=torch.int32, device=device)
            self.sc_out = torch.empty((nq_max, k), dtype=torch.float32, device=device)
        self.nq_max = nq_max

    def search(self, q_ptr: int, nq: int, nprobe: int, stream: int):
        if nq != self.nq_max:
            raise ValueError("ShardedIvf buffers are sized for nq_max queries per call")
        self.index.search_dev(q_ptr, nq, self.k, nprobe, self.ids_loc.data_ptr(), self.sc_loc.data_ptr(),
                              self.cnt_loc.data_ptr(), stream)
        if self.world == 1:
            return self.ids_loc, self.sc_loc, self.cnt_loc
        dist.all_gather_into_tensor(self.ids_all, self.ids_loc, group=self.group)
        dist.all_gather_into_tensor(self.sc_all, self.sc_loc, group=self.group)
        cnt = self.cnt_loc.clone()
        dist.all_reduce(cnt, group=self.group)  # candidates found over all ranks; the answer holds min(k, that)
        self.vsb.merge_topk_dev(self.ids_all.data_ptr(), self.sc_all.data_ptr(), self.world, nq, self.k, False,
                                self.ids_out.data_ptr(), self.sc_out.data_ptr(), stream)
        return self.ids_out, self.sc_out, torch.clamp(cnt, max=self.k)


class ShardedInt8:
    """INT8 brute force over base rows partitioned contiguously across the ranks (every rank must be created with the
    SAME weight scale, e.g. the all-reduced max of the base / 255).  search() -> (ids, u8 scores) on every rank."""

    def __init__(self, vsb, index, nq_max: int, k: int, device: torch.device, group=None):
        self.vsb, self.index, self.k, self.group = vsb, index, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.ids_loc = torch.empty((nq_max, k), dtype=torch.int32, device=device)
        self.sc_loc = torch.empty((nq_max, k), dtype=torch.uint8, device=device)
        if self.world > 1:
            self.ids_all = torch.empty((self.world * nq_max, k), dtype=torch.int32, device=device)
            self.sc_all = torch.empty((self.world * nq_max, k), dtype=torch.float32, device=device)
            self.ids_out = torch.empty((nq_max, k), dtype=torch.int32, device=device)
            self.sc_out = torch.empty((nq_max, k), dtype=torch.float32, device=device)
        self.nq_max = nq_max

    def search(self, q_ptr: int, nq: int, stream: int):
        if nq != self.nq_max:
            raise ValueError("ShardedInt8 buffers are sized for nq_max queries per call")
        self.index.search_dev(q_ptr, nq, self.k, self.ids_loc.data_ptr(), self.sc_loc.data_ptr(), stream)
        if self.world == 1:
            return self.ids_loc, self.sc_loc
        dist.all_gather_into_tensor(self.ids_all, self.ids_loc, group=self.group)
        dist.all_gather_into_tensor(self.sc_all, self.sc_loc.to(torch.float32), group=self.group)  # merge keys are fp32
        self.vsb.merge_topk_dev(self.ids_all.data_ptr(), self.sc_all.data_ptr(), self.world, nq, self.k, False,
                                self.ids_out.data_ptr(), self.sc_out.data_ptr(), stream)
        return self.ids_out, self.sc_out.to(torch.uint8)
