"""Row-sharded multi-GPU search: one process per GPU (torch.distributed), base rows partitioned contiguously across
ranks, queries replicated, per-rank local top-k on the local shard, ONE exchange step — an all-gather of the
[nq x k] (id, key) candidates over NCCL / NVLink — and the merge kernel (vs_merge_topk_dev) on every rank.

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
        self.index.search_dev(q_ptr, nq, self.k, precision, self.ids_loc.data_ptr(), self.d_loc.data_ptr(), stream)
        if self.world == 1:
            return self.ids_loc, self.d_loc
        dist.all_gather_into_tensor(self.ids_all, self.ids_loc, group=self.group)
        dist.all_gather_into_tensor(self.d_all, self.d_loc, group=self.group)
        self.vsb.merge_topk_dev(self.ids_all.data_ptr(), self.d_all.data_ptr(), self.world, nq, self.k, True,
                                self.ids_out.data_ptr(), self.d_out.data_ptr(), stream)
        return self.ids_out, self.d_out
