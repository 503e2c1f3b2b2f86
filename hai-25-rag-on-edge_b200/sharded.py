"""Sharded multi-GPU search: one process per GPU (torch.distributed).  Exact / INT8: base rows partitioned
contiguously across ranks; IVF: inverted lists partitioned across ranks.  Queries are replicated, every rank computes
its local top-k, and there is ONE exchange step — ONE in-place all-gather over NCCL / NVLink of the per-shard exchange
blocks (ids | keys | trailer, vs_topk_block_bytes) — followed by the merge kernel (vs_merge_blocks_dev) on every rank.

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
    """Exact L2 kNN over a base sharded by rows: this process holds `index` (one vsb200.ExactIndex, or a list of them —
    several shards on one device) created with id_base = first row of the shard; the other ranks of the process group
    hold the other shards.  One exchange step: every shard writes its exchange block (ids | dists | uncertified count,
    vsb200.topk_block_bytes) IN PLACE into its slot of the gathered buffer, ONE in-place all-gather replicates the
    slots, the merge kernel reads them all.  The begin / merge / finish sequence is the C ABI's (vs_exact_group_*);
    `exchange` replaces the collective (tests: all shards local -> nothing to do).

    search()   device query pointer -> (ids, dists) tensors on this rank's device holding the GLOBAL top-k
    enqueue() + finish()   the same in two halves: enqueue() never blocks the host; finish() waits for the 4-byte
               total of the uncertified counts and, when some shard had to redo queries, exchanges and merges again.
    """

    def __init__(self, vsb, index, nq_max: int, k: int, device: torch.device, group=None, exchange=None, push=None):
        """exchange: callable replacing the collective (tests).  push: True = exchange by peer stores into symmetric memory
        (vs_push_block_dev + a cross-GPU barrier), False = in-place NCCL all-gather, None = push when torch's symmetric
        memory can be set up for the group (one node, NVLink peers), else NCCL; VSB_EXCHANGE=nccl|push overrides."""
        self.vsb, self.k, self.group = vsb, k, group
        self.shards = list(index) if isinstance(index, (list, tuple)) else [index]
        self.index = self.shards[0]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_local = len(self.shards)
        self.n_slots = self.world * self.n_local
        self.first_slot = self.rank * self.n_local
        self.nq_max = nq_max
        self.exchanges = 0
        self.ids_loc = torch.empty((nq_max, k), dtype=torch.int32, device=device)
        self.d_loc = torch.empty((nq_max, k), dtype=torch.float32, device=device)
        self._exchange = exchange if exchange is not None else self._allgather
        self.exchange_kind = "none" if self.world == 1 else "nccl"
        self._sym = None
        if self.n_slots > 1:
            self.block = vsb.topk_block_bytes(nq_max, k)
            size = self.n_slots * self.block
            import os

            env = os.environ.get("VSB_EXCHANGE", "")
            want_push = (push if push is not None else True) if env == "" else env == "push"
            if exchange is None and self.world > 1 and want_push and device.type == "cuda":
                err = None
                try:
                    self._setup_push(size, device)
                except Exception as e:  # no symmetric memory on this platform
                    err = e
                # every rank must take the same exchange: push only when ALL of them could set it up
                ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
                if int(ok.item()) == 1:
                    self._exchange = self._push_exchange
                    self.exchange_kind = "push"
                else:
                    if err is not None and (push or env == "push"):
                        raise err
                    self._sym = None
                    self.exchange_kind = "nccl (symmetric memory unavailable%s)" % (": " + type(err).__name__ if err else " on a peer")
            if self._sym is None:
                self.gathered = torch.zeros(size, dtype=torch.uint8, device=device)  # [slot][block]
            self.grp = vsb.ExactGroup(self.shards, self.n_slots, self.first_slot)

    def close(self):
        if self.n_slots > 1:
            self.grp.close()

    # ---- push exchange: the gathered buffers of all ranks are ONE symmetric allocation (two halves used alternately), every rank
    # stores its slots straight into the peers' buffers over NVLink (one kernel, vs_push_block_dev) and a device-side barrier
    # on the symmetric memory's signal pad replaces the NCCL all-gather.  Half h is written again two exchanges later; the
    # barrier in between guarantees that every peer has finished merging it.
    def _setup_push(self, size: int, device):
        import torch.distributed._symmetric_memory as symm_mem

        grp = self.group if self.group is not None else dist.group.WORLD
        self._half = (size + 255) // 256 * 256
        # [half 0 | half 1 | flag words: one per sender]
        self._sym = symm_mem.empty(2 * self._half + 256, dtype=torch.uint8, device=device)
        self._sym.zero_()
        self._hdl = symm_mem.rendezvous(self._sym, grp)
        self._peer_ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        if len(self._peer_ptrs) != self.world:
            raise RuntimeError("symmetric memory: unexpected number of peers")
        self._peers = [r for r in range(self.world) if r != self.rank]
        self._flag_dst = [self._peer_ptrs[r] + 2 * self._half + 4 * self.rank for r in self._peers]  # my word on every peer
        self._flags = self._peer_ptrs[self.rank] + 2 * self._half
        self._counter = torch.zeros(1, dtype=torch.int32, device=device)
        self._epoch = 0
        self._parity = 0
        self._stream = 0
        self.gathered = self._sym[: self._half]
        torch.cuda.synchronize(device)
        self._hdl.barrier(channel=0)  # every rank's flags are zero before anybody pushes
        torch.cuda.synchronize(device)

    def _push(self, nbytes: int):
        self._epoch += 1
        off = self._parity * self._half + self.first_slot * self.block
        dst = [self._peer_ptrs[r] + off for r in self._peers]
        self.vsb.push_block_dev(self._peer_ptrs[self.rank] + off, dst, nbytes, self._flag_dst, self._epoch,
                                self._counter.data_ptr(), self._stream)
        self.vsb.wait_flags_dev(self._flags, self.world, self.rank, self._epoch, self._stream)

    def _push_exchange(self, redo: bool = False):
        if redo:  # the same half is pushed twice in one step: the peers must be done merging the first version
            self._push(0)
        self._push(self.n_local * self.block)

    def _allgather(self):
        if self.world > 1:
            mine = self.gathered[self.first_slot * self.block:(self.first_slot + self.n_local) * self.block]
            dist.all_gather_into_tensor(self.gathered, mine, group=self.group)  # in place: slot s comes from its owner

    def enqueue(self, q_ptr: int, nq: int, precision: int, stream: int):
        """Enqueues on `stream` (torch's current stream must be that stream: NCCL orders against it)."""
        if nq != self.nq_max:
            raise ValueError("ShardedExact buffers are sized for nq_max queries per call")
        if self.n_slots == 1:
            self.index.search_dev(q_ptr, nq, self.k, precision, self.ids_loc.data_ptr(), self.d_loc.data_ptr(), stream)
            return self.ids_loc, self.d_loc
        if self._sym is not None:  # next half of the symmetric buffer
            self._parity ^= 1
            self._stream = stream
            self.gathered = self._sym[self._parity * self._half:(self._parity + 1) * self._half]
        self.grp.begin(q_ptr, nq, self.k, precision, self.gathered.data_ptr(), stream)
        self._exchange()
        self.exchanges = 1
        self.grp.merge(self.ids_loc.data_ptr(), self.d_loc.data_ptr())
        return self.ids_loc, self.d_loc

    def finish(self) -> int:
        """-> number of extra exchanges (0 unless a shard could not certify some queries and redid them in fp32)."""
        extra = 0
        if self.n_slots > 1:
            while self.grp.finish():
                if self._sym is not None:
                    self._push_exchange(redo=True)
                else:
                    self._exchange()
                self.grp.merge(self.ids_loc.data_ptr(), self.d_loc.data_ptr())
                extra += 1
                self.exchanges += 1
        return extra

    def search(self, q_ptr: int, nq: int, precision: int, stream: int):
        out = self.enqueue(q_ptr, nq, precision, stream)
        self.finish()
        return out


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
    local_ivf_arrays(); search() -> (ids, scores, counts) tensors holding the GLOBAL answer on every rank.
    One collective: a rank's slot = exchange block (ids | scores | trailer) followed by its candidate counts [nq]."""

    def __init__(self, vsb, index, nq_max: int, k: int, device: torch.device, group=None):
        self.vsb, self.index, self.k, self.group = vsb, index, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.block = vsb.topk_block_bytes(nq_max, k)
        self.slot = (self.block + 4 * nq_max + 15) // 16 * 16
        self.gathered = torch.zeros((self.world, self.slot), dtype=torch.uint8, device=device)
        mine = self.gathered[self.rank]
        self.ids_loc = mine[:4 * nq_max * k].view(torch.int32).view(nq_max, k)
        self.sc_loc = mine[4 * nq_max * k:8 * nq_max * k].view(torch.float32).view(nq_max, k)
        self.cnt_loc = mine[self.block:self.block + 4 * nq_max].view(torch.int32)
        if self.world > 1:
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
        dist.all_gather_into_tensor(self.gathered.view(-1), self.gathered[self.rank], group=self.group)  # in place
        self.vsb.merge_blocks_dev(self.gathered.data_ptr(), self.world, self.slot, nq, self.k, False,
                                  self.ids_out.data_ptr(), self.sc_out.data_ptr(), 0, stream)
        # candidates found over all ranks; the answer holds min(k, that)
        cnt = self.gathered[:, self.block:self.block + 4 * nq].contiguous().view(torch.int32).sum(0, dtype=torch.int32)
        return self.ids_out, self.sc_out, torch.clamp(cnt, max=self.k)


class ShardedInt8:
    """INT8 brute force over base rows partitioned contiguously across the ranks (every rank must be created with the
    SAME weight scale, e.g. the all-reduced max of the base / 255).  search() -> (ids, u8 scores) on every rank.
    One collective: the exchange block of a rank holds its ids and its scores widened to fp32 (the merge's key type)."""

    def __init__(self, vsb, index, nq_max: int, k: int, device: torch.device, group=None):
        self.vsb, self.index, self.k, self.group = vsb, index, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.block = vsb.topk_block_bytes(nq_max, k)
        self.gathered = torch.zeros((self.world, self.block), dtype=torch.uint8, device=device)
        mine = self.gathered[self.rank]
        self.ids_loc = mine[:4 * nq_max * k].view(torch.int32).view(nq_max, k)
        self.key_loc = mine[4 * nq_max * k:8 * nq_max * k].view(torch.float32).view(nq_max, k)
        self.sc_loc = torch.empty((nq_max, k), dtype=torch.uint8, device=device)
        if self.world > 1:
            self.ids_out = torch.empty((nq_max, k), dtype=torch.int32, device=device)
            self.sc_out = torch.empty((nq_max, k), dtype=torch.float32, device=device)
        self.nq_max = nq_max

    def search(self, q_ptr: int, nq: int, stream: int):
        if nq != self.nq_max:
            raise ValueError("ShardedInt8 buffers are sized for nq_max queries per call")
        self.index.search_dev(q_ptr, nq, self.k, self.ids_loc.data_ptr(), self.sc_loc.data_ptr(), stream)
        if self.world == 1:
            return self.ids_loc, self.sc_loc
        self.key_loc.copy_(self.sc_loc)  # u8 -> fp32 keys inside the block
        dist.all_gather_into_tensor(self.gathered.view(-1), self.gathered[self.rank], group=self.group)  # in place
        self.vsb.merge_blocks_dev(self.gathered.data_ptr(), self.world, self.block, nq, self.k, False,
                                  self.ids_out.data_ptr(), self.sc_out.data_ptr(), 0, stream)
        return self.ids_out, self.sc_out.to(torch.uint8)
