"""Host-side logic of the row-sharded multi-GPU path, exercised with world_size = 2 over gloo on CPU: shard ranges,
the all-gather layout handed to the merge, and the invariance of the merged answer w.r.t. the number of shards.
The local searches here are done by the CPU oracle (the checker); on GPUs they are libvsb200 and the merge is
vs_merge_topk_dev (tests/test_exact_gpu.py::test_properties_full_size checks that merge against one index)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_host(ids_all, keys_all, k):
    """numpy statement of vs_merge_topk_dev (smallest first, canonical (key, id) order)."""
    g, nq, kk = ids_all.shape
    ids = ids_all.transpose(1, 0, 2).reshape(nq, g * kk)
    keys = keys_all.transpose(1, 0, 2).reshape(nq, g * kk)
    order = np.lexsort((ids, keys), axis=1)[:, :k]
    return np.take_along_axis(ids, order, 1), np.take_along_axis(keys, order, 1)


def _worker(rank, world, port, n, nq, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vsb200_loader
    from oracle import oracle

    vsb = vsb200_loader.load()
    from vsb200 import sharded

    base = vsb.synth.make("sift", 77, n)       # every rank can regenerate any row: counter-based generator
    qry = vsb.synth.make("sift", 78, nq)
    r0, r1 = sharded.shard_range(n, rank, world)
    ids, d = oracle.exact_search(base[r0:r1], qry, k, mode=1)
    ids = ids + r0                              # id_base of the shard
    ids_all, d_all = sharded.allgather_topk(torch.from_numpy(ids), torch.from_numpy(d))
    assert ids_all.shape == (world, nq, k)
    assert np.array_equal(ids_all[rank].numpy(), ids)     # shard g sits at index g
    mi, md = _merge_host(ids_all.numpy(), d_all.numpy(), k)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=mi, d=md)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    sys.path.insert(0, ROOT)
    import vsb200_loader

    vsb200_loader.load()
    from vsb200 import sharded

    for n in (1, 7, 1000, 1_000_000, 100_000_000):
        for world in (1, 2, 4, 8):
            rs = [sharded.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.shard_range(10, 2, 2)


@pytest.mark.timeout(180)
def test_two_rank_gloo_merge_equals_single_index(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    import vsb200_loader

    vsb = vsb200_loader.load()
    n, nq, k, world = 6001, 37, 10, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, nq, k, str(tmp_path)), nprocs=world, join=True)
    base = vsb.synth.make("sift", 77, n)
    qry = vsb.synth.make("sift", 78, nq)
    want_ids, want_d = oracle.exact_search(base, qry, k, mode=1)
    for r in range(world):   # identical, complete answer on every rank, equal to the unsharded oracle
        g = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(g["ids"], want_ids) and np.array_equal(g["d"], want_d)
