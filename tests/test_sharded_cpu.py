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


# ---------------------------------------------------------------------------------------------- IVF / INT8 sharding
def _merge_host_desc(ids_all, sc_all, k):
    """numpy statement of vs_merge_topk_dev(smallest=0): (score desc, id asc), -1 padding last."""
    g, nq, kk = ids_all.shape
    ids = ids_all.transpose(1, 0, 2).reshape(nq, g * kk)
    sc = sc_all.transpose(1, 0, 2).reshape(nq, g * kk).astype(np.float64)
    key = np.where(ids >= 0, -sc, np.inf)
    order = np.lexsort((ids.astype(np.uint32), key), axis=1)[:, :k]
    return np.take_along_axis(ids, order, 1), np.take_along_axis(sc_all.transpose(1, 0, 2).reshape(nq, g * kk), order, 1)


def _make_ivf(vsb, oracle, n, nlist, seed=5):
    base = vsb.synth.make("mix", seed, n)
    rng = np.random.default_rng(seed)
    cent = base[rng.choice(n, nlist, replace=False)].copy()
    lab, _ = oracle.kmeans_assign(base, cent)
    order = np.argsort(lab, kind="stable").astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(lab, minlength=nlist))]).astype(np.int32)
    return base, cent, order, offsets


def _ivf_worker(rank, world, port, n, nlist, nq, k, nprobe, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vsb200_loader
    from oracle import oracle

    vsb = vsb200_loader.load()
    from vsb200 import sharded

    base, cent, order, offsets = _make_ivf(vsb, oracle, n, nlist)
    qry = vsb.synth.make("mix", 6, nq)
    owner = sharded.assign_lists(offsets, world)
    vec, off, idm = sharded.local_ivf_arrays(base[order], offsets, order, owner, rank)
    coarse = oracle.ivf_coarse(qry, cent)   # replicated coarse stage: identical probe sets on every rank
    ids, sc, cnt, total = oracle.ivf_search(vec, off, idm, True, coarse, qry, k, nprobe, mode=1)
    ids_all, sc_all = sharded.allgather_topk(torch.from_numpy(ids), torch.from_numpy(sc))
    tot = torch.tensor([total], dtype=torch.int64)
    dist.all_reduce(tot)
    mi, ms = _merge_host_desc(ids_all.numpy(), sc_all.numpy(), k)
    np.savez(os.path.join(out_dir, f"ivf{rank}.npz"), ids=mi, sc=ms, total=tot.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_list_assignment_is_a_balanced_partition():
    sys.path.insert(0, ROOT)
    import vsb200_loader

    vsb200_loader.load()
    from vsb200 import sharded

    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 6000, size=1024)
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    for world in (1, 2, 4, 8):
        owner = sharded.assign_lists(offsets, world)
        assert owner.min() >= 0 and owner.max() < world
        loads = np.array([sizes[owner == r].sum() for r in range(world)])
        assert loads.sum() == sizes.sum() and loads.max() - loads.min() <= sizes.max()
        vec = np.arange(offsets[-1], dtype=np.float32)[:, None]
        seen = np.concatenate([sharded.local_ivf_arrays(vec, offsets, np.arange(offsets[-1]), owner, r)[2] for r in range(world)])
        assert np.array_equal(np.sort(seen), np.arange(offsets[-1]))   # every row owned exactly once


@pytest.mark.timeout(180)
def test_two_rank_gloo_ivf_lists_sharded_equals_single_index(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    import vsb200_loader

    vsb = vsb200_loader.load()
    n, nlist, nq, k, nprobe, world = 8000, 64, 29, 10, 8, 2
    mp.spawn(_ivf_worker, args=(world, _free_port(), n, nlist, nq, k, nprobe, str(tmp_path)), nprocs=world, join=True)
    base, cent, order, offsets = _make_ivf(vsb, oracle, n, nlist)
    qry = vsb.synth.make("mix", 6, nq)
    coarse = oracle.ivf_coarse(qry, cent)
    wi, ws, wc, wt = oracle.ivf_search(base[order], offsets, order, True, coarse, qry, k, nprobe, mode=1)
    for r in range(world):
        g = np.load(tmp_path / f"ivf{r}.npz")
        assert np.array_equal(g["ids"], wi) and np.array_equal(g["sc"], ws) and int(g["total"][0]) == wt


def _int8_worker(rank, world, port, n, nq, k, w_scale, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vsb200_loader
    from oracle import oracle

    vsb = vsb200_loader.load()
    from vsb200 import sharded

    base = vsb.synth.make("sift", 808, n)
    qry = vsb.synth.make("sift", 809, nq)
    r0, r1 = sharded.shard_range(n, rank, world)
    m = oracle.int8_multiplier(vsb.QNN_INPUT_SCALE, w_scale, vsb.QNN_OUTPUT_SCALE)
    ids, sc = oracle.int8_search(oracle.quantize_u8(base[r0:r1], w_scale), oracle.quantize_u8(qry, vsb.QNN_INPUT_SCALE), k, m, mode=1)
    ids = ids + r0
    ids_all, sc_all = sharded.allgather_topk(torch.from_numpy(ids), torch.from_numpy(sc.astype(np.float32)))
    mi, ms = _merge_host_desc(ids_all.numpy(), sc_all.numpy(), k)
    np.savez(os.path.join(out_dir, f"i8{rank}.npz"), ids=mi, sc=ms.astype(np.uint8))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_int8_rows_sharded_equals_single_index(tmp_path):
    """INT8 row sharding: one weight scale for all shards, u8 scores merged largest-first with ids ascending inside
    the (massive) score ties — the merged answer equals the unsharded CPU twin's."""
    sys.path.insert(0, ROOT)
    from oracle import oracle
    import vsb200_loader

    vsb = vsb200_loader.load()
    n, nq, k, world = 7001, 23, 10, 2
    base = vsb.synth.make("sift", 808, n)
    qry = vsb.synth.make("sift", 809, nq)
    w_scale = float(np.float32(base.max()) / np.float32(255.0))
    mp.spawn(_int8_worker, args=(world, _free_port(), n, nq, k, w_scale, str(tmp_path)), nprocs=world, join=True)
    m = oracle.int8_multiplier(vsb.QNN_INPUT_SCALE, w_scale, vsb.QNN_OUTPUT_SCALE)
    wi, ws = oracle.int8_search(oracle.quantize_u8(base, w_scale), oracle.quantize_u8(qry, vsb.QNN_INPUT_SCALE), k, m, mode=1)
    for r in range(world):
        g = np.load(tmp_path / f"i8{r}.npz")
        assert np.array_equal(g["ids"], wi) and np.array_equal(g["sc"], ws)


def _block_worker(rank, world, port, n, nq, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vsb200_loader
    from oracle import oracle

    vsb = vsb200_loader.load()
    from vsb200 import sharded

    base = vsb.synth.make("sift", 77, n)
    qry = vsb.synth.make("sift", 78, nq)
    r0, r1 = sharded.shard_range(n, rank, world)
    ids, d = oracle.exact_search(base[r0:r1], qry, k, mode=1)
    B = vsb.topk_block_bytes(nq, k)
    gathered = torch.zeros((world, B), dtype=torch.uint8)
    mine = gathered[rank]                                     # the shard's slot: ids | keys | trailer
    mine[:4 * nq * k].view(torch.int32).view(nq, k).copy_(torch.from_numpy(ids + r0))
    mine[4 * nq * k:8 * nq * k].view(torch.float32).view(nq, k).copy_(torch.from_numpy(d))
    mine[B - 16:B - 12].view(torch.int32)[0] = 3 + rank       # "uncertified count" of this shard
    dist.all_gather_into_tensor(gathered.view(-1), gathered[rank])   # ONE collective, in place
    g = gathered.numpy()
    ids_all = np.stack([g[s, :4 * nq * k].view(np.int32).reshape(nq, k) for s in range(world)])
    d_all = np.stack([g[s, 4 * nq * k:8 * nq * k].view(np.float32).reshape(nq, k) for s in range(world)])
    total = sum(int(g[s, B - 16:B - 12].view(np.int32)[0]) for s in range(world))
    mi, md = _merge_host(ids_all, d_all, k)
    np.savez(os.path.join(out_dir, f"blk{rank}.npz"), ids=mi, d=md, total=np.array([total]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_block_exchange_in_place(tmp_path):
    """The exchange of the row-sharded search as sharded.ShardedExact runs it: every rank fills ITS slot of the gathered
    buffer with an exchange block (vs_topk_block_bytes layout: ids | keys | trailer) and ONE in-place all-gather
    replicates the slots; ids, keys and the uncertified counts arrive together."""
    sys.path.insert(0, ROOT)
    from oracle import oracle
    import vsb200_loader

    vsb = vsb200_loader.load()
    n, nq, k, world = 5003, 31, 10, 2
    assert vsb.topk_block_bytes(nq, k) == (nq * k * 8 + 15) // 16 * 16 + 16
    mp.spawn(_block_worker, args=(world, _free_port(), n, nq, k, str(tmp_path)), nprocs=world, join=True)
    base = vsb.synth.make("sift", 77, n)
    qry = vsb.synth.make("sift", 78, nq)
    wi, wd = oracle.exact_search(base, qry, k, mode=1)
    for r in range(world):
        g = np.load(tmp_path / f"blk{r}.npz")
        assert np.array_equal(g["ids"], wi) and np.array_equal(g["d"], wd) and int(g["total"][0]) == 3 + 4
