"""Row-sharded exact search on real GPUs: one process per GPU over NCCL (the path bench.py --gpus N runs), checked
against the CPU oracle on the unsharded base.  Needs >= 2 visible devices; on a 1-GPU box the N > 1 cases are
reported as skipped (the single-GPU merge itself is covered by tests/test_exact_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from util import assert_topk_matches

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, law, n, nq, k, prec, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist

    import vsb200_loader

    vsb = vsb200_loader.load()
    from vsb200 import sharded

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    r0, r1 = sharded.shard_range(n, rank, world)
    base_d = torch.empty((r1 - r0, 128), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base_d.data_ptr(), r0, r1 - r0, 128, law, 4242)
    torch.cuda.synchronize()
    index = vsb.ExactIndex(base_d.data_ptr(), device=rank, id_base=r0, n=r1 - r0)
    q_dev = torch.from_numpy(vsb.synth.make(law, 4343, nq)).to(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    s = sharded.ShardedExact(vsb, index, nq, k, dev)
    for _ in range(2):  # twice: the second call reuses every workspace
        ids, d = s.search(q_dev.data_ptr(), nq, prec, stream.cuda_stream)
    stream.synchronize()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=ids.cpu().numpy(), d=d.cpu().numpy())
    dist.barrier()
    index.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("law,k,prec_name", [("sift", 10, "auto"), ("cont", 10, "f16cert"), ("cont", 10, "3xtf32"),
                                             ("cont", 100, "auto")])
def test_nccl_sharded_equals_oracle(gpu_vsb, oracle, tmp_path, law, k, prec_name):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    vsb = gpu_vsb
    prec = {"auto": vsb.PREC_AUTO, "3xtf32": vsb.PREC_3XTF32, "f16cert": vsb.PREC_F16_CERT}[prec_name]
    n, nq = 200_003, 300
    mp.spawn(_worker, args=(world, _free_port(), law, n, nq, k, prec, str(tmp_path)), nprocs=world, join=True)
    base = vsb.synth.make(law, 4242, n)
    qry = vsb.synth.make(law, 4343, nq)
    want_ids, want_d = oracle.exact_search(base, qry, k, mode=1)
    got0 = np.load(tmp_path / "rank0.npz")
    for r in range(world):
        g = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(g["ids"], got0["ids"]) and np.array_equal(g["d"], got0["d"]), "ranks disagree"
    rec = oracle.exact_distances_at(base, qry, got0["ids"])
    assert_topk_matches(got0["ids"], got0["d"], want_ids, want_d, rec, exact=(law == "sift"),
                        what=f"nccl x{world} {law} k={k} {prec_name}")


# ---------------------------------------------------------------------------------------------- IVF lists / INT8 rows
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
    import torch
    import torch.distributed as dist

    import vsb200_loader
    from oracle import oracle   # test infrastructure: builds the shared index arrays (assignment to fixed centroids)

    vsb = vsb200_loader.load()
    from vsb200 import sharded

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    base, cent, order, offsets = _make_ivf(vsb, oracle, n, nlist)
    owner = sharded.assign_lists(offsets, world)
    vec, off, idm = sharded.local_ivf_arrays(base[order], offsets, order, owner, rank)
    index = vsb.IvfIndex(vectors=vec, offsets=off, id_map=idm, centroids=cent, device=rank)
    q_dev = torch.from_numpy(vsb.synth.make("mix", 6, nq)).to(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    s = sharded.ShardedIvf(vsb, index, nq, k, dev)
    for _ in range(2):
        ids, sc, cnt = s.search(q_dev.data_ptr(), nq, nprobe, stream.cuda_stream)
    stream.synchronize()
    np.savez(os.path.join(out_dir, f"ivf{rank}.npz"), ids=ids.cpu().numpy(), sc=sc.cpu().numpy(), cnt=cnt.cpu().numpy())
    dist.barrier()
    index.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_nccl_ivf_lists_sharded_equals_oracle(gpu_vsb, oracle, tmp_path):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    vsb = gpu_vsb
    n, nlist, nq, k, nprobe = 60_000, 128, 257, 10, 16
    mp.spawn(_ivf_worker, args=(world, _free_port(), n, nlist, nq, k, nprobe, str(tmp_path)), nprocs=world, join=True)
    base, cent, order, offsets = _make_ivf(vsb, oracle, n, nlist)
    qry = vsb.synth.make("mix", 6, nq)
    coarse = oracle.ivf_coarse(qry, cent)
    wi, ws, wc, _ = oracle.ivf_search(base[order], offsets, order, True, coarse, qry, k, nprobe, mode=1)
    for r in range(world):   # identical, complete answer on every rank: same ids, bit-identical scores, same counts
        g = np.load(tmp_path / f"ivf{r}.npz")
        assert np.array_equal(g["ids"], wi) and np.array_equal(g["sc"], ws) and np.array_equal(g["cnt"], wc)


def _int8_worker(rank, world, port, n, nq, k, w_scale, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist

    import vsb200_loader

    vsb = vsb200_loader.load()
    from vsb200 import sharded

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    r0, r1 = sharded.shard_range(n, rank, world)
    base = vsb.synth.make("sift", 808, n)[r0:r1]
    index = vsb.Int8Index(base, w_scale=w_scale, device=rank, id_base=r0)
    q_dev = torch.from_numpy(vsb.synth.make("sift", 809, nq)).to(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    s = sharded.ShardedInt8(vsb, index, nq, k, dev)
    for _ in range(2):
        ids, sc = s.search(q_dev.data_ptr(), nq, stream.cuda_stream)
    stream.synchronize()
    np.savez(os.path.join(out_dir, f"i8{rank}.npz"), ids=ids.cpu().numpy(), sc=sc.cpu().numpy())
    dist.barrier()
    index.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("nq", [40, 300])
def test_nccl_int8_rows_sharded_equals_cpu_twin(gpu_vsb, oracle, tmp_path, nq):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    vsb = gpu_vsb
    n, k = 50_001, 10
    base = vsb.synth.make("sift", 808, n)
    qry = vsb.synth.make("sift", 809, nq)
    w_scale = float(np.float32(base.max()) / np.float32(255.0))   # one weight scale for every shard
    mp.spawn(_int8_worker, args=(world, _free_port(), n, nq, k, w_scale, str(tmp_path)), nprocs=world, join=True)
    m = oracle.int8_multiplier(vsb.QNN_INPUT_SCALE, w_scale, vsb.QNN_OUTPUT_SCALE)
    wi, ws = oracle.int8_search(oracle.quantize_u8(base, w_scale), oracle.quantize_u8(qry, vsb.QNN_INPUT_SCALE), k, m, mode=1)
    for r in range(world):
        g = np.load(tmp_path / f"i8{r}.npz")
        assert np.array_equal(g["ids"], wi) and np.array_equal(g["sc"], ws)


# ---------------------------------------------------------------------------------------------------------------
# The sharded LOGIC on ONE GPU (runs on the 1-GPU driver box): G shards of one base live on cuda:0, every shard writes
# its exchange block into its slot of the gathered buffer, the exchange itself is a no-op (all slots are local) — the
# begin -> exchange -> merge -> finish -> (redo -> exchange -> merge -> finish) sequence is exactly the one the NCCL
# ranks run (sharded.ShardedExact over vs_exact_group_*; vs_exact_mgpu_* in C++).
# ---------------------------------------------------------------------------------------------------------------
def _base_with_uncertifiable_cluster(vsb, law, n, nq, seed=91):
    rng = np.random.default_rng(seed)
    base = vsb.synth.make(law, seed, n)
    qry = vsb.synth.make(law, seed + 1, nq)
    n_adv = 0
    if law == "cont":
        centre = base[7].copy()
        lo = (n * 5) // 8 + 100          # inside ONE shard for G = 2 and G = 8
        base[lo:lo + 200] = centre[None, :] + rng.uniform(-0.02, 0.02, (200, 128)).astype(np.float32)
        n_adv = 12
        qry[:n_adv] = centre[None, :] + rng.uniform(-0.05, 0.05, (n_adv, 128)).astype(np.float32)
    return base, qry, n_adv


@pytest.mark.parametrize("G", [2, 8])
@pytest.mark.parametrize("law,k,prec_name", [("cont", 10, "f16cert"), ("cont", 10, "3xtf32"), ("sift", 10, "auto"),
                                             ("cont", 100, "auto"), ("sift", 5, "ffma")])
def test_group_sequence_all_shards_on_one_gpu(G, law, k, prec_name, gpu_vsb, oracle):
    import torch

    vsb = gpu_vsb
    from vsb200 import sharded

    prec = {"auto": vsb.PREC_AUTO, "3xtf32": vsb.PREC_3XTF32, "f16cert": vsb.PREC_F16_CERT, "ffma": vsb.PREC_FFMA}[prec_name]
    n, nq = 120_003, 16 if prec_name == "ffma" else 500   # shards of >= 15 K rows: the certified fp16 pass runs in every shard
    base, qry, n_adv = _base_with_uncertifiable_cluster(vsb, law, n, nq)
    dev = torch.device("cuda:0")
    base_d = torch.from_numpy(base).to(dev)
    q_dev = torch.from_numpy(qry).to(dev)
    shards = []
    for g in range(G):
        r0, r1 = sharded.shard_range(n, g, G)
        shards.append(vsb.ExactIndex(base_d.data_ptr() + r0 * 128 * 4, device=0, id_base=r0, n=r1 - r0))
    calls = []
    st = torch.cuda.Stream()
    s = sharded.ShardedExact(vsb, shards, nq, k, dev, exchange=lambda: calls.append(1))
    try:
        assert s.n_slots == G
        with torch.cuda.stream(st):
            for rep in range(2):  # twice: the second call reuses every workspace
                calls.clear()
                ids, d = s.search(q_dev.data_ptr(), nq, prec, st.cuda_stream)
                st.synchronize()
                certified = prec_name == "f16cert" or (prec_name == "auto" and k <= 16 and nq >= 3)
                if certified and n_adv:
                    assert len(calls) == 2, "uncertified queries in one shard must trigger exactly one re-exchange"
                elif certified:   # integer data: a tie at the k-th boundary cannot be certified either
                    assert len(calls) in (1, 2)
                else:
                    assert len(calls) == 1
        want_ids, want_d = oracle.exact_search(base, qry, k, mode=1)
        got_i, got_d = ids.cpu().numpy(), d.cpu().numpy()
        rec = oracle.exact_distances_at(base, qry, got_i)
        if n_adv:  # near-duplicates: distances ~1e-1 after cancellation of ~1e5-sized terms, compared absolutely
            assert np.allclose(got_d[:n_adv], want_d[:n_adv], rtol=0, atol=0.5)
        assert_topk_matches(got_i[n_adv:], got_d[n_adv:], want_ids[n_adv:], want_d[n_adv:], rec[n_adv:],
                            exact=(law == "sift"), what=f"group x{G} {law} k={k} {prec_name}")
    finally:
        s.close()
        for x in shards:
            x.close()


@pytest.mark.parametrize("spg", [1, 3, 8])
def test_mgpu_c_abi_on_one_gpu(spg, gpu_vsb, oracle):
    """vs_exact_mgpu_* (the C-ABI multi-GPU entry point: SURVEY §8b `n_gpus`) with n_gpus = 1 and shards_per_gpu
    shards: the same code path as on 8 GPUs minus the NCCL all-gather."""
    vsb = gpu_vsb
    n, nq, k = 120_001, 700, 10
    base, qry, n_adv = _base_with_uncertifiable_cluster(vsb, "cont", n, nq, seed=33)
    want_ids, want_d = oracle.exact_search(base, qry, k, mode=1)
    m = vsb.ExactMultiGpu(base, n_gpus=1, shards_per_gpu=spg)
    try:
        assert (m.n_gpus, m.n_shards) == (1, spg)
        for prec in (vsb.PREC_AUTO, vsb.PREC_3XTF32):
            for rep in range(2):
                ids, d = m.search(qry, k, prec)
                exch, redone = m.last_stats()
                assert (exch, redone) == ((2, True) if prec == vsb.PREC_AUTO else (1, False))
                rec = oracle.exact_distances_at(base, qry, ids)
                assert np.allclose(d[:n_adv], want_d[:n_adv], rtol=0, atol=0.5)
                assert_topk_matches(ids[n_adv:], d[n_adv:], want_ids[n_adv:], want_d[n_adv:], rec[n_adv:], exact=False,
                                    what=f"mgpu spg={spg} {vsb.PREC_NAMES[prec]}")
        ids, d = m.search(qry[:5], 100, vsb.PREC_AUTO)          # k > 32: multi-pass per shard, any-k merge
        wi, wd = oracle.exact_search(base, qry[:5], 100, mode=1)
        assert np.allclose(d, wd, rtol=1e-5, atol=0.5)
    finally:
        m.close()


def test_mgpu_c_abi_all_gpus(gpu_vsb, oracle):
    """vs_exact_mgpu_* over every visible GPU: ncclCommInitAll + grouped ncclAllGather inside libvsb200."""
    import torch

    vsb = gpu_vsb
    G = min(torch.cuda.device_count(), 8)
    if G < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    n, nq, k = 200_003, 1000, 10
    base, qry, n_adv = _base_with_uncertifiable_cluster(vsb, "cont", n, nq, seed=35)
    want_ids, want_d = oracle.exact_search(base, qry, k, mode=1)
    m = vsb.ExactMultiGpu(base, n_gpus=G)
    try:
        for prec in (vsb.PREC_AUTO, vsb.PREC_3XTF32):
            ids, d = m.search(qry, k, prec)
            rec = oracle.exact_distances_at(base, qry, ids)
            assert np.allclose(d[:n_adv], want_d[:n_adv], rtol=0, atol=0.5)
            assert_topk_matches(ids[n_adv:], d[n_adv:], want_ids[n_adv:], want_d[n_adv:], rec[n_adv:], exact=False,
                                what=f"mgpu x{G} {vsb.PREC_NAMES[prec]}")
    finally:
        m.close()


@pytest.mark.parametrize("G", [2, 8])
def test_ivf_list_shards_and_int8_row_shards_on_one_gpu(G, gpu_vsb, oracle):
    """The IVF (lists partitioned) and INT8 (rows partitioned) exchange on ONE GPU: every emulated rank searches its
    local index straight into its slot of the gathered buffer (exchange block layout), vs_merge_blocks_dev merges."""
    import torch

    vsb = gpu_vsb
    from vsb200 import sharded

    dev = torch.device("cuda:0")
    st = torch.cuda.Stream()          # a real stream: handle 0 would mean "the handle's own stream" to the C ABI
    torch.cuda.set_stream(st)
    sptr = st.cuda_stream
    # ---- IVF
    n, nlist, nq, k, nprobe = 60_000, 128, 257, 10, 16
    base, cent, order, offsets = _make_ivf(vsb, oracle, n, nlist)
    qry = vsb.synth.make("mix", 6, nq)
    q_dev = torch.from_numpy(qry).to(dev)
    owner = sharded.assign_lists(offsets, G)
    B = vsb.topk_block_bytes(nq, k)
    slot = (B + 4 * nq + 15) // 16 * 16
    gathered = torch.zeros((G, slot), dtype=torch.uint8, device=dev)
    for r in range(G):
        vec, off, idm = sharded.local_ivf_arrays(base[order], offsets, order, owner, r)
        idx = vsb.IvfIndex(vectors=vec, offsets=off, id_map=idm, centroids=cent)
        try:
            p = gathered[r].data_ptr()
            idx.search_dev(q_dev.data_ptr(), nq, k, nprobe, p, p + 4 * nq * k, p + B, sptr)
            torch.cuda.synchronize()
        finally:
            idx.close()
    oi = torch.empty((nq, k), dtype=torch.int32, device=dev)
    osc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    vsb.merge_blocks_dev(gathered.data_ptr(), G, slot, nq, k, False, oi.data_ptr(), osc.data_ptr(), 0, sptr)
    cnt = gathered[:, B:B + 4 * nq].contiguous().view(torch.int32).sum(0).clamp(max=k)
    torch.cuda.synchronize()
    coarse = oracle.ivf_coarse(qry, cent)
    wi, ws, wc, _ = oracle.ivf_search(base[order], offsets, order, True, coarse, qry, k, nprobe, mode=1)
    assert np.array_equal(oi.cpu().numpy(), wi) and np.array_equal(osc.cpu().numpy(), ws)
    assert np.array_equal(cnt.cpu().numpy(), wc)
    # ---- INT8
    n, nq, k = 50_001, 64, 10
    base = vsb.synth.make("sift", 808, n)
    qry = vsb.synth.make("sift", 809, nq)
    q_dev = torch.from_numpy(qry).to(dev)
    w_scale = float(np.float32(base.max()) / np.float32(255.0))
    B = vsb.topk_block_bytes(nq, k)
    gathered = torch.zeros((G, B), dtype=torch.uint8, device=dev)
    sc8 = torch.empty((nq, k), dtype=torch.uint8, device=dev)
    for r in range(G):
        r0, r1 = sharded.shard_range(n, r, G)
        idx = vsb.Int8Index(base[r0:r1], w_scale=w_scale, id_base=r0)
        try:
            idx.search_dev(q_dev.data_ptr(), nq, k, gathered[r].data_ptr(), sc8.data_ptr(), sptr)
            gathered[r][4 * nq * k:8 * nq * k].view(torch.float32).view(nq, k).copy_(sc8)
            torch.cuda.synchronize()
        finally:
            idx.close()
    vsb.merge_blocks_dev(gathered.data_ptr(), G, B, nq, k, False, oi[:nq].data_ptr(), osc[:nq].data_ptr(), 0, sptr)
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    m = oracle.int8_multiplier(vsb.QNN_INPUT_SCALE, w_scale, vsb.QNN_OUTPUT_SCALE)
    wi, ws = oracle.int8_search(oracle.quantize_u8(base, w_scale), oracle.quantize_u8(qry, vsb.QNN_INPUT_SCALE), k, m, mode=1)
    assert np.array_equal(oi[:nq].cpu().numpy(), wi) and np.array_equal(osc[:nq].cpu().numpy().astype(np.uint8), ws)
