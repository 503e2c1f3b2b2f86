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
