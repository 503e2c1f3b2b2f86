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
