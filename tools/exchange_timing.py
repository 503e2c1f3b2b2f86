"""Where the time of one sharded step goes (run under torchrun, one rank per GPU): CUDA events on the launching stream
around (1) the local fused search (vs_exact_group_begin), (2) the ONE exchange of the blocks (push over NVLink peer memory + flag barrier, or VSB_EXCHANGE=nccl: in-place all-gather),
(3) the merge kernel + the 4-byte total, and the host-side wall time of finish().  Prints one JSON line (rank 0, max over
ranks).  Usage: torchrun --nproc-per-node N tools/exchange_timing.py [--nq 10000] [--rows 1000000] [--steps 20]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import vsb200_loader


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    rank, world, lrank = (int(os.environ.get(x, d)) for x, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    vsb = vsb200_loader.load()
    from vsb200 import sharded

    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r0, r1 = sharded.shard_range(a.rows, rank, world)
    base = torch.empty((r1 - r0, 128), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base.data_ptr(), r0, r1 - r0, 128, "cont", 2025)
    torch.cuda.synchronize()
    index = vsb.ExactIndex(base.data_ptr(), device=lrank, id_base=r0, n=r1 - r0)
    q = torch.from_numpy(vsb.synth.make("cont", 2026, a.nq)).to(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    s = sharded.ShardedExact(vsb, index, a.nq, a.k, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    acc = np.zeros(5)
    for it in range(a.steps + 3):
        flush.fill_(1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        if world > 1:
            if s._sym is not None:  # push exchange: next half of the symmetric buffer (what ShardedExact.enqueue does)
                s._parity ^= 1
                s._stream = stream.cuda_stream
                s.gathered = s._sym[s._parity * s._half:(s._parity + 1) * s._half]
            s.grp.begin(q.data_ptr(), a.nq, a.k, vsb.PREC_AUTO, s.gathered.data_ptr(), stream.cuda_stream)
            ev[1].record(stream)
            s._exchange()
            ev[2].record(stream)
            s.grp.merge(s.ids_loc.data_ptr(), s.d_loc.data_ptr())
            ev[3].record(stream)
            t0 = time.perf_counter()
            again = s.grp.finish()
            t_fin = 1e3 * (time.perf_counter() - t0)
            assert not again
        else:
            s.enqueue(q.data_ptr(), a.nq, vsb.PREC_AUTO, stream.cuda_stream)
            for e in ev[1:]:
                e.record(stream)
            t_fin = 0.0
        stream.synchronize()
        if it >= 3:
            acc += np.array([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]),
                             ev[0].elapsed_time(ev[3]), t_fin])
    t = torch.tensor(acc / a.steps, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        v = t.cpu().numpy()
        print(json.dumps({"n_gpus": world, "rows_per_gpu": r1 - r0, "nq": a.nq, "k": a.k,
                          "exchange": s.exchange_kind,
                          "ms_local_search_begin": v[0], "ms_exchange_blocks": v[1], "ms_merge_and_total": v[2],
                          "ms_step_device": v[3], "ms_finish_host_wall": v[4],
                          "block_bytes_per_rank": vsb.topk_block_bytes(a.nq, a.k),
                          "note": "CUDA events on the launching stream, max over ranks, L2 flushed between steps"}))
    index.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
