"""BASELINE.json configs[4]: 100M x 128 fp32 exact top-100, base rows sharded across the GPUs of one box, NCCL
all-gather of the local top-k + merge (vs_merge_topk_dev).  Run as
    python tools/bench_config5.py [--rows 100000000] [--nq 1000] [--k 100] [--law sift]
or under torchrun with one rank per GPU.  Prints ONE JSON line (rank 0).

Checks (no CPU oracle at this size: the base never exists on the host):
  * every result row is sorted by (distance, id), ids are unique and in range;
  * for a few queries, a plain torch fp32 matmul over the same shard(s) + torch.topk gives the same distances
    (within 1e-5 relative) — an independent statement of the same arithmetic on the same device.
"""
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
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--nq", type=int, default=1000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--law", default="sift")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", type=int, default=4, help="queries cross-checked against torch fp32")
    a = ap.parse_args()
    rank, world, lrank = (int(os.environ.get(x, d)) for x, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    vsb = vsb200_loader.load()
    from vsb200 import sharded

    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r0, r1 = sharded.shard_range(a.rows, rank, world)
    n_loc = r1 - r0
    base = torch.empty((n_loc, 128), dtype=torch.float32, device=dev)
    t0 = time.time()
    CH = 1 << 22
    for c0 in range(0, n_loc, CH):  # chunked: the generator takes int32-sized launches comfortably
        n = min(CH, n_loc - c0)
        vsb.synth_fill_dev(base.data_ptr() + c0 * 512, r0 + c0, n, 128, a.law, 31337)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    index = vsb.ExactIndex(base.data_ptr(), device=lrank, id_base=r0, n=n_loc)
    t_build = time.time() - t0
    q = torch.from_numpy(vsb.synth.make(a.law, 31338, a.nq)).to(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    s = sharded.ShardedExact(vsb, index, a.nq, a.k, dev)

    def step():
        return s.search(q.data_ptr(), a.nq, vsb.PREC_AUTO, stream.cuda_stream)

    ids, d = step()
    stream.synchronize()
    ts = []
    for _ in range(a.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ids, d = step()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches, prec_used = index.last_launches()

    # ---- checks
    ids_h, d_h = ids.cpu().numpy(), d.cpu().numpy()
    ok_sorted = bool(np.all((d_h[:, 1:] > d_h[:, :-1]) | ((d_h[:, 1:] == d_h[:, :-1]) & (ids_h[:, 1:] > ids_h[:, :-1]))))
    ok_range = bool(ids_h.min() >= 0 and ids_h.max() < a.rows)
    ok_unique = all(len(set(r.tolist())) == a.k for r in ids_h[:64])
    # torch fp32 reference for a few queries: local top-k per shard, gathered and merged on the host
    nc = min(a.check, a.nq)
    qn = (q[:nc] * q[:nc]).sum(1, keepdim=True)
    best = None
    for c0 in range(0, n_loc, 1 << 21):
        blk = base[c0:c0 + (1 << 21)]
        dd = qn + (blk * blk).sum(1)[None, :] - 2.0 * (q[:nc] @ blk.T)
        v, i = torch.topk(dd, min(a.k, blk.shape[0]), dim=1, largest=False)
        i = i + (r0 + c0)
        if best is None:
            best = (v, i)
        else:
            v2, sel = torch.topk(torch.cat([best[0], v], 1), a.k, dim=1, largest=False)
            best = (v2, torch.gather(torch.cat([best[1], i], 1), 1, sel))
    ref_v = best[0]
    if world > 1:
        allv = [torch.empty_like(ref_v) for _ in range(world)]
        dist.all_gather(allv, ref_v.contiguous())
        ref_v = torch.topk(torch.cat(allv, 1), a.k, dim=1, largest=False)[0]
    ref_v = torch.sort(ref_v, dim=1)[0].cpu().numpy()
    rel = float(np.max(np.abs(ref_v - d_h[:nc]) / np.maximum(np.abs(ref_v), 1e-30)))
    if rank == 0:
        flop = 2.0 * a.nq * a.rows * 128
        print(json.dumps({
            "path": f"exact top-{a.k} over {a.rows}x128 fp32 ({a.law}), {a.nq} queries, {world} GPU(s), rows sharded",
            "ms_per_batch": ms, "qps": a.nq / (ms * 1e-3), "algorithmic_tflops": flop / (ms * 1e-3) / 1e12,
            "precision": vsb.PREC_NAMES[prec_used], "launches_per_rank": launches,
            "rows_per_gpu": n_loc, "gen_s": t_gen, "index_build_s": t_build,
            "checks": {"sorted_by_dist_id": ok_sorted, "ids_in_range": ok_range, "ids_unique": ok_unique,
                       "torch_fp32_max_rel_diff_of_distances": rel, "queries_checked": nc}}))
        assert ok_sorted and ok_range and ok_unique and rel <= 1e-5, "config-5 checks failed"
    index.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
