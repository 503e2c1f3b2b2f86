#!/usr/bin/env python
"""bench.py — headline benchmark: QPS of exact L2 top-10 on a 1M x 128 fp32 base (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: 10 000 queries against the whole base, top-10, fp32-faithful.
Default precision AUTO = certified fp16 tcgen05 candidate generation (sample pass over 1/16 of the base tiles ->
per-query key threshold -> threshold-filter pass over the whole base) + exact fp32 refine of the <= 32 best
candidates + per-query certificate (uncertified queries are redone in 3xTF32); the pure 3xTF32 tcgen05 path is timed
in the same run and printed beside it (`fp32_3xtf32_path`).
N > 1 (torchrun, one rank per GPU): base rows sharded across ranks, queries replicated (every rank uploads 1/N of
them, one all-gather replicates the slices over NVLink), per-rank local top-k written in place into the rank's slot
of the gathered buffer, ONE exchange of the blocks (ids | dists | uncertified count) — by default every rank pushes its
block into the peers' symmetric-memory buffers with one kernel over NVLink and a device-side barrier follows
(VSB_EXCHANGE=nccl: one in-place NCCL all-gather) — + merge kernel: strong scaling on the fixed 1M x 128 problem.

Prints ONE JSON line (rank 0):
  value        whole-job QPS, queries already resident in HBM when the timed region starts
  e2e          the same through the host-buffer C-ABI call (H2D of queries + D2H of results inside the timing)
  roofline     dominant kernel (the fused distance + filter pass over the whole base) timed live with CUDA events on its
               launching stream; `prepass_ms` = the sample pass + threshold selection that run before it
  cpu_baseline the UNMODIFIED reference (oracle/_ref) timed on this box's host cores on a bounded query sample
--impl reference times that reference as the main metric (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BASE = 1_000_000
N_QUERY = 10_000
DIM = 128
TOPK = 10
LAW = "cont"  # continuous-valued SIFT-shaped synthetic data: exercises the real 3xTF32 split (lo != 0)
BASE_SEED, QUERY_SEED = 2025, 2026
METRIC = "QPS exact top-10 on 1Mx128"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def pause(self, on):
        """The polling nvidia-smi takes the driver lock every 20 ms, which delays HOST-side CUDA calls by up to a millisecond:
        harmless for the device-timed loops (CUDA events), but it would be charged to the wall-clock e2e loop.  The sampler
        therefore sleeps (SIGSTOP) while the e2e steps are timed."""
        import signal
        if self.proc and self.proc.poll() is None:
            try:
                os.kill(self.proc.pid, signal.SIGSTOP if on else signal.SIGCONT)
            except OSError:
                pass

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_dram_bytes(path):
    """dram__bytes_read.sum + dram__bytes_write.sum of a tools/ncu_summary.py excerpt, or None."""
    try:
        tot, seen = 0.0, 0
        for ln in open(path):
            f = ln.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(f[2], None)
                if mult is None:
                    return None
                tot += float(f[1]) * mult
                seen += 1
        return tot if seen == 2 else None
    except OSError:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained"),
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


# ----------------------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED cpu_baseline.cpp (oracle/_ref/ref_driver -> run_benchmark), bounded sample
# ----------------------------------------------------------------------------------------------------------
def reference_run(vsb, n_queries, repeats, threads, keep_dir=None):
    """Times run_benchmark() (its own timed span: the query loop, cpu_baseline.cpp:220-257) on the bench base
    and `n_queries` of the bench queries. Returns list of QPS values (one per repeat)."""
    from oracle import oracle

    if not oracle.have_ref():
        return None
    td = keep_dir or tempfile.mkdtemp(prefix="vsb_ref_")
    bf, qf, rt = (os.path.join(td, x) for x in ("base.fvecs", "query.fvecs", "results.txt"))
    if not os.path.exists(bf):
        out = np.empty((N_BASE, DIM + 1), dtype=np.float32)
        out[:, 0] = np.array([DIM], dtype=np.int32).view(np.float32)[0]
        for r0 in range(0, N_BASE, 1 << 16):
            n = min(1 << 16, N_BASE - r0)
            out[r0:r0 + n, 1:] = vsb.synth.rows(LAW, BASE_SEED, r0, n)
        out.tofile(bf)
        del out
    vsb.synth.write_fvecs(qf, vsb.synth.rows(LAW, QUERY_SEED, 0, n_queries))
    res = []
    for _ in range(repeats):
        r = oracle.ref_bench(bf, qf, TOPK, rt, threads=threads)
        res.append(r["qps"])
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "f16cert", "3xtf32", "1xtf32", "ffma"])
    ap.add_argument("--nq", type=int, default=N_QUERY)
    ap.add_argument("--cpu-sample", type=int, default=128, help="queries timed on the CPU reference (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    ncores = os.cpu_count() or 1

    import vsb200_loader
    vsb = vsb200_loader.load()

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = min(args.cpu_sample, args.nq)
        qps = reference_run(vsb, sample, args.warmup + args.steps, ncores)
        if qps is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver was not built"}))
            return 0
        qps = qps[args.warmup:]
        v = float(np.mean(qps))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / v,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{N_BASE}x{DIM} fp32 base, exact L2 top-{TOPK}, law={LAW}",
                           "queries_per_step": sample},
                "cpu_baseline": {"value": v, "unit": "queries/s", "cores": ncores, "kind": "reference",
                                 "sample": f"{sample} of the {args.nq} bench queries per step against the full "
                                           f"{N_BASE}x{DIM} base; the reference's own timed span "
                                           f"(cpu_baseline.cpp:220-257)"},
                "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available() or vsb.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libvsb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    prec = {"3xtf32": vsb.PREC_3XTF32, "auto": vsb.PREC_AUTO, "1xtf32": vsb.PREC_TF32_1X, "ffma": vsb.PREC_FFMA,
            "f16cert": vsb.PREC_F16_CERT}[args.precision]
    nq, k = args.nq, TOPK

    # base shard of this rank, generated on the device (bit-identical to the numpy generator)
    from vsb200 import sharded as vsb_sharded

    r0, r1 = vsb_sharded.shard_range(N_BASE, rank, world)
    base_d = torch.empty((r1 - r0, DIM), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base_d.data_ptr(), r0, r1 - r0, DIM, LAW, BASE_SEED)
    torch.cuda.synchronize()
    index = vsb.ExactIndex(base_d.data_ptr(), device=local_rank, id_base=r0, n=r1 - r0)
    index.set_profile(True)

    q_host = torch.from_numpy(vsb.synth.make(LAW, QUERY_SEED, nq)).pin_memory()
    q_dev = q_host.to(dev)
    ids_h = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    d_h = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # a non-default stream: its handle is what the C ABI launches on and what the CUDA events are recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    searcher = vsb_sharded.ShardedExact(vsb, index, nq, k, dev)

    def device_step(q_ptr, after_enqueue=None):
        # local fused search; for N > 1: ONE in-place NCCL all-gather of the exchange blocks + merge kernel.  Everything
        # the step does on the GPU is enqueued by enqueue(); finish() is the host-side check of the 4-byte total of the
        # uncertified counts (0 on this data: nothing is redone), it enqueues nothing in that case
        out = searcher.enqueue(q_ptr, nq, prec, sptr)
        if after_enqueue is not None:
            after_enqueue()
        assert searcher.finish() == 0
        return out

    # N > 1: every rank uploads ITS 1/N of the queries (they cross PCIe once in total) and one all-gather replicates
    # the slices over NVLink
    q_rows = (nq + world - 1) // world
    q_stage = torch.empty((q_rows * world, DIM), dtype=torch.float32, device=dev)
    q_r0, q_r1 = min(nq, q_rows * rank), min(nq, q_rows * (rank + 1))

    def e2e_step():
        if world == 1:
            # the reference-facing C-ABI call with HOST buffers: H2D + search + D2H inside
            index.search(q_host.numpy(), k, prec, out_ids=ids_h.numpy(), out_dists=d_h.numpy())
        else:
            if q_r1 > q_r0:
                q_stage[q_r0:q_r1].copy_(q_host[q_r0:q_r1], non_blocking=True)
            dist.all_gather_into_tensor(q_stage, q_stage[q_rows * rank:q_rows * (rank + 1)])   # in place
            oi, od = searcher.enqueue(q_stage.data_ptr(), nq, prec, sptr)
            if rank == 0:
                ids_h.copy_(oi, non_blocking=True)
                d_h.copy_(od, non_blocking=True)
            if searcher.finish() and rank == 0:   # rare: something was redone and merged again
                ids_h.copy_(oi, non_blocking=True)
                d_h.copy_(od, non_blocking=True)
            stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, use_events=True):
        """per-step timing with an L2 flush between steps; returns list of ms"""
        out = []
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            if use_events:
                # device time of the step: the closing event is recorded right after the step's last GPU operation was
                # enqueued (before the host waits for anything)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn(lambda: e1.record(stream))
                e1.synchronize()
                out.append(e0.elapsed_time(e1))
            else:
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                out.append(1e3 * (time.perf_counter() - t0))
        return out

    for _ in range(args.warmup):
        device_step(q_dev.data_ptr())
        e2e_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kernel_ms, prepass_ms = [], []

    def dev_fn(after_enqueue=None):
        device_step(q_dev.data_ptr(), after_enqueue)

    t_dev = []
    for _ in range(args.steps):
        t_dev += timed(dev_fn, 1)
        kernel_ms.append(index.last_kernel_ms())
        prepass_ms.append(index.last_prepass_ms())
    launches, prec_used = index.last_launches()
    fallbacks = index.last_fallbacks()
    if rank == 0:
        sampler.pause(True)
    t_e2e = timed(e2e_step, args.steps, use_events=False)  # wall clock: the host-buffer call blocks the host
    if rank == 0:
        sampler.pause(False)
    barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev = max_over_ranks(float(np.sum(t_dev))) / args.steps
    ms_e2e = max_over_ranks(float(np.sum(t_e2e))) / args.steps
    ms_kernel = max_over_ranks(float(np.mean(kernel_ms)))
    ms_prepass = max_over_ranks(float(np.mean(prepass_ms)))

    # the pure fp32-faithful tensor-core path (3xTF32 split, no fp16 candidate pass) timed in the same run, for reference
    alt = None
    if prec == vsb.PREC_AUTO:
        def alt_fn(after_enqueue=None):
            searcher.enqueue(q_dev.data_ptr(), nq, vsb.PREC_3XTF32, sptr)
            if after_enqueue is not None:
                after_enqueue()
            searcher.finish()
        for _ in range(3):
            alt_fn()
        barrier()
        ms_alt = max_over_ranks(float(np.sum(timed(alt_fn, args.steps)))) / args.steps
        alt = {"precision": vsb.PREC_NAMES[vsb.PREC_3XTF32], "value": nq / (ms_alt * 1e-3), "unit": "queries/s",
               "ms_per_step": ms_alt, "note": "same workload through VS_PREC_FP32_3XTF32 only (device-resident queries)"}

    clocks = sampler.stop() if rank == 0 else None  # sampled over the device-timed loops (headline and 3xTF32), not the e2e loop

    # sanity: the timed path produced a plausible answer (ascending distances, ids in range)
    oi, od = device_step(q_dev.data_ptr())
    torch.cuda.synchronize()
    assert bool((od[:, 1:] >= od[:, :-1]).all()) and int(oi.min()) >= 0 and int(oi.max()) < N_BASE

    if rank == 0:
        pk = peaks()
        split = 3 if prec_used == vsb.PREC_3XTF32 else 1
        n_local = (N_BASE + world - 1) // world
        if prec_used == vsb.PREC_FFMA:
            passes = (nq + 7) // 8
            alg_bytes = n_local * (DIM * 4 + 4)  # first launch: one pass over the shard
            roof = {"bound": "hbm", "achieved": alg_bytes / (ms_kernel * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                    "unit": "GB/s", "traffic": None, "kernel": "exact_stream_kernel (first of %d passes)" % passes}
        elif prec_used == vsb.PREC_F16_CERT:
            # one fp16 product per (query, row, dim): 2*Q*N*128 flop on the f16/bf16 tensor pipe (DESIGN.md "Roofline")
            flops = 2.0 * nq * n_local * DIM
            roof = {"bound": "tensor", "achieved": flops / (ms_kernel * 1e-3) / 1e12, "peak": pk["bf16"],
                    "unit": "TFLOP/s", "traffic": None, "kernel": "exact_tc_kernel<32, F16> (threshold-filter pass)",
                    "tf32_products": 0, "algorithmic_fp32_tflops": flops / (ms_kernel * 1e-3) / 1e12,
                    "prepass_ms": ms_prepass,
                    "frac_incl_prepass": flops / ((ms_kernel + ms_prepass) * 1e-3) / 1e12 / pk["bf16"],
                    "prepass_note": "exact_tc_kernel<1, F16> sample pass over one base tile in 16 + tc_select_thr_kernel "
                                    "(per-query thresholds); their flops are not counted as algorithmic work",
                    "peak_note": f"fp16/bf16 dense = MEASURED_PEAKS bf16 burst ({pk['src']})"}
        else:
            # tensor work issued by the fused kernel for this rank's shard: 2*Q*N*128 flop per TF32 product, 3
            # products for the fp32-faithful split (DESIGN.md "Roofline")
            flops = 2.0 * nq * n_local * DIM * split
            tf32_peak = pk["bf16"] / 2.0
            roof = {"bound": "tensor", "achieved": flops / (ms_kernel * 1e-3) / 1e12, "peak": tf32_peak,
                    "unit": "TFLOP/s", "traffic": None, "kernel": "exact_tc_kernel",
                    "tf32_products": split, "algorithmic_fp32_tflops": 2.0 * nq * n_local * DIM / (ms_kernel * 1e-3) / 1e12,
                    "peak_note": f"TF32 dense = MEASURED_PEAKS bf16 burst / 2 ({pk['src']})"}
        if world == 1 and nq == N_QUERY:
            # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture of the same command
            prof = {vsb.PREC_F16_CERT: "r2_ncu_full_exact_tc_f16_filter.txt", vsb.PREC_3XTF32: "r1f_ncu_full_exact_tc_3xtf32.txt"}.get(prec_used)
            roof["traffic"] = ncu_dram_bytes(os.path.join(ROOT, "profiles", prof)) if prof else None
            roof["traffic_unit"] = "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/%s)" % prof if prof else None
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["kernel_ms"] = ms_kernel
        line = {
            "metric": METRIC, "value": nq / (ms_dev * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{N_BASE}x{DIM} fp32 base, {nq} queries/step, exact L2 top-{k}, law={LAW}",
                       "precision": vsb.PREC_NAMES[prec_used], "uncertified_queries_redone_in_fp32": fallbacks,
                       "precision_note": "the fp16 tensor-core passes only PROPOSE candidates (every row whose fp16 key lies below a "
                                         "per-query threshold; the <= 32 best are kept); every returned distance is recomputed in "
                                         "fp32 and the top-k is certified complete per query against the bound actually used (else "
                                         "redone in 3xTF32): results equal the fp32 path's (tests/test_exact_gpu.py)" if prec_used == vsb.PREC_F16_CERT else "",
                       "base_rows_per_gpu": n_local,
                       "parallelism": (f"base rows sharded x{world}, queries replicated, ONE exchange of the (ids|dists|count) blocks ["
                                       + ("push: every rank stores its block into the peers' symmetric-memory buffers over NVLink "
                                          "(vs_push_block_dev) + device-side barrier" if searcher.exchange_kind == "push"
                                          else "in-place NCCL all-gather: " + searcher.exchange_kind) + "] + merge") if world > 1 else "single GPU",
                       "cache": "L2 flushed (256 MB write) between timed steps; operands (0.5-1 GB) exceed the 126 MB L2"},
            "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(q_host.numel() * 4),
                    "d2h_bytes_per_step": int(nq * k * 8),
                    "api": "vs_exact_search_f32 (host buffers)" if world == 1 else "H2D of 1/N of the queries per rank + all-gather of the slices + vs_exact_group_begin + ONE exchange of the blocks (see config.parallelism) + vs_exact_group_merge + D2H + vs_exact_group_finish"},
            "gpu_launches": int(launches + (1 if world > 1 else 0)) * args.steps,
            "roofline": roof,
            "clocks": clocks,
        }
        if alt is not None:
            line["fp32_3xtf32_path"] = alt
        if not args.no_cpu and world == 1:
            t0 = time.time()
            sample = min(args.cpu_sample, nq)
            qps = reference_run(vsb, sample, 1, ncores)
            if qps is not None:
                line["cpu_baseline"] = {"value": float(qps[0]), "unit": "queries/s", "cores": ncores, "kind": "reference",
                                        "sample": f"unmodified cpu_baseline.cpp run_benchmark(): {sample} of the {nq} bench "
                                                  f"queries against the full {N_BASE}x{DIM} base, its own timed span "
                                                  f"(query loop only); wall {time.time() - t0:.0f}s incl. file I/O"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": ncores, "kind": "reference",
                                        "sample": "oracle/_ref/ref_driver missing"}
        print(json.dumps(line))
    index.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
