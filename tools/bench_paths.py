"""Secondary measurements (BASELINE.json configs 2-3): batch-1 exact scan (HBM-bound), IVF nlist=1024 nprobe 8/32
list scan (HBM-bound), the GPU k-means builder.  One JSON line per path; kernel times are CUDA-event times of the
dominant kernel on its launching stream, L2 flushed between repetitions."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vsb200_loader

vsb = vsb200_loader.load()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]
BF16 = peaks.get("bf16_tflops", 1590.0)   # dense bf16 burst; TF32 proxy = half of it
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
what = sys.argv[1:] or ["batch1", "ivf", "int8"]
N, NQ, K = 1_000_000, 10_000, 10

if "batch1" in what:
    base = torch.empty((N, 128), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base.data_ptr(), 0, N, 128, "cont", 2025)
    q = torch.from_numpy(vsb.synth.make("cont", 2026, 64)).to(dev)
    qh = vsb.synth.make("cont", 2026, 64)
    ids = torch.empty((64, K), dtype=torch.int32, device=dev)
    d = torch.empty((64, K), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    idx = vsb.ExactIndex(base.data_ptr(), n=N)
    idx.set_profile(True)
    st = torch.cuda.Stream()
    for nq in (1, 2, 4, 8):
        ts, lat = [], []
        for it in range(12):
            flush.fill_(1)
            torch.cuda.synchronize()
            idx.search_dev(q.data_ptr(), nq, K, vsb.PREC_FFMA, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
            st.synchronize()
            ts.append(idx.last_kernel_ms())
        idx.set_profile(False)  # the host-buffer call then runs as ONE CUDA graph per (nq, k) (H2D + 3 kernels + D2H)
        for it in range(30):
            t0 = time.perf_counter()
            idx.search(qh[:nq], K, vsb.PREC_FFMA)
            lat.append(1e3 * (time.perf_counter() - t0))
        lat16 = []
        for it in range(30):   # the same host-buffer call through the certified fp16 path (streams the half-size fp16 base)
            t0 = time.perf_counter()
            idx.search(qh[:nq], K, vsb.PREC_F16_CERT)
            lat16.append(1e3 * (time.perf_counter() - t0))
        idx.set_profile(True)
        ms = float(np.median(ts[2:]))
        gb = N * (128 * 4 + 4) / 1e9
        print(json.dumps({"path": "exact batch-%d (exact_stream_kernel)" % nq, "kernel_ms": ms,
                          "roofline": {"bound": "hbm", "achieved": gb / (ms * 1e-3), "peak": HBM, "unit": "GB/s",
                                       "frac": gb / (ms * 1e-3) / HBM},
                          "e2e_call_ms_median": float(np.median(lat[5:])), "e2e_qps": nq / (np.median(lat[5:]) * 1e-3),
                          "e2e_call_ms_median_f16cert_path": float(np.median(lat16[5:]))}))
    idx.close()
    del base

if "ivf" in what:
    nlist = 1024
    t0 = time.time()
    base_h = vsb.synth.make("mix", 2025, N)
    t_gen = time.time() - t0
    d_ = tempfile.mkdtemp(prefix="vsb_ivf_")
    t0 = time.time()
    info = vsb.ivf_build(base_h, nlist, d_, max_iter=10, seed=42, reordered=True)
    t_build = time.time() - t0
    off = np.load(os.path.join(d_, "cluster_offsets.npy"))
    sizes = np.diff(off)
    print(json.dumps({"path": "ivf build (k-means 1Mx128, nlist 1024)", "iters": info["iters"], "inertia": info["inertia"],
                      "build_s_incl_io": t_build, "list_len_min_avg_max": [int(sizes.min()), float(sizes.mean()), int(sizes.max())]}))
    idx = vsb.IvfIndex(d_)
    idx.set_profile(True)
    qh = vsb.synth.make("mix", 2026, NQ)
    q = torch.from_numpy(qh).to(dev)
    ids = torch.empty((NQ, K), dtype=torch.int32, device=dev)
    sc = torch.empty((NQ, K), dtype=torch.float32, device=dev)
    cnt = torch.empty((NQ,), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    for nprobe in (8, 32):
        _, _, _, total = idx.search_batch(qh, K, nprobe)
        ts, tot = [], []
        for it in range(8):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            idx.search_dev(q.data_ptr(), NQ, K, nprobe, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), st.cuda_stream)
            e1.record(st)
            st.synchronize()
            ts.append(idx.last_kernel_ms())
            tot.append(e0.elapsed_time(e1))
        # host-buffer C-ABI call with PINNED host buffers (what vs_host_alloc gives a caller), median of 5
        qp = torch.from_numpy(qh).pin_memory()
        outp = (torch.empty((NQ, K), dtype=torch.int32).pin_memory().numpy(), torch.empty((NQ, K), dtype=torch.float32).pin_memory().numpy(),
                torch.empty((NQ,), dtype=torch.int32).pin_memory().numpy())
        e2es = []
        for it in range(6):
            t0 = time.perf_counter()
            idx.search_batch(qp.numpy(), K, nprobe, out=outp)
            e2es.append(time.perf_counter() - t0)
        e2e = float(np.median(e2es[1:]))
        ms = float(np.median(ts[2:]))
        gb = total * 512 / 1e9
        lm = os.environ.get("VSB_IVF_LM", "1") != "0"   # batches >= 256 queries take a list-major scan by default
        tc = lm and os.environ.get("VSB_IVF_TC", "1") != "0"
        kern = ("exact_tc_kernel<IVF> (K9, tensor-core list-major) + grouping" if tc else
                "ivf_lm_kernel (K8, FFMA list-major) + grouping" if lm else "ivf_scan_kernel (K6, query-major)")
        fp32_peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12   # FFMA lanes x 2 flop x clock
        flops = total * 256.0
        line = {"path": "ivf nlist=1024 nprobe=%d top-10, 10K queries (%s)" % (nprobe, kern), "kernel_ms": ms,
                "search_ms_all_kernels": float(np.median(tot[2:])), "qps_device": NQ / (np.median(tot[2:]) * 1e-3),
                "qps_e2e_host_buffers": NQ / e2e, "rows_scanned": total, "algorithmic_bytes": total * 512,
                "note": "kernel_ms = pair grouping + scan (+ per-query counts); search_ms_all_kernels adds the coarse scores, "
                        "probe selection and the merge / re-score; L2 flushed before every timed call"}
        if tc:   # integer-valued `mix` data: vectors and queries are TF32-exact, one TF32 product
            line["roofline"] = {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": BF16 / 2.0, "unit": "TFLOP/s",
                                "frac": flops / (ms * 1e-3) / 1e12 / (BF16 / 2.0),
                                "note": "256 flop per probed (query, row) pair, one TF32 product; the scan is bound by per-item "
                                        "latency (lists average 7.6 tiles), not by the tensor pipe; a list is read once per 128 "
                                        "pairs, so DRAM bytes are far below the algorithmic bytes (no HBM fraction is quoted)"}
        elif lm:
            line["roofline"] = {"bound": "fp32 FFMA", "achieved": flops / (ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                "frac": flops / (ms * 1e-3) / 1e12 / fp32_peak}
        else:
            line["roofline"] = {"bound": "hbm", "achieved": gb / (ms * 1e-3), "peak": HBM, "unit": "GB/s",
                                "frac": gb / (ms * 1e-3) / HBM,
                                "note": "algorithmic bytes = probed rows x 512 B per query (SURVEY.md 8d); lists shared between "
                                        "queries hit L2, so frac can exceed the DRAM share"}
        print(json.dumps(line))
    idx.close()

if "int8" in what:
    # BASELINE configs[3]: 10M x 128 u8 base (1.28 GB), batches of 32 and 1024 queries, top-10
    NI = 10_000_000
    base = torch.empty((NI, 128), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base.data_ptr(), 0, NI, 128, "mix", 2025)
    torch.cuda.synchronize()
    idx = vsb.Int8Index(base.data_ptr(), 218.0 / 255.0, 218.0 / 255.0, 9000.0, n=NI)
    del base
    torch.cuda.empty_cache()
    idx.set_profile(True)
    peak_bf16 = peaks.get("bf16_tflops", 1590.0)
    st = torch.cuda.Stream()
    for nq in (32, 1024):
        qh = vsb.synth.make("mix", 2026, nq)
        q = torch.from_numpy(qh).to(dev)
        ids = torch.empty((nq, K), dtype=torch.int32, device=dev)
        sc = torch.empty((nq, K), dtype=torch.uint8, device=dev)
        ts, tot = [], []
        for it in range(10):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            idx.search_dev(q.data_ptr(), nq, K, ids.data_ptr(), sc.data_ptr(), st.cuda_stream)
            e1.record(st)
            st.synchronize()
            ts.append(idx.last_kernel_ms())
            tot.append(e0.elapsed_time(e1))
        lat = []
        for it in range(10):
            t0 = time.perf_counter()
            idx.search(qh, K)
            lat.append(time.perf_counter() - t0)
        ms = float(np.median(ts[2:]))
        line = {"path": "int8 brute force 10Mx128 batch-%d top-10 (%s)" % (nq, "int8_tc_pair_kernel" if nq > 128 and os.environ.get("VSB_INT8_PAIR", "1") != "0" else "int8_tc_kernel"), "kernel_ms": ms,
                "search_ms_all_kernels": float(np.median(tot[2:])), "qps_device": nq / (np.median(tot[2:]) * 1e-3),
                "qps_e2e_host_buffers": nq / float(np.median(lat[2:]))}
        gb = NI * 128 / 1e9
        tops = 2.0 * nq * NI * 128 / 1e12
        line["roofline_hbm"] = {"bound": "hbm", "achieved": gb / (ms * 1e-3), "peak": HBM, "unit": "GB/s", "frac": gb / (ms * 1e-3) / HBM}
        line["roofline_tensor"] = {"bound": "tensor", "achieved": tops / (ms * 1e-3), "peak": 2 * peak_bf16, "unit": "TOP/s",
                                   "frac": tops / (ms * 1e-3) / (2 * peak_bf16), "issued_frac_m128_tiles": (2.0 * ((nq + 127) // 128 * 128) * NI * 128 / 1e12) / (ms * 1e-3) / (2 * peak_bf16),
                                   "peak_note": "int8 dense = 2 x MEASURED_PEAKS bf16 burst"}
        print(json.dumps(line))
    idx.close()
