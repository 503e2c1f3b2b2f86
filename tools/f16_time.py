"""Kernel time of the certified fp16 candidate generation (sample + threshold + filter pass, CUDA events of the library)
and whole-call time at the shard sizes of interest.  VSB200_LIB selects a build variant.  Usage: python tools/f16_time.py [law] [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vsb200_loader

vsb = vsb200_loader.load()
law = sys.argv[1] if len(sys.argv) > 1 else "cont"
sizes = [int(x) for x in sys.argv[2:]] or [1_000_000, 125_000]
NQ, K = 10_000, 10
dev = torch.device("cuda:0")
q = torch.from_numpy(vsb.synth.make(law, 2026, NQ)).to(dev)
ids = torch.empty((NQ, K), dtype=torch.int32, device=dev)
d = torch.empty((NQ, K), dtype=torch.float32, device=dev)
st = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for N in sizes:
    base = torch.empty((N, 128), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base.data_ptr(), 0, N, 128, law, 2025)
    torch.cuda.synchronize()
    idx = vsb.ExactIndex(base.data_ptr(), n=N)
    idx.set_profile(True)
    ks, ws = [], []
    for _ in range(8):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            idx.search_dev(q.data_ptr(), NQ, K, vsb.PREC_F16_CERT, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
            e1.record(st)
        st.synchronize()
        ks.append(idx.last_kernel_ms())
        ws.append(e0.elapsed_time(e1))
    out.append(f"N={N}: candidates {np.min(ks[2:]):.3f} ms, call {np.min(ws[2:]):.3f} ms, uncertified {idx.last_fallbacks()}")
    idx.close()
    del base
print(os.environ.get("VSB200_LIB", "default").split("/")[-1], os.environ.get("VSB_F16_STRIDE", "-"), os.environ.get("VSB_F16_M", "-"),
      os.environ.get("VSB_TC_DBG", "-"), " | ".join(out), flush=True)
