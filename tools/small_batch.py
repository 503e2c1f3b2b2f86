"""Whole-call device time of an exact top-10 search at small batch sizes, per precision path (decides the AUTO
thresholds).  Wall clock around search_dev + stream sync, L2 flushed between calls."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vsb200_loader

vsb = vsb200_loader.load()
N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 10
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
base = torch.empty((N, 128), dtype=torch.float32, device=dev)
vsb.synth_fill_dev(base.data_ptr(), 0, N, 128, "cont", 2025)
NQMAX = 4096
q = torch.from_numpy(vsb.synth.make("cont", 2026, NQMAX)).to(dev)
ids = torch.empty((NQMAX, K), dtype=torch.int32, device=dev)
d = torch.empty((NQMAX, K), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
idx = vsb.ExactIndex(base.data_ptr(), n=N)
st = torch.cuda.Stream()
print(f"N={N}  ms per call (median of 10): nq, ffma, f16cert, 3xtf32")
for nq in (1, 2, 4, 8, 9, 16, 32, 128, 256, 384, 512, 768, 1024, 2048, 4096):
    row = []
    for prec in (vsb.PREC_FFMA, vsb.PREC_F16_CERT, vsb.PREC_3XTF32):
        if prec == vsb.PREC_FFMA and nq > 32:
            row.append(float("nan"))
            continue
        ts = []
        for it in range(12):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            idx.search_dev(q.data_ptr(), nq, K, prec, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
            st.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        row.append(float(np.median(ts[2:])))
    print(f"{nq:4d}  {row[0]:8.3f} {row[1]:8.3f} {row[2]:8.3f}", flush=True)
idx.close()
