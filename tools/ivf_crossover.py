"""IVF fine scan: query-major (K6) vs list-major (K8) whole-search time per batch size (decides the AUTO rule).
1M x 128 mixture-law base, nlist 1024, top-10."""
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
N, K, NLIST = 1_000_000, 10, 1024
dev = torch.device("cuda:0")
base_h = vsb.synth.make("mix", 2025, N)
d_ = tempfile.mkdtemp(prefix="vsb_ivf_")
vsb.ivf_build(base_h, NLIST, d_, max_iter=10, seed=42, reordered=True)
idx = vsb.IvfIndex(d_)
qh = vsb.synth.make("mix", 2026, 4096)
q = torch.from_numpy(qh).to(dev)
ids = torch.empty((4096, K), dtype=torch.int32, device=dev)
sc = torch.empty((4096, K), dtype=torch.float32, device=dev)
cnt = torch.empty((4096,), dtype=torch.int32, device=dev)
st = torch.cuda.Stream()
print("nq nprobe  K6_ms  K8_ms   (whole search_dev call, median of 8)")
for nprobe in (8, 32):
    for nq in (32, 64, 128, 256, 512, 1024, 2048, 4096):
        row = []
        for lm in ("0", "1"):
            os.environ["VSB_IVF_LM"] = lm
            ts = []
            for it in range(10):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                idx.search_dev(q.data_ptr(), nq, K, nprobe, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), st.cuda_stream)
                st.synchronize()
                ts.append(1e3 * (time.perf_counter() - t0))
            row.append(float(np.median(ts[2:])))
        print(f"{nq:5d} {nprobe:3d}  {row[0]:7.3f} {row[1]:7.3f}", flush=True)
idx.close()
