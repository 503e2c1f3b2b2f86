"""Timing ablation of the fused kernel (VSB_TC_DBG flags: 1 no epilogue work, 2 no candidate hand-off / inserts,
4 no MMA issue, 8 no B loads, 64 list keepers consume queue entries without merging them).
Results are meaningless under a flag, timings are not.  Usage: python tools/tc_ablation.py [N] [flags...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vsb200_loader

vsb = vsb200_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
flags = [int(x) for x in sys.argv[2:]] or [0, 2, 64, 1]
nq, k = 10_000, 10
dev = torch.device("cuda:0")
base = torch.empty((n, 128), dtype=torch.float32, device=dev)
vsb.synth_fill_dev(base.data_ptr(), 0, n, 128, "cont", 2025)
q = torch.from_numpy(vsb.synth.make("cont", 2026, nq)).to(dev)
ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
d = torch.empty((nq, k), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
idx = vsb.ExactIndex(base.data_ptr(), n=n)
idx.set_profile(True)
st = torch.cuda.Stream()
PRECS = {"f16": vsb.PREC_F16_CERT, "1x": vsb.PREC_TF32_1X, "3x": vsb.PREC_3XTF32}
for prec, name in ((PRECS[os.environ.get("PREC", "f16")], os.environ.get("PREC", "f16")),):
    for dbg in flags:
        os.environ["VSB_TC_DBG"] = str(dbg)
        ts = []
        for it in range(4):
            try:
                idx.search_dev(q.data_ptr(), nq, k, prec, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
            except Exception:
                pass
            st.synchronize()
            ts.append(idx.last_kernel_ms())
        print(f"N={n} {name} dbg={dbg:3d}  kernel ms: {np.min(ts[1:]):8.3f}", flush=True)
os.environ["VSB_TC_DBG"] = "0"
