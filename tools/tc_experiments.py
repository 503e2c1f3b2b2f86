"""Timing experiments on the fused tensor-core kernel: component costs via VSB_TC_DBG flags
(1 no epilogue work, 2 no inserts, 4 no MMA issue, 8 no B loads). Results are meaningless, timings are not."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vsb200_loader

vsb = vsb200_loader.load()
n, nq, k = 1_000_000, 10_000, 10
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
for prec, name in ((vsb.PREC_TF32_1X, "1x"), (vsb.PREC_3XTF32, "3x")):
    for dbg in (0, 2, 1, 4, 5, 8, 9, 12, 13):
        os.environ["VSB_TC_DBG"] = str(dbg)
        ts = []
        for it in range(4):
            idx.search_dev(q.data_ptr(), nq, k, prec, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
            st.synchronize()
            ts.append(idx.last_kernel_ms())
        print(f"{name} dbg={dbg:2d}  kernel ms: {np.min(ts[1:]):8.3f}   (flags: "
              f"{'noepi ' if dbg & 1 else ''}{'noinsert ' if dbg & 2 else ''}{'nomma ' if dbg & 4 else ''}{'noload' if dbg & 8 else ''})")
os.environ["VSB_TC_DBG"] = "0"
