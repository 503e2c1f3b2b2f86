"""Timing ablations of the fused tensor-core kernel (VSB_TC_DBG bits, see exact_tc.cuh TcParams::dbg): which of
MMA / operand loads / epilogue TMEM reads / epilogue math bounds the kernel.  Results with a bit set are WRONG by
construction; only the kernel time is read."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vsb200_loader

vsb = vsb200_loader.load()
N, NQ, K = 1_000_000, 10_000, 10
dev = torch.device("cuda:0")
base = torch.empty((N, 128), dtype=torch.float32, device=dev)
vsb.synth_fill_dev(base.data_ptr(), 0, N, 128, "cont", 2025)
q = torch.from_numpy(vsb.synth.make("cont", 2026, NQ)).to(dev)
ids = torch.empty((NQ, K), dtype=torch.int32, device=dev)
d = torch.empty((NQ, K), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
idx = vsb.ExactIndex(base.data_ptr(), n=N)
idx.set_profile(True)
st = torch.cuda.Stream()
NAMES = {1: "noepi", 2: "noinsert", 4: "nomma", 8: "noload", 16: "ldonly", 32: "mathonly"}
combos = [int(x) for x in sys.argv[1:]] or [0, 2, 1, 4, 5, 8, 16, 32, 4 | 16, 4 | 32, 4 | 8 | 16, 4 | 8 | 32, 4 | 8 | 1]
for prec, pname in ((vsb.PREC_F16_CERT, "f16"), (vsb.PREC_TF32_1X, "1x"), (vsb.PREC_3XTF32, "3x")):
    for dbg in combos:
        os.environ["VSB_TC_DBG"] = str(dbg)
        ts = []
        for it in range(5):
            idx.search_dev(q.data_ptr(), NQ, K, prec, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
            st.synchronize()
            ts.append(idx.last_kernel_ms())
        flags = " ".join(v for b, v in NAMES.items() if dbg & b)
        print(f"{pname} dbg={dbg:3d}  kernel ms: {np.median(ts[1:]):8.3f}   (flags: {flags})", flush=True)
os.environ["VSB_TC_DBG"] = "0"
idx.close()
