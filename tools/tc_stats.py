"""Debug: insertion counters (VSB_TC_STATS, printed by the library to stderr) and kernel time of the fused kernel for
the candidate-list sizes / shard sizes of interest.  Usage: python tools/tc_stats.py [law]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vsb200_loader

vsb = vsb200_loader.load()
law = sys.argv[1] if len(sys.argv) > 1 else "cont"
NQ, K = 10_000, 10
dev = torch.device("cuda:0")
q = torch.from_numpy(vsb.synth.make(law, 2026, NQ)).to(dev)
ids = torch.empty((NQ, K), dtype=torch.int32, device=dev)
d = torch.empty((NQ, K), dtype=torch.float32, device=dev)
st = torch.cuda.Stream()
for N in (1_000_000, 500_000, 125_000):
    base = torch.empty((N, 128), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base.data_ptr(), 0, N, 128, law, 2025)
    torch.cuda.synchronize()
    idx = vsb.ExactIndex(base.data_ptr(), n=N)
    idx.set_profile(True)
    for prec, name in ((vsb.PREC_F16_CERT, "f16"), (vsb.PREC_TF32_1X, "1x")):
        for ktop in (0,):
            os.environ.pop("VSB_TC_STATS", None)
            ts = []
            for _ in range(5):
                idx.search_dev(q.data_ptr(), NQ, K, prec, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
                st.synchronize()
                ts.append(idx.last_kernel_ms())
            print(f"law={law} N={N} {name} ktop={ktop}: kernel ms {np.min(ts[1:]):.3f}  uncertified {idx.last_fallbacks()}",
                  file=sys.stderr, flush=True)
            os.environ["VSB_TC_STATS"] = "1"
            idx.search_dev(q.data_ptr(), NQ, K, prec, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
            st.synchronize()
    idx.close()
    del base
