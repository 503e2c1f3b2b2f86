"""Debug: insertion counters of the fused kernel (VSB_TC_STATS, printed by the library to stderr)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
st = torch.cuda.Stream()
os.environ["VSB_TC_STATS"] = "1"
for prec in (vsb.PREC_F16_CERT, vsb.PREC_TF32_1X):
    idx.search_dev(q.data_ptr(), NQ, K, prec, ids.data_ptr(), d.data_ptr(), st.cuda_stream)
    st.synchronize()
idx.close()
