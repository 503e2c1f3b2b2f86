"""A/B of the fused kernel with independent CTAs (VSB_TC_CL=1) and with CTA pairs that share the streamed base tiles
through TMA multicast (VSB_TC_CL=2): kernel time (CUDA events on the launching stream, min of 3 after a warm-up) per
precision path, base size and VSB_TC_DBG flag (0 = real kernel, 2 = no candidate hand-off, 1 = no epilogue work).
Usage: python tools/tc_cluster_ab.py [precs=f16,1x,3x] [sizes=1000000,125000] [dbg=0,1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vsb200_loader

vsb = vsb200_loader.load()
precs = (sys.argv[1] if len(sys.argv) > 1 else "f16,1x,3x").split(",")
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1000000,125000").split(",")]
flags = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,1").split(",")]
nq, k = 10_000, 10
dev = torch.device("cuda:0")
PREC = {"f16": vsb.PREC_F16_CERT, "1x": vsb.PREC_TF32_1X, "3x": vsb.PREC_3XTF32}
q = torch.from_numpy(vsb.synth.make("cont", 2026, nq)).to(dev)
ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
d = torch.empty((nq, k), dtype=torch.float32, device=dev)
st = torch.cuda.Stream()
for n in sizes:
    base = torch.empty((n, 128), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base.data_ptr(), 0, n, 128, "cont", 2025)
    torch.cuda.synchronize()
    idx = vsb.ExactIndex(base.data_ptr(), n=n)
    idx.set_profile(True)
    ref = {}
    for name in precs:
        for dbg in flags:
            for cl in (1, 2):
                os.environ["VSB_TC_DBG"], os.environ["VSB_TC_CL"] = str(dbg), str(cl)
                ts = []
                for it in range(4):
                    try:
                        idx.search_dev(q.data_ptr(), nq, k, PREC[name], ids.data_ptr(), d.data_ptr(), st.cuda_stream)
                    except Exception as e:
                        if dbg == 0:
                            raise
                    st.synchronize()
                    ts.append(idx.last_kernel_ms())
                same = ""
                if dbg == 0:
                    got = (ids.cpu().numpy().copy(), d.cpu().numpy().copy())
                    if cl == 1:
                        ref[name] = got
                    else:
                        same = "  results == CL1: %s" % (np.array_equal(got[0], ref[name][0]) and np.array_equal(got[1], ref[name][1]))
                print(f"N={n} {name:>3s} dbg={dbg} CL={cl}  kernel ms: {np.min(ts[1:]):8.3f}{same}", flush=True)
    idx.close()
    del base
os.environ["VSB_TC_DBG"] = "0"
