"""First-contact diagnostics for the CUDA kernels: runs small cases stage by stage, each in a subprocess with a
timeout (a hung kernel must not take the box down), and prints mismatch statistics instead of pass/fail."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["synth", "ffma", "tc1x_tile", "tc1x", "tc3x", "tc3x_big", "k100"]


def stage(name):
    import numpy as np

    import vsb200_loader
    from oracle import oracle

    vsb = vsb200_loader.load()

    def cmp(tag, base, qry, k, prec):
        idx = vsb.ExactIndex(base)
        ids, d = idx.search(qry, k, prec)
        idx.close()
        oi, od = oracle.exact_search(base, qry, k, mode=1)
        rel = np.abs(d - od) / np.maximum(np.abs(od), 1e-30)
        print(f"[{tag}] ids equal {np.mean(ids == oi):.4f}  max rel dist err {rel.max():.3e}  "
              f"rows all-equal {np.mean((ids == oi).all(1)):.4f}")
        if not (ids == oi).all():
            r = int(np.argwhere((ids != oi).any(1))[0][0])
            print("  first bad row", r, "\n  gpu ids", ids[r], "\n  ref ids", oi[r], "\n  gpu d", d[r], "\n  ref d", od[r])

    if name == "synth":
        import torch
        t = torch.empty((1000, 128), dtype=torch.float32, device="cuda")
        for law in ("sift", "cont", "mix"):
            vsb.synth_fill_dev(t.data_ptr(), 37, 1000, 128, law, 99)
            torch.cuda.synchronize()
            print("[synth]", law, "equal:", np.array_equal(t.cpu().numpy(), vsb.synth.rows(law, 99, 37, 1000)))
    elif name == "ffma":
        base = vsb.synth.make("cont", 1, 5000)
        for nq in (1, 2, 3, 8, 11):
            cmp(f"ffma nq={nq}", base, vsb.synth.make("cont", 2, nq), 10, vsb.PREC_FFMA)
    elif name == "tc1x_tile":
        cmp("tc1x 128x128", vsb.synth.make("sift", 1, 128), vsb.synth.make("sift", 2, 128), 10, vsb.PREC_TF32_1X)
    elif name == "tc1x":
        cmp("tc1x 5000x300", vsb.synth.make("sift", 1, 5000), vsb.synth.make("sift", 2, 300), 10, vsb.PREC_TF32_1X)
    elif name == "tc3x":
        cmp("tc3x 5000x300 cont", vsb.synth.make("cont", 1, 5000), vsb.synth.make("cont", 2, 300), 10, vsb.PREC_3XTF32)
        cmp("tc3x 5000x300 sift", vsb.synth.make("sift", 1, 5000), vsb.synth.make("sift", 2, 300), 10, vsb.PREC_3XTF32)
    elif name == "tc3x_big":
        cmp("tc3x 200000x1000", vsb.synth.make("cont", 1, 200000), vsb.synth.make("cont", 2, 1000), 10, vsb.PREC_3XTF32)
    elif name == "k100":
        cmp("k100 tc", vsb.synth.make("cont", 1, 20000), vsb.synth.make("cont", 2, 40), 100, vsb.PREC_3XTF32)
        cmp("k100 ffma", vsb.synth.make("cont", 1, 20000), vsb.synth.make("cont", 2, 5), 100, vsb.PREC_FFMA)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        stage(sys.argv[1])
        sys.exit(0)
    for s in STAGES:
        try:
            r = subprocess.run([sys.executable, __file__, s], timeout=120, capture_output=True, text=True)
            print(r.stdout.strip())
            if r.returncode != 0:
                print(f"[{s}] FAILED rc={r.returncode}\n{r.stderr[-2000:]}")
        except subprocess.TimeoutExpired:
            print(f"[{s}] TIMEOUT (kernel hang?)")
            break
