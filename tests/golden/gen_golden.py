"""Generates tests/golden/hp1_*.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/ref_driver, built by
oracle/Makefile from /root/reference/cpu/cpu_baseline.cpp) on seeded synthetic inputs.  Run in the build
container (the reference sources are not on the GPU box):

    python tests/golden/gen_golden.py

Each fixture stores the generator parameters (law, seeds, shapes, k) — inputs are regenerated from them —
plus what the reference produced: ids/dists from its functions (binary, exact float32), the ids/dists parsed
from run_benchmark()'s own results file (6 significant digits), and the norms.
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "hai-25-rag-on-edge_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

CASES = [
    # name,            law,   nb,    nq,  k, base_seed, query_seed
    ("hp1_cfg0_sift", "sift", 10000, 100, 5, 1234, 4321),   # BASELINE.json configs[0]
    ("hp1_cfg0_cont", "cont", 10000, 100, 5, 1234, 4321),   # same shape, continuous values (3xTF32 check)
    ("hp1_k10_sift", "sift", 4096, 64, 10, 11, 12),
    ("hp1_k10_cont", "cont", 5000, 37, 10, 21, 22),         # ragged: nb, nq not multiples of any tile
    ("hp1_k1_cont", "cont", 777, 5, 1, 31, 32),
    ("hp1_k32_sift", "sift", 3000, 33, 32, 41, 42),
    ("hp1_k100_sift", "sift", 2500, 9, 100, 51, 52),
    ("hp1_k_eq_n", "cont", 16, 3, 16, 61, 62),              # k == N (k > N is UB in the reference)
]


def main():
    oracle.build(ref=True)
    assert oracle.have_ref(), "oracle/_ref/ref_driver missing (needs /root/reference)"
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, law, nb, nq, k, bs, qs in CASES:
        base = synth.make(law, bs, nb)
        qry = synth.make(law, qs, nq)
        with tempfile.TemporaryDirectory() as td:
            bf, qf, rt = (os.path.join(td, x) for x in ("base.fvecs", "query.fvecs", "results.txt"))
            synth.write_fvecs(bf, base)
            synth.write_fvecs(qf, qry)
            ids, dists, qn, bn = oracle.ref_dump(bf, qf, k)
            oracle.ref_bench(bf, qf, k, rt)
            tids, tdists = oracle.parse_results_txt(rt)
        # the reference's text output and its functions must agree (ids exactly, dists to 6 digits)
        assert np.array_equal(np.sort(ids, 1), np.sort(tids, 1)) or law == "sift", name
        assert np.allclose(dists, tdists, rtol=6e-6), name
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), law=law, nb=nb, nq=nq, k=k, base_seed=bs,
                            query_seed=qs, ids=ids, dists=dists, text_ids=tids, text_dists=tdists, qnorms=qn,
                            bnorms=bn)
        print(f"{name}: nb={nb} nq={nq} k={k} ok")


if __name__ == "__main__":
    main()
