"""Generates tests/golden/hp2_*.npz by RUNNING THE UNMODIFIED REFERENCE IVF searcher — class IVFSearcher of
/root/reference/qidk_ivf/prepare/benchmark_ivf.py:53-140, the Python twin of IVFIndex::search — on index
directories written in the reference's on-disk format (scattered layout).

    python tests/golden/gen_golden_ivf.py            (build container only: needs /root/reference)

benchmark_ivf.py imports onnxruntime (absent here) for ONE call: centroid_session.run(None, {"query": q}), whose
graph is the single MatMul  query[1,dim] x centroids^T[dim,nlist]  (create_ivf_model.py:121-131).  A stub module
with exactly that arithmetic is placed in sys.modules before the import; everything else — the list gather, the
per-list dot products, argpartition / argsort top-k, the candidate count — is the reference's own code, unmodified.

All data is integer valued (law "mix", integer centroids), so every fp32 summation order gives the same bits
(SURVEY.md Appendix D): ids / scores / candidate counts pin bit-exactly, ties aside.  The generator checks that no
query has a coarse-score tie at the nprobe boundary (argpartition's choice there would be arbitrary).

Each fixture stores the generator parameters (inputs are regenerated from seeds), the centroids, the labels the
directory was built from, and per nprobe what the reference returned: ids, scores, candidates, recall@k against the
exact-L2 ground truth (HP1 reference, compute_recall of benchmark_ivf.py:166-170).
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/qidk_ivf/prepare/benchmark_ivf.py"

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "hai-25-rag-on-edge_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)


def load_reference():
    """import benchmark_ivf.py unmodified behind a stub onnxruntime (InferenceSession.run = query @ centroids^T)."""
    ort = types.ModuleType("onnxruntime")

    class InferenceSession:
        def __init__(self, path, providers=None):
            # centroids.onnx holds centroids^T as its only initializer; the builder saves the same numbers next to it
            self.cent = np.load(os.path.join(os.path.dirname(path), "centroids.npy")).astype(np.float32)

        def run(self, outputs, feeds):
            q = np.asarray(feeds["query"], dtype=np.float32)
            return [q @ self.cent.T]

    ort.InferenceSession = InferenceSession
    sys.modules["onnxruntime"] = ort
    s = importlib.util.spec_from_file_location("benchmark_ivf_ref", REF)
    m = importlib.util.module_from_spec(s)
    s.loader.exec_module(m)
    return m


def write_scattered_dir(d, base, cent, labels):
    """create_ivf_model.py:112-166: lists = np.where(cluster_ids == i) (ascending ids), CSR offsets, int32 dtypes."""
    n, nlist = base.shape[0], cent.shape[0]
    lists = [np.where(labels == i)[0].astype(np.int32) for i in range(nlist)]
    sizes = np.array([len(x) for x in lists])
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    indices = np.concatenate(lists).astype(np.int32)
    json.dump({"n_vectors": int(n), "n_clusters": int(nlist), "dim": int(base.shape[1]), "batch_size": 1,
               "avg_cluster_size": float(sizes.mean()), "min_cluster_size": int(sizes.min()),
               "max_cluster_size": int(sizes.max())}, open(os.path.join(d, "ivf_config.json"), "w"), indent=2)
    np.save(os.path.join(d, "cluster_ids.npy"), labels.astype(np.int32))
    np.save(os.path.join(d, "cluster_offsets.npy"), offsets)
    np.save(os.path.join(d, "cluster_indices.npy"), indices)
    np.save(os.path.join(d, "vectors.npy"), base.astype(np.float32))
    np.save(os.path.join(d, "centroids.npy"), cent.astype(np.float32))
    open(os.path.join(d, "centroids.onnx"), "wb").close()  # path only; the stub reads centroids.npy
    return offsets, indices


def assign_l2(base, cent):
    """exact nearest centroid (integer data: float64 is exact), ties -> lowest cluster id"""
    b = base.astype(np.float64)
    c = cent.astype(np.float64)
    lab = np.empty(b.shape[0], dtype=np.int32)
    cn = (c * c).sum(1)
    for r0 in range(0, b.shape[0], 8192):
        d = cn[None, :] - 2.0 * (b[r0:r0 + 8192] @ c.T)
        lab[r0:r0 + 8192] = np.argmin(d, axis=1)
    return lab


CASES = [
    # name,           n,      nlist, nq, k,  nprobes,    base_seed, query_seed, centroid rule
    ("hp2_small",     20000,  64,    60, 10, (8, 32),    3,         99,         "rows"),
    ("hp2_short",     3000,   300,   25, 10, (3, 8),     5,         98,         "rows"),    # lists shorter than k exist
    ("hp2_k5_all",    5000,   16,    33, 5,  (8, 100),   7,         97,         "rows"),    # nprobe > nlist -> all lists
    ("hp2_kmeans",    60000,  256,   80, 10, (8, 32),    2025,      2026,       "kmeans"),  # rounded Lloyd centroids
]


def main():
    ref = load_reference()
    from oracle import oracle
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, n, nlist, nq, k, nprobes, bs, qs, rule in CASES:
        base = synth.make("mix", bs, n)
        qry = synth.make("mix", qs, nq)
        rng = np.random.default_rng(bs)
        cent = base[np.sort(rng.choice(n, nlist, replace=False))].copy()
        if rule == "kmeans":  # a few Lloyd steps, then rounded to integers (keeps every summation order exact)
            for _ in range(5):
                lab = assign_l2(base, cent)
                for c in range(nlist):
                    if (lab == c).any():
                        cent[c] = base[lab == c].mean(0)
            cent = np.clip(np.rint(cent), 0, 218).astype(np.float32)
        labels = assign_l2(base, cent)
        gt, _ = oracle.exact_search(base, qry, k, mode=1)  # exact-L2 ground truth (HP1 oracle, pinned by hp1_*.npz)
        with tempfile.TemporaryDirectory() as td:
            write_scattered_dir(td, base, cent, labels)
            searcher = ref.IVFSearcher(td)
            coarse = qry.astype(np.float64) @ cent.astype(np.float64).T
            out = {}
            for nprobe in nprobes:
                if nprobe < nlist:  # no tie at the selection boundary
                    srt = -np.sort(-coarse, axis=1)
                    assert (srt[:, nprobe - 1] > srt[:, nprobe]).all(), (name, nprobe, "coarse tie at the boundary")
                ids = np.full((nq, k), -1, dtype=np.int32)
                sc = np.full((nq, k), -np.inf, dtype=np.float32)
                cand = np.zeros(nq, dtype=np.int64)
                rec = np.zeros(nq, dtype=np.float64)
                for i in range(nq):
                    a, s, c = searcher.search(qry[i], k=k, nprobe=nprobe)
                    ids[i, :len(a)] = a
                    sc[i, :len(s)] = s
                    cand[i] = c
                    rec[i] = ref.compute_recall(a, gt[i], k=k)
                out[f"ids_np{nprobe}"] = ids
                out[f"scores_np{nprobe}"] = sc
                out[f"cand_np{nprobe}"] = cand
                out[f"recall_np{nprobe}"] = rec
                print(f"{name}: nprobe={nprobe} recall@{k}={rec.mean():.4f} candidates/query={cand.mean():.0f}")
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), law="mix", n=n, nlist=nlist, nq=nq, k=k,
                            nprobes=np.array(nprobes), base_seed=bs, query_seed=qs, centroids=cent,
                            labels=labels.astype(np.int16 if nlist < 32768 else np.int32), gt=gt, **out)


if __name__ == "__main__":
    main()
