"""GPU parity tests for the IVF path (coarse scores, probe selection, list scan, builder, on-disk format) against
the CPU restatement of IVFIndex::searchBatch (oracle/vs_oracle.c) on the same index arrays."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_index(vsb, oracle, n, nlist, seed=3, law="mix"):
    base = vsb.synth.make(law, seed, n)
    rng = np.random.default_rng(seed)
    cent = base[rng.choice(n, nlist, replace=False)].copy()
    lab, _ = oracle.kmeans_assign(base, cent)
    order = np.argsort(lab, kind="stable").astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(lab, minlength=nlist))]).astype(np.int32)
    return base, cent, order, offsets


def _set_scan(monkeypatch, scan):
    """query_major = K6; list_major = K8 (FFMA, scores bit-identical by construction); list_major_tc = the tensor-core
    list-major scan (TF32 candidates + re-score in the reference's summation order)."""
    monkeypatch.setenv("VSB_IVF_LM", "0" if scan == "query_major" else "1")
    monkeypatch.setenv("VSB_IVF_TC", "1" if scan == "list_major_tc" else "0")


@pytest.mark.parametrize("scan", ["query_major", "list_major", "list_major_tc"])
@pytest.mark.parametrize("law", ["mix", "cont"])
@pytest.mark.parametrize("n,nlist,nq,k,nprobe", [(20000, 64, 50, 10, 8), (5000, 16, 33, 5, 16), (3000, 300, 7, 10, 3),
                                                 (100000, 256, 200, 10, 32), (2000, 8, 5, 32, 100), (30000, 40, 700, 16, 5),
                                                 (120000, 64, 400, 30, 32),   # up to 3 x 32 x 32 candidates per query (K9)
                                                 (9000, 700, 65, 1, 9)])
def test_search_matches_restatement(scan, law, n, nlist, nq, k, nprobe, gpu_vsb, oracle, monkeypatch):
    """Both fine-scan kernels (K6 query-major, K8 list-major) against the CPU restatement: same probe sets, bit-identical
    scores, canonical ids, counts and total candidates."""
    _set_scan(monkeypatch, scan)
    vsb = gpu_vsb
    base, cent, order, offsets = make_index(vsb, oracle, n, nlist, law=law)
    qry = vsb.synth.make(law, 99, nq)
    idx = vsb.IvfIndex(vectors=base[order], offsets=offsets, id_map=order, centroids=cent)
    try:
        assert (idx.num_vectors, idx.num_clusters, idx.dim) == (n, nlist, 128)
        coarse = idx.coarse_scores(qry)
        ref_coarse = oracle.ivf_coarse(qry, cent)
        assert np.array_equal(coarse, ref_coarse)  # same fp32 summation order => bit-exact on any data
        ids, sc, cnt, total = idx.search_batch(qry, k, nprobe)
    finally:
        idx.close()
    rids, rsc, rcnt, rtotal = oracle.ivf_search(base[order], offsets, order, True, ref_coarse, qry, k, nprobe, mode=1)
    assert total == rtotal
    assert np.array_equal(cnt, rcnt)
    assert np.array_equal(sc, rsc)      # bit-exact scores
    assert np.array_equal(ids, rids)    # canonical (score desc, id asc) order
    # recall@k equals the CPU IVF reference at the same nprobe and centroids (BASELINE.json north_star)
    # and the literal heap variant agrees up to tie order
    lids, lsc, _, _ = oracle.ivf_search(base[order], offsets, order, True, ref_coarse, qry, k, nprobe, mode=0)
    assert np.array_equal(lsc, sc)


def test_recall_equals_cpu_reference(gpu_vsb, oracle):
    """recall@10 against exact-L2 ground truth (the reference's metric mix, main_ivf.cpp:52-59) is identical for
    the GPU path and the CPU restatement."""
    vsb = gpu_vsb
    n, nlist, nq, k = 200_000, 256, 500, 10
    base, cent, order, offsets = make_index(vsb, oracle, n, nlist)
    qry = vsb.synth.make("mix", 123, nq)
    gt, _ = oracle.exact_search(base, qry, k, mode=1)
    idx = vsb.IvfIndex(vectors=base[order], offsets=offsets, id_map=order, centroids=cent)
    coarse = oracle.ivf_coarse(qry, cent)
    try:
        for nprobe in (8, 32):
            ids, _, _, _ = idx.search_batch(qry, k, nprobe)
            rids, _, _, _ = oracle.ivf_search(base[order], offsets, order, True, coarse, qry, k, nprobe, mode=1)
            rec = lambda a: np.mean([len(set(a[i]) & set(gt[i])) / k for i in range(nq)])
            assert rec(ids) == rec(rids)
            assert 0.05 < rec(ids) <= 1.0
    finally:
        idx.close()


@pytest.mark.parametrize("reordered", [True, False])
def test_build_and_directory_format(reordered, tmp_path, gpu_vsb, oracle):
    vsb = gpu_vsb
    n, nlist = 30_000, 64
    base = vsb.synth.make("mix", 11, n)
    d = str(tmp_path / "idx")
    info = vsb.ivf_build(base, nlist, d, max_iter=20, seed=42, reordered=reordered)
    assert info["nlist"] == nlist and 1 <= info["iters"] <= 20
    cfg = json.load(open(os.path.join(d, "ivf_config.json")))
    assert cfg["n_vectors"] == n and cfg["n_clusters"] == nlist and cfg["dim"] == 128
    assert bool(cfg.get("reordered", False)) == reordered
    off = np.load(os.path.join(d, "cluster_offsets.npy"))
    cent = np.load(os.path.join(d, "centroids.npy"))
    lab = np.load(os.path.join(d, "cluster_ids.npy"))
    assert off.dtype == np.int32 and off.shape == (nlist + 1,) and off[0] == 0 and off[-1] == n
    assert cent.dtype == np.float32 and cent.shape == (nlist, 128) and lab.dtype == np.int32
    # labels are the exact arg-min assignment of the final centroids (ties -> lowest id)
    olab, od = oracle.kmeans_assign(base, cent)
    assert np.array_equal(lab, olab)
    assert abs(info["inertia"] - float(od.astype(np.float64).sum())) <= 1e-6 * info["inertia"]
    # k-means made progress: inertia below that of the initial random sample of rows
    rng = np.random.default_rng(0)
    _, d0 = oracle.kmeans_assign(base, base[rng.choice(n, nlist, replace=False)])
    assert info["inertia"] < float(d0.astype(np.float64).sum())
    if reordered:
        perm = np.load(os.path.join(d, "reorder_to_original.npy"))
        vec = np.load(os.path.join(d, "vectors_reordered.npy"))
        assert np.array_equal(vec, base[perm])
        assert np.array_equal(np.load(os.path.join(d, "cluster_sizes.npy")), np.diff(off))
    else:
        perm = np.load(os.path.join(d, "cluster_indices.npy"))
        assert np.array_equal(np.load(os.path.join(d, "vectors.npy")), base)
    assert np.array_equal(np.sort(perm), np.arange(n))
    for c in range(nlist):  # lists hold ascending original ids (np.where order, create_ivf_model.py:117)
        seg = perm[off[c]:off[c + 1]]
        assert (lab[seg] == c).all() and (np.diff(seg) > 0).all()
    # the directory loads back and searches like the arrays it was built from
    qry = vsb.synth.make("mix", 12, 40)
    idx = vsb.IvfIndex(d)
    try:
        ids, sc, cnt, total = idx.search_batch(qry, 10, 8)
    finally:
        idx.close()
    coarse = oracle.ivf_coarse(qry, cent)
    rids, rsc, rcnt, rtotal = oracle.ivf_search(base[perm], off, perm, True, coarse, qry, 10, 8, mode=1)
    assert np.array_equal(ids, rids) and np.array_equal(sc, rsc) and total == rtotal


def test_parity_mode_fixed_centroids_and_numpy_written_dir(tmp_path, gpu_vsb, oracle):
    """An index directory written with numpy exactly as the reference builders do (np.save / json.dump) loads,
    and build with init_centroids + max_iter=0 keeps the centroids unchanged."""
    vsb = gpu_vsb
    n, nlist = 8000, 32
    base, cent, order, offsets = make_index(vsb, oracle, n, nlist)
    d = tmp_path / "npdir"
    d.mkdir()
    json.dump({"n_vectors": n, "n_clusters": nlist, "dim": 128, "batch_size": 1, "avg_cluster_size": n / nlist,
               "min_cluster_size": 1, "max_cluster_size": n}, open(d / "ivf_config.json", "w"), indent=2)
    np.save(d / "cluster_offsets.npy", offsets)
    np.save(d / "cluster_indices.npy", order)
    np.save(d / "vectors.npy", base)
    np.save(d / "centroids.npy", cent)
    qry = vsb.synth.make("mix", 5, 25)
    idx = vsb.IvfIndex(str(d))
    try:
        ids, sc, _, _ = idx.search_batch(qry, 10, 4)
    finally:
        idx.close()
    rids, rsc, _, _ = oracle.ivf_search(base, offsets, order, False, oracle.ivf_coarse(qry, cent), qry, 10, 4, mode=1)
    assert np.array_equal(ids, rids) and np.array_equal(sc, rsc)
    d2 = str(tmp_path / "fixed")
    info = vsb.ivf_build(base, nlist, d2, max_iter=0, init_centroids=cent)
    assert info["iters"] == 0 and np.array_equal(np.load(os.path.join(d2, "centroids.npy")), cent)
    assert np.array_equal(np.load(os.path.join(d2, "cluster_offsets.npy")), offsets)
    with pytest.raises(vsb.VsbError):
        vsb.IvfIndex(str(tmp_path / "missing"))  # ctor throws on missing files, like IVFIndex.cpp:184-197


def test_full_size_list_major_equals_query_major(tmp_path, gpu_vsb, monkeypatch):
    """BASELINE configs[2] at full size (1M x 128, nlist 1024, 10 000 queries, nprobe 8 / 32, top-10): the list-major
    kernel (K8: pairs grouped by list, cross-list bounds, 296 CTAs pulling work items) must return exactly what the
    query-major kernel (K6) returns — same ids, bit-identical scores, same counts and total — and do so repeatably."""
    vsb = gpu_vsb
    n, nlist, nq, k = 1_000_000, 1024, 10_000, 10
    base = vsb.synth.make("mix", 2025, n)
    d_ = str(tmp_path / "idx")
    os.makedirs(d_)
    vsb.ivf_build(base, nlist, d_, max_iter=4, seed=42, reordered=True)
    qry = vsb.synth.make("mix", 2026, nq)
    idx = vsb.IvfIndex(d_)
    try:
        for nprobe in (8, 32):
            _set_scan(monkeypatch, "query_major")
            want = idx.search_batch(qry, k, nprobe)
            for scan in ("list_major", "list_major_tc"):
                _set_scan(monkeypatch, scan)
                for _ in range(3):
                    got = idx.search_batch(qry, k, nprobe)
                    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (scan, nprobe)
                    assert np.array_equal(got[2], want[2]) and got[3] == want[3]
    finally:
        idx.close()


def _hp2_cases():
    from util import golden_cases

    return golden_cases("hp2_")


@pytest.mark.parametrize("scan", ["query_major", "list_major", "list_major_tc"])
@pytest.mark.parametrize("path", _hp2_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_gpu_matches_reference_ivfsearcher_golden(path, scan, tmp_path, gpu_vsb, oracle, monkeypatch):
    """HP2 pin on the GPU: tests/golden/hp2_*.npz hold what the reference's OWN searcher (IVFSearcher.search,
    qidk_ivf/prepare/benchmark_ivf.py:96-140, imported unmodified) returned on a scattered-layout directory.  The same
    directory is written here by vs_ivf_build (parity mode: fixed centroids), must equal the one the reference consumed
    array by array, and searching it must give the reference's scores bit for bit, its candidate counts, its ids outside
    ties and its recall@k."""
    from util import assert_ivf_matches_golden, load_golden_ivf

    _set_scan(monkeypatch, scan)
    vsb = gpu_vsb
    g, base, qry, cent, labels, offsets, indices = load_golden_ivf(path, vsb.synth)
    k, nlist = int(g["k"]), int(g["nlist"])
    d = str(tmp_path / "idx")
    info = vsb.ivf_build(base, nlist, d, max_iter=0, init_centroids=cent, reordered=False)
    assert info["nlist"] == nlist and info["iters"] == 0
    assert np.array_equal(np.load(os.path.join(d, "centroids.npy")), cent)
    assert np.array_equal(np.load(os.path.join(d, "cluster_ids.npy")), labels)
    assert np.array_equal(np.load(os.path.join(d, "cluster_offsets.npy")), offsets)
    assert np.array_equal(np.load(os.path.join(d, "cluster_indices.npy")), indices)
    assert np.array_equal(np.load(os.path.join(d, "vectors.npy")), base)
    for how in ("dir", "arrays"):
        idx = vsb.IvfIndex(d) if how == "dir" else vsb.IvfIndex(vectors=base[indices], offsets=offsets, id_map=indices,
                                                                  centroids=cent)
        try:
            for nprobe in g["nprobes"]:
                ids, sc, cnt, total = idx.search_batch(qry, k, int(nprobe))
                rec = oracle.ivf_scores_at(base, qry, ids)
                assert_ivf_matches_golden(g, int(nprobe), ids, sc, cnt, total, rec, g["gt"],
                                          what=f"{path} {how} nprobe={nprobe} {scan}")
        finally:
            idx.close()


@pytest.mark.parametrize("law,n,nlist", [("mix", 60_000, 128), ("cont", 40_000, 64)])
def test_builder_matches_sklearn_kmeans(law, n, nlist, tmp_path, gpu_vsb, oracle):
    """build_ivf_index (qidk_ivf/prepare/create_ivf_model.py:102-110) is KMeans(n_clusters, random_state=42, n_init=1,
    max_iter=100): greedy k-means++ seeding + Lloyd with sklearn's tolerance.  The random streams differ, so the centroids
    differ; what must agree is the QUALITY of the clustering — final inertia within 2 % of sklearn's on the same data —
    and the build must be a pure function of (data, seed)."""
    from sklearn.cluster import KMeans

    vsb = gpu_vsb
    base = vsb.synth.make(law, 21, n)
    d1, d2 = str(tmp_path / "a"), str(tmp_path / "b")
    info = vsb.ivf_build(base, nlist, d1, max_iter=100, seed=42)
    km = KMeans(n_clusters=nlist, random_state=42, n_init=1, max_iter=100).fit(base)
    ratio = info["inertia"] / float(km.inertia_)
    print(f"\n[kmeans] {law} n={n} k={nlist}: inertia {info['inertia']:.6g} after {info['iters']} iterations, "
          f"sklearn {km.inertia_:.6g} after {km.n_iter_}: ratio {ratio:.4f}")
    assert 0.97 <= ratio <= 1.02
    # the reported inertia is the true one of the written centroids / labels
    cent = np.load(os.path.join(d1, "centroids.npy"))
    lab = np.load(os.path.join(d1, "cluster_ids.npy"))
    olab, od = oracle.kmeans_assign(base, cent)
    assert np.array_equal(lab, olab)
    assert abs(info["inertia"] - float(od.astype(np.float64).sum())) <= 1e-6 * info["inertia"]
    # k-means++ beats a uniform row sample as a starting point (same Lloyd, same data)
    rng = np.random.default_rng(0)
    info_u = vsb.ivf_build(base, nlist, str(tmp_path / "u"), max_iter=0, init_centroids=base[rng.choice(n, nlist, replace=False)])
    info_p = vsb.ivf_build(base, nlist, str(tmp_path / "p"), max_iter=0, seed=42)
    assert info_p["inertia"] < info_u["inertia"]
    # deterministic
    info2 = vsb.ivf_build(base, nlist, d2, max_iter=100, seed=42)
    assert info2 == info and np.array_equal(np.load(os.path.join(d2, "centroids.npy")), cent)
    assert np.array_equal(np.load(os.path.join(d2, "cluster_indices.npy" if not os.path.exists(os.path.join(d2, "reorder_to_original.npy")) else "reorder_to_original.npy")),
                          np.load(os.path.join(d1, "cluster_indices.npy" if not os.path.exists(os.path.join(d1, "reorder_to_original.npy")) else "reorder_to_original.npy")))
