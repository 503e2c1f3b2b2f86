"""CPU tests (no GPU): the oracle restatement against the golden vectors produced by the UNMODIFIED reference
(tests/golden/gen_golden.py ran oracle/_ref/ref_driver = /root/reference/cpu/cpu_baseline.cpp)."""
import numpy as np
import pytest

from util import assert_topk_matches, golden_cases, load_golden


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_reference_golden(path, vsb, oracle):
    g, law, base, qry, k = load_golden(path, vsb.synth)
    exact = law == "sift"  # integer-valued data: every fp32 summation order is exact
    # norms: the restatement follows the reference's summation order => bit-exact on any data
    assert np.array_equal(oracle.norms(base), g["bnorms"])
    assert np.array_equal(oracle.norms(qry), g["qnorms"])
    for mode in (0, 1):  # literal select_topk ties / canonical order
        ids, d = oracle.exact_search(base, qry, k, mode=mode)
        rec = oracle.exact_distances_at(base, qry, ids)
        assert_topk_matches(ids, d, g["ids"], g["dists"], rec, exact=exact, what=f"{path} mode={mode}")
    # the reference's own text output agrees with its functions to the 6 digits it prints
    assert np.allclose(g["text_dists"], g["dists"], rtol=6e-6)


def test_literal_topk_equals_reference_on_ties(vsb, oracle):
    """select_topk's slot-replacement tie behaviour (cpu_baseline.cpp:140-150) is restated literally: on the
    tie-heavy integer fixtures mode 0 must reproduce the reference's ids exactly, not only tie-equivalently."""
    for path in golden_cases():
        g, law, base, qry, k = load_golden(path, vsb.synth)
        if law != "sift":
            continue
        ids, d = oracle.exact_search(base, qry, k, mode=0)
        assert np.array_equal(d, g["dists"])
        # std::sort among equal distances is unspecified: compare as sets per distance value
        for r in range(ids.shape[0]):
            for v in np.unique(d[r]):
                assert set(ids[r][d[r] == v]) == set(g["ids"][r][g["dists"][r] == v]), (path, r, v)


def test_synth_laws(vsb):
    s = vsb.synth
    a = s.make("sift", 5, 1000)
    assert a.dtype == np.float32 and a.min() >= 0 and a.max() <= 127 and np.array_equal(a, np.round(a))
    assert np.array_equal(s.rows("sift", 5, 100, 50), a[100:150])
    c = s.make("cont", 5, 1000)
    assert np.abs(c - a).max() <= 0.5 and not np.array_equal(c, np.round(c))
    m = s.make("mix", 5, 1000)
    assert m.min() >= 0 and m.max() <= 218 and np.array_equal(m, np.round(m))


def test_fvecs_roundtrip(tmp_path, vsb):
    s = vsb.synth
    a = s.make("cont", 9, 33)
    p = str(tmp_path / "a.fvecs")
    s.write_fvecs(p, a)
    assert np.array_equal(s.read_fvecs(p), a)
    with open(p, "ab") as f:
        f.write(b"\x80\x00\x00\x00\x00\x00")
    with pytest.raises(ValueError):
        s.read_fvecs(p)


def test_int8_quantiser_rule(oracle):
    """QnnRunner.cpp:13-55: trunc(x * (1/scale) + 0.5) saturated to [0, 255]."""
    x = np.array([0.0, 0.33, 0.34, 1.0, 100.0, 168.9, 169.0, 500.0, -3.0, 0.6627451 * 7.5], dtype=np.float32)
    q = oracle.quantize_u8(x, 0.6627451)
    inv = np.float32(1.0) / np.float32(0.6627451)
    want = np.clip(np.trunc((x * inv).astype(np.float32) + np.float32(0.5)), 0, 255).astype(np.uint8)
    assert np.array_equal(q, want)
    assert q[-2] == 0 and q[7] == 255


def test_ivf_restatement_selfconsistent(vsb, oracle):
    s = vsb.synth
    base = s.make("mix", 3, 4000)
    qry = s.make("mix", 4, 20)
    rng = np.random.default_rng(0)
    cent = base[rng.choice(4000, 32, replace=False)]
    lab, _ = oracle.kmeans_assign(base, cent)
    order = np.argsort(lab, kind="stable").astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(lab, minlength=32))]).astype(np.int32)
    coarse = oracle.ivf_coarse(qry, cent)
    # reordered layout and scattered layout give identical answers
    a = oracle.ivf_search(base[order], offsets, order, True, coarse, qry, 10, 4, mode=1)
    b = oracle.ivf_search(base, offsets, order, False, coarse, qry, 10, 4, mode=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[3] == b[3]
    # nprobe == nlist degenerates to exact maximum inner product
    full = oracle.ivf_search(base[order], offsets, order, True, coarse, qry, 10, 32, mode=1)
    ip = qry.astype(np.float64) @ base.astype(np.float64).T
    want = np.argsort(-ip, axis=1, kind="stable")[:, :10]
    assert np.array_equal(np.sort(full[0], 1), np.sort(want, 1))
    assert full[3] == 20 * 4000


@pytest.mark.parametrize("path", golden_cases("hp2_"), ids=lambda p: p.split("/")[-1][:-4])
def test_ivf_restatement_matches_reference_golden(path, vsb, oracle):
    """HP2 pin: the C restatement of IVFIndex::searchBatch (oracle/vs_oracle.c) against what the reference's own
    IVFSearcher (qidk_ivf/prepare/benchmark_ivf.py:53-140, run unmodified by tests/golden/gen_golden_ivf.py) returned
    on the same scattered-layout index: scores bit-exact, candidate counts equal, ids equal outside ties, recall equal."""
    from util import assert_ivf_matches_golden, load_golden_ivf

    g, base, qry, cent, labels, offsets, indices = load_golden_ivf(path, vsb.synth)
    k = int(g["k"])
    # the directory the reference consumed was assigned by exact nearest centroid: the restatement agrees
    olab, _ = oracle.kmeans_assign(base, cent)
    assert np.array_equal(olab, labels)
    coarse = oracle.ivf_coarse(qry, cent)
    assert np.array_equal(coarse, (qry.astype(np.float64) @ cent.astype(np.float64).T).astype(np.float32))
    for nprobe in g["nprobes"]:
        for reordered, vec in ((False, base), (True, base[indices])):
            for mode in (0, 1):
                ids, sc, cnt, total = oracle.ivf_search(vec, offsets, indices, reordered, coarse, qry, k, int(nprobe), mode=mode)
                rec = oracle.ivf_scores_at(base, qry, ids)
                assert_ivf_matches_golden(g, int(nprobe), ids, sc, cnt, total, rec, g["gt"],
                                          what=f"{path} nprobe={nprobe} reordered={reordered} mode={mode}")
