"""Helpers shared by the parity tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5  # BASELINE.json north_star: fp32 distances within 1e-5 relative; ids bit-exact except ties within 1e-5


def golden_cases(prefix="hp1_"):
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(path, synth):
    g = np.load(path)
    law = str(g["law"])
    base = synth.make(law, int(g["base_seed"]), int(g["nb"]))
    qry = synth.make(law, int(g["query_seed"]), int(g["nq"]))
    return g, law, base, qry, int(g["k"])


def assert_topk_matches(ids, keys, ref_ids, ref_keys, recomputed, *, exact: bool, what=""):
    """Tie-aware comparison.
      1. key vectors equal position by position (bit-exact when `exact`, else within RTOL);
      2. every returned id's key, recomputed by the oracle, equals the reported key;
      3. ids identical wherever the reference key is not part of a tie group (a neighbour within RTOL in the same
         row) and is not the last position (whose tie partner may be the excluded (k+1)-th element).
    """
    ids, ref_ids = np.asarray(ids), np.asarray(ref_ids)
    keys, ref_keys, recomputed = (np.asarray(x, dtype=np.float32) for x in (keys, ref_keys, recomputed))
    assert ids.shape == ref_ids.shape, what
    if exact:
        assert np.array_equal(keys, ref_keys), f"{what}: keys not bit-exact (max abs diff {np.abs(keys - ref_keys).max()})"
        assert np.array_equal(recomputed, keys), f"{what}: reported key != oracle key of the returned id"
    else:
        assert np.allclose(keys, ref_keys, rtol=RTOL, atol=0), f"{what}: keys differ beyond {RTOL}"
        assert np.allclose(recomputed, keys, rtol=RTOL, atol=0), f"{what}: reported key != oracle key of the returned id"
    mism = ids != ref_ids
    if not mism.any():
        return
    tol = RTOL * np.abs(ref_keys) if not exact else np.zeros_like(ref_keys)
    k = ids.shape[1]
    tied = np.zeros_like(mism)
    if k > 1:
        close_next = np.abs(ref_keys[:, 1:] - ref_keys[:, :-1]) <= np.maximum(tol[:, 1:], tol[:, :-1])
        tied[:, 1:] |= close_next
        tied[:, :-1] |= close_next
    tied[:, -1] = True
    bad = mism & ~tied
    assert not bad.any(), f"{what}: {int(bad.sum())} id mismatches outside tie groups, first at {np.argwhere(bad)[0]}"
    # a mismatching id must not appear twice in a row
    for r in np.unique(np.argwhere(mism)[:, 0]):
        assert len(set(ids[r].tolist())) == k, f"{what}: duplicate ids in row {r}"


def load_golden_ivf(path, synth):
    """hp2_*.npz (tests/golden/gen_golden_ivf.py: the reference's IVFSearcher run on a scattered-layout directory)
    -> (golden, base, queries, centroids, labels, offsets, cluster_indices)."""
    g = np.load(path)
    base = synth.make("mix", int(g["base_seed"]), int(g["n"]))
    qry = synth.make("mix", int(g["query_seed"]), int(g["nq"]))
    cent = g["centroids"].astype(np.float32)
    labels = g["labels"].astype(np.int32)
    nlist = int(g["nlist"])
    # create_ivf_model.py:112-119: list i = np.where(cluster_ids == i) -> ascending original ids inside a list
    indices = np.argsort(labels, kind="stable").astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(labels, minlength=nlist))]).astype(np.int32)
    return g, base, qry, cent, labels, offsets, indices


def assert_ivf_matches_golden(g, nprobe, ids, scores, counts, total, recomputed, gt, what=""):
    """Reference (benchmark_ivf.IVFSearcher.search) vs ours on integer data: scores bit-exact position by position,
    candidate counts and their sum equal, ids equal outside score ties, recall@k equal per query set-wise whenever the
    k-th and (k+1)-th scores are not tied (ids of a tie group are interchangeable in the reference's argpartition)."""
    k = int(g["k"])
    rids, rsc, rcand = g[f"ids_np{nprobe}"], g[f"scores_np{nprobe}"], g[f"cand_np{nprobe}"]
    assert total == int(rcand.sum()), what
    assert np.array_equal(counts, np.minimum(rcand, k)), what
    assert_topk_matches(ids, scores, rids, rsc, recomputed, exact=True, what=what)
    rec = np.array([len(set(ids[i][ids[i] >= 0].tolist()) & set(gt[i][:k].tolist())) / k for i in range(ids.shape[0])])
    same_set = np.array([set(ids[i].tolist()) == set(rids[i].tolist()) for i in range(ids.shape[0])])
    assert np.array_equal(rec[same_set], g[f"recall_np{nprobe}"][same_set]), what
    assert same_set.mean() > 0.9, what
