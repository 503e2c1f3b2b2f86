"""GPU parity tests for the INT8 brute-force path (QNN quantisation scheme), through the C ABI, against the CPU
twin in oracle/ (oracle.quantize_u8 / int8_scores / int8_search).  Everything here is integer/byte work: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _calibrated_scales(vsb, oracle, base, qry):
    """Mirror of convert_to_qnn.sh using the query file as calibration data: encodings = max/255, offset 0."""
    s_in = float(qry.max()) / 255.0
    s_w = float(base.max()) / 255.0
    acc_max = (oracle.quantize_u8(qry, s_in).astype(np.int64) @ oracle.quantize_u8(base[:2048], s_w).astype(np.int64).T).max()
    s_out = float(acc_max) * s_in * s_w / 255.0
    return s_in, s_w, s_out


def test_quantiser_matches_quantize_buffer_neon(gpu_vsb, oracle):
    vsb = gpu_vsb
    rng = np.random.default_rng(5)
    x = np.concatenate([
        rng.uniform(-50, 400, 100_003).astype(np.float32),
        (np.arange(0, 512, dtype=np.float32) * np.float32(0.6627451) * np.float32(0.5)),  # .5 boundaries
        np.array([0.0, -0.0, 1e30, -1e30, np.nan, np.inf, -np.inf, 168.99998, 169.0, 169.33], dtype=np.float32),
    ])
    for scale in (vsb.QNN_INPUT_SCALE, 1.0, 0.8549, 218.0 / 255.0):
        assert np.array_equal(vsb.int8_quantize(x, scale), oracle.quantize_u8(x, scale)), scale


@pytest.mark.parametrize("law", ["sift", "mix"])
@pytest.mark.parametrize("nb,nq,k", [(1, 1, 1), (127, 3, 5), (129, 33, 10), (4097, 129, 10), (20000, 257, 32), (3001, 32, 16)])
def test_search_and_raw_scores_vs_cpu_twin(law, nb, nq, k, gpu_vsb, oracle):
    vsb = gpu_vsb
    k = min(k, nb)
    base = vsb.synth.make(law, 7000 + nb, nb)
    qry = vsb.synth.make(law, 8000 + nq, nq)
    for scales in ("reference", "calibrated"):
        if scales == "reference":  # the constants baked into QnnRunner.cpp:70-71, weights by min/max
            s_in, s_w, s_out = vsb.QNN_INPUT_SCALE, 0.0, vsb.QNN_OUTPUT_SCALE
        else:
            s_in, s_w, s_out = _calibrated_scales(vsb, oracle, base, qry)
        idx = vsb.Int8Index(base, s_in, s_w, s_out)
        try:
            if s_w == 0.0:
                assert idx.w_scale == np.float32(base.max()) / np.float32(255.0) or base.max() == 0
            m = oracle.int8_multiplier(idx.in_scale, idx.w_scale, idx.out_scale)
            assert m == idx.multiplier
            b8 = oracle.quantize_u8(base, idx.w_scale)
            q8 = oracle.quantize_u8(qry, idx.in_scale)
            want_ids, want_sc = oracle.int8_search(b8, q8, k, m, mode=1)
            ids, sc = idx.search(qry, k)
            assert np.array_equal(sc, want_sc), f"{law} {nb}x{nq} k={k} {scales}: scores differ"
            assert np.array_equal(ids, want_ids), f"{law} {nb}x{nq} k={k} {scales}: ids differ"
            if nb * nq <= 2_000_000:
                assert np.array_equal(idx.scores_raw(qry), oracle.int8_scores(b8, q8, m))
        finally:
            idx.close()


def test_literal_heap_order_is_a_permutation_within_ties(gpu_vsb, oracle):
    """find_top_k_int8 (main.cpp:36-57) breaks ties by heap order; the GPU returns the canonical (score desc, id asc)
    order. The score vectors must be identical and ids may differ only inside groups of equal score."""
    vsb = gpu_vsb
    base = vsb.synth.make("sift", 31, 50_000)
    qry = vsb.synth.make("sift", 32, 64)
    idx = vsb.Int8Index(base)
    try:
        b8, q8 = oracle.quantize_u8(base, idx.w_scale), oracle.quantize_u8(qry, idx.in_scale)
        lit_ids, lit_sc = oracle.int8_search(b8, q8, 10, idx.multiplier, mode=0)
        ids, sc = idx.search(qry, 10)
        assert np.array_equal(sc, lit_sc)
        raw = oracle.int8_scores(b8, q8, idx.multiplier)
        assert np.array_equal(np.take_along_axis(raw, ids.astype(np.int64), 1), sc)
        strictly_better = sc[:, :1] > sc[:, -1:]  # ids above the k-th score level are forced
        for r in range(qry.shape[0]):
            forced = sc[r] > sc[r, -1]
            assert set(ids[r][forced]) == set(lit_ids[r][forced])
        assert strictly_better.shape == (64, 1)
    finally:
        idx.close()


def test_sharding_invariance_and_big_batch(gpu_vsb, oracle):
    """BASELINE configs[3] at reduced N here (2M x 128, batch 1024 and 32); the full 10M case is timed by
    tools/bench_paths.py. Properties: merge of two shard handles == one handle; sorted scores; scores of the returned
    ids recomputed by the CPU twin."""
    import torch

    vsb = gpu_vsb
    n, k = 2_000_000, 10
    dev = torch.device("cuda:0")
    base_d = torch.empty((n, 128), dtype=torch.float32, device=dev)
    vsb.synth_fill_dev(base_d.data_ptr(), 0, n, 128, "mix", 99)
    torch.cuda.synchronize()
    qry = vsb.synth.make("mix", 98, 1024)
    s_in, s_w, s_out = 218.0 / 255.0, 218.0 / 255.0, 9000.0
    one = vsb.Int8Index(base_d.data_ptr(), s_in, s_w, s_out, n=n)
    h = n // 2 + 77
    a = vsb.Int8Index(base_d.data_ptr(), s_in, s_w, s_out, n=h)
    b = vsb.Int8Index(base_d.data_ptr() + h * 512, s_in, s_w, s_out, n=n - h, id_base=h)
    try:
        ids, sc = one.search(qry, k)
        ids32, sc32 = one.search(qry[:32], k)
        assert np.array_equal(ids32, ids[:32]) and np.array_equal(sc32, sc[:32])
        assert (np.diff(sc.astype(np.int32), axis=1) <= 0).all()
        ia, sa = a.search(qry, k)
        ib, sb = b.search(qry, k)
        cat_ids = np.concatenate([ia, ib], 1)
        cat_sc = np.concatenate([sa, sb], 1).astype(np.int32)
        order = np.lexsort((cat_ids, -cat_sc), axis=1)[:, :k]
        assert np.array_equal(np.take_along_axis(cat_ids, order, 1), ids)
        assert np.array_equal(np.take_along_axis(cat_sc, order, 1), sc.astype(np.int32))
        # recompute the returned scores with the CPU twin on regenerated rows
        m = one.multiplier
        q8 = oracle.quantize_u8(qry[:64], one.in_scale)
        for r in range(0, 64, 7):
            rows = np.concatenate([vsb.synth.rows("mix", 99, int(i), 1) for i in ids[r]])
            raw = oracle.int8_scores(oracle.quantize_u8(rows, one.w_scale), q8[r:r + 1], m)
            assert np.array_equal(raw[0], sc[r])
        # no row of a 100K sample beats the k-th score
        samp = vsb.synth.rows("mix", 99, 1_234_000, 100_000)
        raw = oracle.int8_scores(oracle.quantize_u8(samp, one.w_scale), q8[:16], m)
        assert (raw.max(1) <= sc[:16, 0]).all()
        assert ((raw > sc[:16, -1:]).sum(1) <= k).all()
    finally:
        one.close()
        a.close()
        b.close()


def test_argument_errors(gpu_vsb):
    vsb = gpu_vsb
    base = vsb.synth.make("sift", 1, 100)
    with pytest.raises(vsb.VsbError):
        vsb.Int8Index(base, in_scale=0.0)
    idx = vsb.Int8Index(base)
    try:
        with pytest.raises(vsb.VsbError):
            idx.search(base[:2], 101)
        with pytest.raises(vsb.VsbError):
            idx.search(base[:2], 33)  # k > 32 not implemented on the INT8 path
        ids, sc = idx.search(base[:0], 5)
        assert ids.shape == (0, 5)
    finally:
        idx.close()


def test_full_size_10m_parity(gpu_vsb, oracle):
    """BASELINE configs[3] at its full size: 10M x 128 rows (generated on the device), batches of 1024 and 32 queries,
    top-10.  Exactness is established in three links, each bit-exact:
      1. a 1M-row window of the base, searched through its own handle, equals the CPU twin's top-10 on the same rows
         (ids and u8 scores, canonical order) for 96 queries;
      2. the 10M answer equals the merge of the ten 1M-window handles' answers (row sharding is invisible);
      3. every score the 10M handle returns equals the CPU twin's score of that (query, row) pair, for all 1024 queries;
    plus batch 32 == the first 32 rows of batch 1024 and descending order."""
    import torch

    vsb = gpu_vsb
    n, k, W = 10_000_000, 10, 1_000_000
    dev = torch.device("cuda:0")
    base_d = torch.empty((n, 128), dtype=torch.float32, device=dev)
    for c0 in range(0, n, 1 << 22):
        vsb.synth_fill_dev(base_d.data_ptr() + c0 * 512, c0, min(1 << 22, n - c0), 128, "mix", 99)
    torch.cuda.synchronize()
    qry = vsb.synth.make("mix", 98, 1024)
    s_in, s_w, s_out = 218.0 / 255.0, 218.0 / 255.0, 9000.0
    one = vsb.Int8Index(base_d.data_ptr(), s_in, s_w, s_out, n=n)
    try:
        ids, sc = one.search(qry, k)
        ids32, sc32 = one.search(qry[:32], k)
        m, in_scale, w_scale = one.multiplier, one.in_scale, one.w_scale
    finally:
        one.close()
    assert np.array_equal(ids32, ids[:32]) and np.array_equal(sc32, sc[:32])
    assert (np.diff(sc.astype(np.int32), axis=1) <= 0).all() and ids.min() >= 0 and ids.max() < n
    q8 = oracle.quantize_u8(qry, in_scale)
    # 2. merge of the ten window handles
    parts_i, parts_s = [], []
    for w0 in range(0, n, W):
        h = vsb.Int8Index(base_d.data_ptr() + w0 * 512, s_in, s_w, s_out, n=W, id_base=w0)
        try:
            wi, ws = h.search(qry, k)
        finally:
            h.close()
        parts_i.append(wi)
        parts_s.append(ws)
        if w0 == 3 * W:   # 1. this window against the CPU twin
            rows = base_d[w0:w0 + W].cpu().numpy()
            ti, ts = oracle.int8_search(oracle.quantize_u8(rows, w_scale), q8[:96], k, m, mode=1)
            assert np.array_equal(wi[:96], ti + w0) and np.array_equal(ws[:96], ts)
    cat_i = np.concatenate(parts_i, 1)
    cat_s = np.concatenate(parts_s, 1).astype(np.int32)
    order = np.lexsort((cat_i, -cat_s), axis=1)[:, :k]
    assert np.array_equal(np.take_along_axis(cat_i, order, 1), ids)
    assert np.array_equal(np.take_along_axis(cat_s, order, 1), sc.astype(np.int32))
    # 3. every returned score recomputed by the CPU twin
    uniq, inv = np.unique(ids.ravel(), return_inverse=True)
    rows_u8 = oracle.quantize_u8(base_d[torch.from_numpy(uniq.astype(np.int64)).to(dev)].cpu().numpy(), w_scale)
    assert np.array_equal(rows_u8[:64], oracle.quantize_u8(np.concatenate([vsb.synth.rows("mix", 99, int(i), 1) for i in uniq[:64]]), w_scale))
    pos = inv.reshape(ids.shape)
    want = np.stack([oracle.int8_scores(rows_u8[pos[r]], q8[r:r + 1], m)[0] for r in range(ids.shape[0])])
    assert np.array_equal(want, sc)
