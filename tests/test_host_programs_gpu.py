"""The C++ host programs that keep the reference's call shapes (hai-25-rag-on-edge_b200/host -> bin/), run as
subprocesses on synthetic .fvecs files and compared with the reference's own output / the CPU oracle:
  bin/cpu_baseline     vs the text written by the UNMODIFIED cpu/cpu_baseline.cpp (oracle/_ref/ref_driver)
  bin/qidk_ivf_search  vs the CPU restatement of IVFIndex::searchBatch (same results.txt format, recall in metrics.txt)
  bin/qidk_rag_demo    vs the CPU twin of the INT8 scheme (results.txt prints u8 * output_scale with 4 decimals)
"""
import os
import re
import subprocess

import numpy as np
import pytest

from util import RTOL, assert_topk_matches

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "hai-25-rag-on-edge_b200", "bin")


def run(cmd, cwd=None, expect_rc=0):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == expect_rc, f"{cmd}: rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return r


@pytest.mark.parametrize("law,k,extra", [("sift", 5, []), ("cont", 10, []), ("sift", 5, ["--batch", "1"]),
                                         ("cont", 10, ["--precision", "3xtf32", "--batch", "32"]),
                                         ("cont", 10, ["--gpus", "1", "--shards-per-gpu", "4"]),   # vs_exact_mgpu_* path
                                         ("sift", 5, ["--gpus", "0"])])                             # every visible GPU
def test_cpu_baseline_cli_matches_reference_text(law, k, extra, tmp_path, gpu_vsb, oracle):
    vsb = gpu_vsb
    assert os.path.exists(os.path.join(BIN, "cpu_baseline")), "run __graft_entry__.build()"
    base = vsb.synth.make(law, 1234, 10_000)   # BASELINE configs[0] shape
    qry = vsb.synth.make(law, 4321, 100)
    bf, qf = str(tmp_path / "base.fvecs"), str(tmp_path / "query.fvecs")
    vsb.synth.write_fvecs(bf, base)
    vsb.synth.write_fvecs(qf, qry)
    out = str(tmp_path / "ours.txt")
    r = run([os.path.join(BIN, "cpu_baseline"), bf, qf, str(k), out] + extra)
    assert "Throughput:" in r.stdout and "queries/sec" in r.stdout
    ids, d = oracle.parse_results_txt(out)
    assert ids.shape == (100, k)
    if oracle.have_ref():   # the unmodified reference's own text output
        ref_txt = str(tmp_path / "ref.txt")
        oracle.ref_bench(bf, qf, k, ref_txt, threads=4)
        rids, rd = oracle.parse_results_txt(ref_txt)
        if law == "sift":   # integer data: same characters, line by line, wherever no tie reorders ids
            ours, theirs = open(out).read().splitlines(), open(ref_txt).read().splitlines()
            assert len(ours) == len(theirs)
            same = sum(a == b for a, b in zip(ours, theirs))
            assert same >= 0.9 * len(ours)
            assert np.array_equal(d, rd)
    else:
        rids, rd = oracle.exact_search(base, qry, k, mode=1)
    # 6 significant digits in the text: compare at that resolution
    full_ids, full_d = oracle.exact_search(base, qry, k, mode=1)
    rec = oracle.exact_distances_at(base, qry, ids)
    assert np.allclose(d, rec, rtol=2e-6 + RTOL, atol=0)
    assert np.allclose(d, rd, rtol=2e-6 + RTOL, atol=0)
    assert (ids == full_ids).mean() > 0.99


def test_cpu_baseline_no_argument_mode_and_errors(tmp_path, gpu_vsb, oracle):
    vsb = gpu_vsb
    os.makedirs(tmp_path / "siftsmall")
    base = vsb.synth.make("sift", 5, 2000)
    qry = vsb.synth.make("sift", 6, 20)
    vsb.synth.write_fvecs(str(tmp_path / "siftsmall" / "siftsmall_base.fvecs"), base)
    vsb.synth.write_fvecs(str(tmp_path / "siftsmall" / "siftsmall_query.fvecs"), qry)
    r = run([os.path.join(BIN, "cpu_baseline")], cwd=str(tmp_path))   # sift/ is missing: reported, skipped, rc 0
    assert "Cannot open file" in r.stderr
    ids, d = oracle.parse_results_txt(str(tmp_path / "siftsmall_results.txt"))
    oi, od = oracle.exact_search(base, qry, 5, mode=1)   # hard-coded k = 5 (cpu_baseline.cpp:329)
    assert np.array_equal(ids, oi) and np.array_equal(d, od)
    assert not os.path.exists(tmp_path / "sift_results.txt")
    # truncated file
    raw = open(tmp_path / "siftsmall" / "siftsmall_base.fvecs", "rb").read()
    open(tmp_path / "trunc.fvecs", "wb").write(raw[:-7])
    r = run([os.path.join(BIN, "cpu_baseline"), str(tmp_path / "trunc.fvecs"),
             str(tmp_path / "siftsmall" / "siftsmall_query.fvecs"), "5", str(tmp_path / "x.txt")], expect_rc=1)
    assert "File seems truncated." in r.stderr and not os.path.exists(tmp_path / "x.txt")
    run([os.path.join(BIN, "cpu_baseline"), "a", "b"], expect_rc=1)


def test_ivf_driver_matches_restatement(tmp_path, gpu_vsb, oracle):
    vsb = gpu_vsb
    n, nlist, nq, k, nprobe, batch = 50_000, 128, 203, 10, 8, 16
    base = vsb.synth.make("mix", 21, n)
    qry = vsb.synth.make("mix", 22, nq)
    idx_dir = str(tmp_path / "ivf")
    vsb.ivf_build(base, nlist, idx_dir, max_iter=5, seed=1, reordered=True)
    gt, _ = oracle.exact_search(base, qry, k, mode=1)
    qf, gf, rd = str(tmp_path / "q.fvecs"), str(tmp_path / "gt.ivecs"), str(tmp_path / "res")
    vsb.synth.write_fvecs(qf, qry)
    vsb.synth.write_ivecs(gf, gt)
    r = run([os.path.join(BIN, "qidk_ivf_search"), idx_dir, qf, rd, "./libQnnHtp.so", str(k), str(nprobe), gf, str(batch)])
    assert "IVF Search Complete" in r.stdout
    ids, sc = oracle.parse_results_txt(os.path.join(rd, "results.txt"))
    assert ids.shape == (nq, k)
    # CPU restatement on the same index files
    vec = np.load(os.path.join(idx_dir, "vectors_reordered.npy"))
    off = np.load(os.path.join(idx_dir, "cluster_offsets.npy"))
    r2o = np.load(os.path.join(idx_dir, "reorder_to_original.npy"))
    cent = np.load(os.path.join(idx_dir, "centroids.npy"))
    coarse = oracle.ivf_coarse(qry, cent)
    rids, rsc, _, rtotal = oracle.ivf_search(vec, off, r2o, True, coarse, qry, k, nprobe, mode=1)
    assert np.array_equal(ids, rids)
    assert np.allclose(sc, rsc, rtol=0, atol=5.1e-5 + 1e-7 * np.abs(rsc))  # text has 4 decimals
    met = open(os.path.join(rd, "metrics.txt")).read()
    for label in ("Index Configuration:", "Avg candidates searched:", "Candidate reduction:", "Recall@10:", "Batch P50:",
                  "QPS:", "FLOPs per query (fine):", "Time Breakdown:"):
        assert label in met, label
    rec = float(re.search(r"Recall@10: ([0-9.]+)%", met).group(1))
    want = 100.0 * np.mean([len(set(rids[i]) & set(gt[i])) / k for i in range(nq)])
    assert abs(rec - want) < 1e-3
    # candidates: padded last batch (203 = 12*16 + 11 -> 5 zero queries) counts too, like the reference
    avg = float(re.search(r"Avg candidates searched: ([0-9.]+)", met).group(1))
    assert avg * nq >= rtotal
    run([os.path.join(BIN, "qidk_ivf_search"), str(tmp_path / "nope"), qf, rd, "x.so", "10"], expect_rc=1)


def test_rag_demo_driver_matches_int8_twin(tmp_path, gpu_vsb, oracle):
    vsb = gpu_vsb
    docs = vsb.synth.make("sift", 31, 30_000)
    qry = vsb.synth.make("sift", 32, 70)
    df, qf, rd = str(tmp_path / "docs.fvecs"), str(tmp_path / "q.fvecs"), str(tmp_path / "res")
    vsb.synth.write_fvecs(df, docs)
    vsb.synth.write_fvecs(qf, qry)
    k = 10
    for batch in ("1", "32"):
        r = run([os.path.join(BIN, "qidk_rag_demo"), "model.bin", qf, rd, "./libQnnHtp.so", df, str(k), batch])
        assert "Search Complete" in r.stdout
        ids, sc = oracle.parse_results_txt(os.path.join(rd, "results.txt"))
        s_w = float(np.float32(docs.max()) / np.float32(255.0))
        m = oracle.int8_multiplier(vsb.QNN_INPUT_SCALE, s_w, vsb.QNN_OUTPUT_SCALE)
        want_ids, want_sc = oracle.int8_search(oracle.quantize_u8(docs, s_w), oracle.quantize_u8(qry, vsb.QNN_INPUT_SCALE), k, m, mode=1)
        assert np.array_equal(ids, want_ids)
        assert np.allclose(sc, want_sc.astype(np.float32) * np.float32(vsb.QNN_OUTPUT_SCALE), rtol=1e-6, atol=6e-5)
        assert "QPS:" in open(os.path.join(rd, "metrics.txt")).read()
