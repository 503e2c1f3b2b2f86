"""TEST INFRASTRUCTURE — ctypes front-end to the CPU oracle (oracle/_build/liborc.so) and to the compiled,
unmodified reference (oracle/_ref/ref_driver).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liborc.so")
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")

_lib = None


def build(ref: bool = True) -> None:
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_ivf_search.restype = C.c_int64
        _lib.orc_int8_multiplier.restype = C.c_float
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def have_ref() -> bool:
    return os.path.exists(REF_DRIVER)


# ------------------------------------------------------------------ HP1
def norms(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty(x.shape[0], dtype=np.float32)
    lib().orc_norms(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(out))
    return out


def exact_search(base, queries, k: int, mode: int = 1):
    """mode 0: literal select_topk ties, 1: canonical (dist asc, id asc). -> (ids int32 [nq,k], dists f32)."""
    base, queries = _f32(base), _f32(queries)
    nq = queries.shape[0]
    ids = np.empty((nq, k), dtype=np.int32)
    d = np.empty((nq, k), dtype=np.float32)
    rc = lib().orc_exact_search(_p(base), C.c_int64(base.shape[0]), C.c_int(base.shape[1]), _p(queries),
                                C.c_int64(nq), C.c_int(k), C.c_int(mode), None, _p(ids), _p(d))
    if rc != 0:
        raise ValueError("orc_exact_search: bad arguments")
    return ids, d


def exact_distances_at(base, queries, ids) -> np.ndarray:
    base, queries = _f32(base), _f32(queries)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    out = np.empty(ids.shape, dtype=np.float32)
    lib().orc_exact_distances_at(_p(base), C.c_int(base.shape[1]), _p(queries), C.c_int64(queries.shape[0]),
                                 C.c_int(ids.shape[1]), _p(ids), _p(out))
    return out


# ------------------------------------------------------------------ the real reference (HP1 only)
def ref_dump(base_fvecs: str, query_fvecs: str, k: int, threads: int | None = None):
    """Run the unmodified reference functions (oracle/ref_driver.cpp `dump`) -> ids, dists, qnorms, bnorms."""
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "dump.bin")
        env = dict(os.environ)
        if threads:
            env["OMP_NUM_THREADS"] = env["OPENBLAS_NUM_THREADS"] = str(threads)
        subprocess.run([REF_DRIVER, "dump", base_fvecs, query_fvecs, str(k), out], check=True, env=env)
        raw = np.fromfile(out, dtype=np.int32)
    nq, kk = int(raw[0]), int(raw[1])
    o = 2
    ids = raw[o : o + nq * kk].reshape(nq, kk).copy(); o += nq * kk
    dists = raw[o : o + nq * kk].view(np.float32).reshape(nq, kk).copy(); o += nq * kk
    qn = raw[o : o + nq].view(np.float32).copy(); o += nq
    nb = int(raw[o]); o += 1
    bn = raw[o : o + nb].view(np.float32).copy()
    return ids, dists, qn, bn


def ref_bench(base_fvecs: str, query_fvecs: str, k: int, results_txt: str, threads: int | None = None) -> dict:
    """run_benchmark() of the unmodified reference (its own timed loop, cpu_baseline.cpp:220-257).
    Returns {'qps', 'total_s', 'stdout'}; results_txt gets the reference's text output."""
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = env["OPENBLAS_NUM_THREADS"] = str(threads)
    r = subprocess.run([REF_DRIVER, "bench", "synthetic", base_fvecs, query_fvecs, str(k), results_txt],
                       check=True, env=env, capture_output=True, text=True)
    qps = float(re.search(r"Throughput:\s*([0-9.eE+-]+)\s*queries/sec", r.stdout).group(1))
    tot = float(re.search(r"Total execution time:\s*([0-9.eE+-]+)\s*s", r.stdout).group(1))
    return {"qps": qps, "total_s": tot, "stdout": r.stdout}


def parse_results_txt(path: str):
    """'Query i: (idx, dist) ...' (cpu_baseline.cpp:167-172) -> ids int32 [nq,k], dists float32 (6 sig. digits)."""
    ids, dists = [], []
    pat = re.compile(r"\((-?\d+), ([^)]+)\)")
    with open(path) as f:
        for line in f:
            m = pat.findall(line)
            ids.append([int(a) for a, _ in m])
            dists.append([float(b) for _, b in m])
    return np.array(ids, dtype=np.int32), np.array(dists, dtype=np.float32)


# ------------------------------------------------------------------ HP2
def ivf_coarse(queries, centroids) -> np.ndarray:
    queries, centroids = _f32(queries), _f32(centroids)
    out = np.empty((queries.shape[0], centroids.shape[0]), dtype=np.float32)
    lib().orc_ivf_coarse(_p(queries), C.c_int64(queries.shape[0]), _p(centroids), C.c_int(centroids.shape[0]),
                         C.c_int(queries.shape[1]), _p(out))
    return out


def ivf_select_probes(scores_row, nprobe: int) -> np.ndarray:
    s = _f32(scores_row)
    out = np.empty(nprobe, dtype=np.int32)
    lib().orc_ivf_select_probes(_p(s), C.c_int(s.shape[0]), C.c_int(nprobe), _p(out))
    return out


def ivf_search(vectors, offsets, id_map, reordered: bool, coarse_scores, queries, k: int, nprobe: int, mode: int = 1):
    vectors, queries, coarse_scores = _f32(vectors), _f32(queries), _f32(coarse_scores)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    id_map = np.ascontiguousarray(id_map, dtype=np.int32)
    nq, nlist = queries.shape[0], offsets.shape[0] - 1
    ids = np.empty((nq, k), dtype=np.int32)
    sc = np.empty((nq, k), dtype=np.float32)
    cnt = np.empty(nq, dtype=np.int32)
    total = lib().orc_ivf_search(_p(vectors), _p(offsets), _p(id_map), C.c_int(int(reordered)), C.c_int(nlist),
                                 C.c_int(vectors.shape[1]), _p(coarse_scores), _p(queries), C.c_int64(nq),
                                 C.c_int(k), C.c_int(nprobe), C.c_int(mode), _p(ids), _p(sc), _p(cnt))
    return ids, sc, cnt, int(total)


def ivf_scores_at(vectors_original, queries, ids) -> np.ndarray:
    v, q = _f32(vectors_original), _f32(queries)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    out = np.empty(ids.shape, dtype=np.float32)
    lib().orc_ivf_scores_at(_p(v), C.c_int(v.shape[1]), _p(q), C.c_int64(q.shape[0]), C.c_int(ids.shape[1]),
                            _p(ids), _p(out))
    return out


def kmeans_assign(x, centroids):
    x, centroids = _f32(x), _f32(centroids)
    lab = np.empty(x.shape[0], dtype=np.int32)
    bd = np.empty(x.shape[0], dtype=np.float32)
    lib().orc_kmeans_assign(_p(x), C.c_int64(x.shape[0]), _p(centroids), C.c_int(centroids.shape[0]),
                            C.c_int(x.shape[1]), _p(lab), _p(bd))
    return lab, bd


# ------------------------------------------------------------------ INT8
def quantize_u8(x, scale: float) -> np.ndarray:
    x = _f32(x)
    inv = np.float32(1.0) / np.float32(scale)
    out = np.empty(x.shape, dtype=np.uint8)
    lib().orc_quantize_u8(_p(x), C.c_int64(x.size), C.c_float(float(inv)), _p(out))
    return out


def int8_multiplier(s_in: float, s_w: float, s_out: float) -> float:
    return float(lib().orc_int8_multiplier(C.c_float(s_in), C.c_float(s_w), C.c_float(s_out)))


def int8_scores(base_u8, q_u8, m: float) -> np.ndarray:
    b = np.ascontiguousarray(base_u8, dtype=np.uint8)
    q = np.ascontiguousarray(q_u8, dtype=np.uint8)
    out = np.empty((q.shape[0], b.shape[0]), dtype=np.uint8)
    lib().orc_int8_scores(_p(b), C.c_int64(b.shape[0]), C.c_int(b.shape[1]), _p(q), C.c_int64(q.shape[0]),
                          C.c_float(m), _p(out))
    return out


def int8_search(base_u8, q_u8, k: int, m: float, mode: int = 1):
    b = np.ascontiguousarray(base_u8, dtype=np.uint8)
    q = np.ascontiguousarray(q_u8, dtype=np.uint8)
    ids = np.empty((q.shape[0], k), dtype=np.int32)
    sc = np.empty((q.shape[0], k), dtype=np.uint8)
    rc = lib().orc_int8_search(_p(b), C.c_int64(b.shape[0]), C.c_int(b.shape[1]), _p(q), C.c_int64(q.shape[0]),
                               C.c_int(k), C.c_float(m), C.c_int(mode), _p(ids), _p(sc))
    if rc != 0:
        raise ValueError("orc_int8_search: bad arguments")
    return ids, sc
