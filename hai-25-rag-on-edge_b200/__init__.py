"""vsb200 — thin ctypes front-end over the C ABI of libvsb200.so (include/vsb200.h).

The product is the shared library (hand-written sm_100a CUDA behind plain C entry points) plus the C++ host
programs in host/.  This module only exists so that the tests, bench.py and __graft_entry__ can drive the C ABI
from Python; it adds no compute of its own and has NO fallback: if the library is missing or no CUDA device is
present, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import synth  # noqa: F401  (re-exported: seeded synthetic SIFT-shaped data)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSB200_LIB") or os.path.join(HERE, "libvsb200.so")  # VSB200_LIB: build variants (tools/)

PREC_AUTO, PREC_3XTF32, PREC_FFMA, PREC_TF32_1X, PREC_F16_CERT = 0, 1, 2, 3, 4
PREC_NAMES = {0: "auto", 1: "fp32_3xtf32", 2: "fp32_ffma", 3: "tf32_1x", 4: "f16_certified+fp32_refine"}


class VsbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vsb200 error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> None:
    """Compile every CUDA source for sm_100a into libvsb200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", HERE, "-j8", "all"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("libvsb200 build failed:\n" + (r.stdout or "") + (r.stderr or ""))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VsbError(-1, f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.vs_last_error.restype = C.c_char_p
        L.vs_exact_size.restype = C.c_int64
        L.vs_topk_block_bytes.restype = C.c_size_t
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise VsbError(rc, lib().vs_last_error().decode(errors="replace"))


def device_count() -> int:
    c = C.c_int(0)
    rc = lib().vs_device_count(C.byref(c))
    return c.value if rc == 0 else 0


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw device pointer


class ExactIndex:
    """Exact squared-L2 kNN over a device-resident base; mirrors run_benchmark()'s hot triple
    (cpu/cpu_baseline.cpp:211-248) for a whole query batch."""

    def __init__(self, base, device: int = 0, id_base: int = 0, n: int | None = None, dim: int = 128):
        self._h = C.c_void_p()
        if isinstance(base, np.ndarray):
            base = np.ascontiguousarray(base, dtype=np.float32)
            n, dim = base.shape
            _check(lib().vs_exact_create(C.byref(self._h), _ptr(base), C.c_int64(n), C.c_int(dim), C.c_int(device),
                                         C.c_int64(id_base)))
        else:  # device pointer
            _check(lib().vs_exact_create_dev(C.byref(self._h), _ptr(base), C.c_int64(n), C.c_int(dim), C.c_int(device),
                                             C.c_int64(id_base)))
        self.n, self.dim, self.device = int(n), int(dim), device

    def close(self) -> None:
        if self._h:
            lib().vs_exact_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def base_is_tf32_exact(self) -> bool:
        return bool(lib().vs_exact_base_is_tf32_exact(self._h))

    def search(self, queries: np.ndarray, k: int, precision: int = PREC_AUTO, out_ids=None, out_dists=None):
        """Host buffers in, host buffers out (H2D + search + D2H inside the call)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        ids = out_ids if out_ids is not None else np.empty((nq, k), dtype=np.int32)
        d = out_dists if out_dists is not None else np.empty((nq, k), dtype=np.float32)
        _check(lib().vs_exact_search_f32(self._h, _ptr(q), C.c_int64(nq), C.c_int(k), C.c_int(precision), _ptr(ids),
                                         _ptr(d)))
        return ids, d

    def search_dev(self, q_ptr: int, nq: int, k: int, precision: int, ids_ptr: int, dists_ptr: int, stream: int = 0):
        """Device pointers, asynchronous on `stream` (a raw cudaStream_t value)."""
        _check(lib().vs_exact_search_dev(self._h, _ptr(q_ptr), C.c_int64(nq), C.c_int(k), C.c_int(precision),
                                         _ptr(ids_ptr), _ptr(dists_ptr), C.c_void_p(stream)))

    def search_dev_begin(self, q_ptr: int, nq: int, k: int, precision: int, ids_ptr: int, dists_ptr: int, stream: int = 0):
        """First half of search_dev: enqueues only (no host synchronisation); pair with search_dev_finish()."""
        _check(lib().vs_exact_search_dev_begin(self._h, _ptr(q_ptr), C.c_int64(nq), C.c_int(k), C.c_int(precision),
                                               _ptr(ids_ptr), _ptr(dists_ptr), C.c_void_p(stream)))

    def search_dev_finish(self) -> int:
        """Second half: waits for the certification count, redoes uncertified queries; -> result rows rewritten."""
        n = C.c_int(0)
        _check(lib().vs_exact_search_dev_finish(self._h, C.byref(n)))
        return n.value

    def debug_f16_candidates(self, queries: np.ndarray):
        """Test hook: the fp16 candidate pass alone -> (ids [nq,32] local, keys [nq,32] as the kernel ranked them,
        bound [nq] = the certificate's E_q)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, 32), dtype=np.int32)
        keys = np.empty((nq, 32), dtype=np.float32)
        bound = np.empty(nq, dtype=np.float32)
        _check(lib().vs_exact_debug_f16_candidates(self._h, _ptr(q), C.c_int64(nq), _ptr(ids), _ptr(keys), _ptr(bound)))
        return ids, keys, bound

    def set_profile(self, enable: bool = True) -> None:
        _check(lib().vs_exact_set_profile(self._h, C.c_int(int(enable))))

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib().vs_exact_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def last_prepass_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib().vs_exact_last_prepass_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def last_launches(self):
        a, b = C.c_int(0), C.c_int(0)
        _check(lib().vs_exact_last_launches(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_fallbacks(self) -> int:
        """queries of the last certified search that had to be redone on the fp32 path"""
        a = C.c_int(0)
        _check(lib().vs_exact_last_fallbacks(self._h, C.byref(a)))
        return a.value


def merge_topk_dev(ids_ptr: int, keys_ptr: int, n_shards: int, nq: int, k: int, smallest: bool, out_ids_ptr: int,
                   out_keys_ptr: int, stream: int = 0) -> None:
    _check(lib().vs_merge_topk_dev(_ptr(ids_ptr), _ptr(keys_ptr), C.c_int(n_shards), C.c_int64(nq), C.c_int(k),
                                   C.c_int(int(smallest)), _ptr(out_ids_ptr), _ptr(out_keys_ptr), C.c_void_p(stream)))


def topk_block_bytes(nq: int, k: int) -> int:
    """bytes of one shard's exchange block: ids [nq,k] i32 | keys [nq,k] f32 | 16-byte trailer (uncertified count)"""
    return int(lib().vs_topk_block_bytes(C.c_int64(nq), C.c_int(k)))


def push_block_dev(src_ptr: int, dst_ptrs, nbytes: int, flag_ptrs, epoch: int, counter_ptr: int, stream: int = 0) -> None:
    """one kernel: `nbytes` from src_ptr to the same bytes at every pointer of dst_ptrs (peer-mapped device memory), then
    `epoch` into every flag word of flag_ptrs"""
    n = len(flag_ptrs)
    arr = (C.c_void_p * n)(*[int(p) for p in dst_ptrs]) if nbytes else (C.c_void_p * n)()
    flg = (C.c_void_p * n)(*[int(p) for p in flag_ptrs])
    _check(lib().vs_push_block_dev(_ptr(src_ptr), arr, C.c_int(n), C.c_size_t(nbytes), flg, C.c_uint32(epoch & 0xffffffff),
                                   _ptr(counter_ptr), C.c_void_p(stream)))


def wait_flags_dev(flags_ptr: int, n: int, self_index: int, epoch: int, stream: int = 0) -> None:
    _check(lib().vs_wait_flags_dev(_ptr(flags_ptr), C.c_int(n), C.c_int(self_index), C.c_uint32(epoch & 0xffffffff),
                                   C.c_void_p(stream)))


def merge_blocks_dev(blocks_ptr: int, n_shards: int, stride: int, nq: int, k: int, smallest: bool, out_ids_ptr: int,
                     out_keys_ptr: int, total_ptr: int = 0, stream: int = 0) -> None:
    _check(lib().vs_merge_blocks_dev(_ptr(blocks_ptr), C.c_int(n_shards), C.c_size_t(stride), C.c_int64(nq), C.c_int(k),
                                     C.c_int(int(smallest)), _ptr(out_ids_ptr), _ptr(out_keys_ptr),
                                     _ptr(total_ptr) if total_ptr else None, C.c_void_p(stream)))


class ExactGroup:
    """The exact-search shards of ONE device as one participant of the row-sharded search (vs_exact_group_*):
    begin -> [the caller's exchange of the gathered buffer] -> merge -> finish."""

    def __init__(self, shards, n_slots: int, first_slot: int):
        self._h = C.c_void_p()
        self._shards = list(shards)  # keep the ExactIndex objects alive: the group does not own them
        arr = (C.c_void_p * len(self._shards))(*[s._h for s in self._shards])
        _check(lib().vs_exact_group_create_from(C.byref(self._h), C.c_int(len(self._shards)), arr, C.c_int(n_slots),
                                                C.c_int(first_slot)))

    def close(self) -> None:
        if self._h:
            lib().vs_exact_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def begin(self, q_ptr: int, nq: int, k: int, precision: int, gathered_ptr: int, stream: int = 0) -> None:
        _check(lib().vs_exact_group_begin(self._h, _ptr(q_ptr), C.c_int64(nq), C.c_int(k), C.c_int(precision),
                                          _ptr(gathered_ptr), C.c_void_p(stream)))

    def merge(self, out_ids_ptr: int, out_dists_ptr: int) -> None:
        _check(lib().vs_exact_group_merge(self._h, _ptr(out_ids_ptr), _ptr(out_dists_ptr)))

    def finish(self) -> bool:
        """-> True when some shard redid uncertified queries: exchange, merge and finish again."""
        need = C.c_int(0)
        _check(lib().vs_exact_group_finish(self._h, C.byref(need)))
        return bool(need.value)


class ExactMultiGpu:
    """Single-process multi-GPU exact search (vs_exact_mgpu_*): host buffers in, host buffers out."""

    def __init__(self, base: np.ndarray, n_gpus: int = 0, shards_per_gpu: int = 1):
        self._h = C.c_void_p()
        base = np.ascontiguousarray(base, dtype=np.float32)
        _check(lib().vs_exact_mgpu_create(C.byref(self._h), _ptr(base), C.c_int64(base.shape[0]), C.c_int(base.shape[1]),
                                          C.c_int(n_gpus), C.c_int(shards_per_gpu)))
        self.n_gpus = int(lib().vs_exact_mgpu_num_gpus(self._h))
        self.n_shards = int(lib().vs_exact_mgpu_num_shards(self._h))

    def close(self) -> None:
        if self._h:
            lib().vs_exact_mgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, queries: np.ndarray, k: int, precision: int = PREC_AUTO, out_ids=None, out_dists=None):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        ids = out_ids if out_ids is not None else np.empty((nq, k), dtype=np.int32)
        d = out_dists if out_dists is not None else np.empty((nq, k), dtype=np.float32)
        _check(lib().vs_exact_mgpu_search_f32(self._h, _ptr(q), C.c_int64(nq), C.c_int(k), C.c_int(precision), _ptr(ids),
                                              _ptr(d)))
        return ids, d

    def last_stats(self):
        """-> (exchanges of the last search, whether any shard redid uncertified queries)"""
        a, b = C.c_int(0), C.c_int(0)
        _check(lib().vs_exact_mgpu_last_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, bool(b.value)

    def set_profile(self, enable: bool = True) -> None:
        _check(lib().vs_exact_mgpu_set_profile(self._h, C.c_int(int(enable))))

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib().vs_exact_mgpu_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)


def synth_fill_dev(out_ptr: int, row0: int, nrows: int, dim: int, law: str, seed: int, centre_seed: int = 7,
                   stream: int = 0) -> None:
    _check(lib().vs_synth_fill_dev(_ptr(out_ptr), C.c_int64(row0), C.c_int64(nrows), C.c_int(dim),
                                   C.c_int(synth.LAWS[law]), C.c_uint64(seed), C.c_uint64(centre_seed),
                                   C.c_void_p(stream)))


class IvfIndex:
    """Two-stage IVF search (inner product, largest first); mirrors the reference's IVFIndex class surface
    (qidk_ivf/android/app/main/jni/IVFIndex.h:14-54) over the C ABI."""

    def __init__(self, index_dir: str | None = None, *, vectors=None, offsets=None, id_map=None, centroids=None,
                 device: int = 0):
        self._h = C.c_void_p()
        L = lib()
        L.vs_ivf_num_vectors.restype = C.c_int64
        L.vs_ivf_avg_cluster_size.restype = C.c_float
        if index_dir is not None:
            _check(L.vs_ivf_open(C.byref(self._h), index_dir.encode(), C.c_int(device)))
        else:
            v = np.ascontiguousarray(vectors, dtype=np.float32)
            o = np.ascontiguousarray(offsets, dtype=np.int32)
            m = np.ascontiguousarray(id_map, dtype=np.int32)
            c = np.ascontiguousarray(centroids, dtype=np.float32)
            _check(L.vs_ivf_create(C.byref(self._h), _ptr(v), C.c_int64(v.shape[0]), C.c_int(v.shape[1]), _ptr(o),
                                   C.c_int(o.shape[0] - 1), _ptr(m), _ptr(c), C.c_int(device)))
        self.num_vectors = int(L.vs_ivf_num_vectors(self._h))
        self.num_clusters = int(L.vs_ivf_num_clusters(self._h))
        self.dim = int(L.vs_ivf_dim(self._h))
        self.avg_cluster_size = float(L.vs_ivf_avg_cluster_size(self._h))

    def close(self) -> None:
        if self._h:
            lib().vs_ivf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search_batch(self, queries, k: int, nprobe: int, out=None):
        """-> (ids [nq,k] int32 (-1 padded), scores [nq,k] f32 descending, counts [nq], total candidates).
        `out` = (ids, scores, counts) arrays to fill, e.g. views of pinned host memory (vs_host_alloc)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        if out is not None:
            ids, sc, cnt = out
        else:
            ids = np.empty((nq, k), dtype=np.int32)
            sc = np.empty((nq, k), dtype=np.float32)
            cnt = np.empty(nq, dtype=np.int32)
        total = C.c_uint64(0)
        _check(lib().vs_ivf_search(self._h, _ptr(q), C.c_int64(nq), C.c_int(k), C.c_int(nprobe), _ptr(ids), _ptr(sc),
                                   _ptr(cnt), C.byref(total)))
        return ids, sc, cnt, int(total.value)

    def search_dev(self, q_ptr: int, nq: int, k: int, nprobe: int, ids_ptr: int, scores_ptr: int, counts_ptr: int,
                   stream: int = 0) -> None:
        _check(lib().vs_ivf_search_dev(self._h, _ptr(q_ptr), C.c_int64(nq), C.c_int(k), C.c_int(nprobe), _ptr(ids_ptr),
                                       _ptr(scores_ptr), _ptr(counts_ptr), C.c_void_p(stream)))

    def coarse_scores(self, queries) -> np.ndarray:
        q = np.ascontiguousarray(queries, dtype=np.float32)
        out = np.empty((q.shape[0], self.num_clusters), dtype=np.float32)
        _check(lib().vs_ivf_coarse_scores(self._h, _ptr(q), C.c_int64(q.shape[0]), _ptr(out)))
        return out

    def set_profile(self, enable: bool = True) -> None:
        _check(lib().vs_ivf_set_profile(self._h, C.c_int(int(enable))))

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib().vs_ivf_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)


def ivf_build(base, nlist: int, out_dir: str, *, max_iter: int = 100, seed: int = 42, reordered: bool = True,
              device: int = 0, init_centroids=None):
    """build_ivf_index() of the reference (create_ivf_model*.py) on the GPU -> {'nlist','iters','inertia'}."""
    b = np.ascontiguousarray(base, dtype=np.float32)
    ic = None if init_centroids is None else np.ascontiguousarray(init_centroids, dtype=np.float32)
    nl, it, inertia = C.c_int(0), C.c_int(0), C.c_double(0)
    _check(lib().vs_ivf_build(_ptr(b), C.c_int64(b.shape[0]), C.c_int(b.shape[1]), C.c_int(nlist), C.c_int(max_iter),
                              C.c_uint64(seed), out_dir.encode(), C.c_int(int(reordered)), C.c_int(device), _ptr(ic),
                              C.byref(nl), C.byref(it), C.byref(inertia)))
    return {"nlist": nl.value, "iters": it.value, "inertia": inertia.value}


QNN_INPUT_SCALE = 0.6627451    # QnnRunner.cpp:70 default u8 encoding of the query tensor (169/255)
QNN_OUTPUT_SCALE = 1013.4312   # QnnRunner.cpp:71 default u8 encoding of the score tensor


class Int8Index:
    """INT8 brute force (u8 x u8 -> s32 MatMul + requantisation + largest-k); mirrors the QnnRunner surface of the
    reference's qidk_rag_demo (qidk_bruteforce/android/app/main/jni/QnnRunner.h:20-52) plus find_top_k_int8
    (main.cpp:36-71) over the C ABI."""

    def __init__(self, base, in_scale: float = QNN_INPUT_SCALE, w_scale: float = 0.0, out_scale: float = QNN_OUTPUT_SCALE,
                 device: int = 0, id_base: int = 0, n: int | None = None, dim: int = 128):
        self._h = C.c_void_p()
        L = lib()
        L.vs_int8_num_docs.restype = C.c_int64
        L.vs_int8_output_scale.restype = C.c_float
        args = (C.c_float(in_scale), C.c_float(w_scale), C.c_float(out_scale), C.c_int(device), C.c_int64(id_base))
        if isinstance(base, np.ndarray):
            base = np.ascontiguousarray(base, dtype=np.float32)
            n, dim = base.shape
            _check(L.vs_int8_create(C.byref(self._h), _ptr(base), C.c_int64(n), C.c_int(dim), *args))
        else:
            _check(L.vs_int8_create_dev(C.byref(self._h), _ptr(base), C.c_int64(n), C.c_int(dim), *args))
        self.num_docs = int(L.vs_int8_num_docs(self._h))
        self.dim = int(L.vs_int8_dim(self._h))
        a, b, c, m = C.c_float(0), C.c_float(0), C.c_float(0), C.c_float(0)
        _check(L.vs_int8_scales(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(m)))
        self.in_scale, self.w_scale, self.out_scale, self.multiplier = a.value, b.value, c.value, m.value

    def close(self) -> None:
        if self._h:
            lib().vs_int8_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, queries, k: int):
        """fp32 queries (raw SIFT values) -> (ids [nq,k] int32, raw u8 scores [nq,k]), (score desc, id asc)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int32)
        sc = np.empty((nq, k), dtype=np.uint8)
        _check(lib().vs_int8_search(self._h, _ptr(q), C.c_int64(nq), C.c_int(k), _ptr(ids), _ptr(sc)))
        return ids, sc

    def search_dev(self, q_ptr: int, nq: int, k: int, ids_ptr: int, scores_ptr: int, stream: int = 0) -> None:
        _check(lib().vs_int8_search_dev(self._h, _ptr(q_ptr), C.c_int64(nq), C.c_int(k), _ptr(ids_ptr), _ptr(scores_ptr),
                                        C.c_void_p(stream)))

    def scores_raw(self, queries) -> np.ndarray:
        q = np.ascontiguousarray(queries, dtype=np.float32)
        out = np.empty((q.shape[0], self.num_docs), dtype=np.uint8)
        _check(lib().vs_int8_scores_raw(self._h, _ptr(q), C.c_int64(q.shape[0]), _ptr(out)))
        return out

    def set_profile(self, enable: bool = True) -> None:
        _check(lib().vs_int8_set_profile(self._h, C.c_int(int(enable))))

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib().vs_int8_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)


def int8_quantize(x, scale: float, device: int = 0) -> np.ndarray:
    """quantize_buffer_neon (QnnRunner.cpp:13-55) on the device."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.uint8)
    _check(lib().vs_int8_quantize(_ptr(a), C.c_int64(a.size), C.c_float(scale), _ptr(out), C.c_int(device)))
    return out
