"""Seeded synthetic SIFT-shaped data (the INRIA corpus is not available offline).

Counter-based: element (row, col) of a set is a pure function of (seed, row, col), so any slice can be
regenerated on the CPU (here, numpy) or on the GPU (csrc/synth.cu implements the same integer arithmetic,
bit for bit) without storing the set.

Laws
  "sift"  integer components in [0,127], skewed to small values like SIFT histograms (row norm ~ 480).
          Every fp32/TF32/u8 path is exact on it (sums < 2^24), so it checks indexing/top-k/sharding
          logic bit-exactly but cannot check fp32 faithfulness.
  "cont"  the same plus a 16-bit uniform fraction in [-0.5, 0.5): continuous-valued, exercises the
          3xTF32 split and the 1e-5 tolerance.
  "mix"   IVF mixture: 4096 latent centres drawn with law "sift" (shared `centre_seed`), point = clip(centre +
          ~N(0, 12^2) integer noise, 0, 218); gives nlist=1024 / nprobe 8..32 a meaningful recall band.
"""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
N_CENTRES = 4096
LAWS = {"sift": 0, "cont": 1, "mix": 2}


def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _hash(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return _mix64(idx.astype(np.uint64) + np.uint64(seed) * _GOLD)


def _sift_from_hash(h: np.ndarray) -> np.ndarray:
    b0 = (h & np.uint64(0xFF)).astype(np.int64)
    b1 = ((h >> np.uint64(8)) & np.uint64(0xFF)).astype(np.int64)
    return (b0 * b1) >> 9


def rows(law: str, seed: int, row0: int, nrows: int, dim: int = 128, centre_seed: int = 7) -> np.ndarray:
    """float32 [nrows, dim] = rows row0 .. row0+nrows-1 of the set (law, seed)."""
    r = np.arange(row0, row0 + nrows, dtype=np.uint64)[:, None]
    c = np.arange(dim, dtype=np.uint64)[None, :]
    h = _hash(seed, r * np.uint64(dim) + c)
    if law == "sift":
        return _sift_from_hash(h).astype(np.float32)
    if law == "cont":
        v = _sift_from_hash(h).astype(np.float32)
        frac = ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.float32) / np.float32(65536.0)
        return (v + (frac - np.float32(0.5))).astype(np.float32)
    if law == "mix":
        cid = _hash(seed ^ 0x5BD1E995, r) % np.uint64(N_CENTRES)  # [nrows,1]
        hc = _hash(centre_seed, cid * np.uint64(dim) + c)
        centre = _sift_from_hash(hc)
        s = np.zeros(h.shape, dtype=np.int64)
        for sh in (16, 24, 32, 40):
            s += ((h >> np.uint64(sh)) & np.uint64(0xFF)).astype(np.int64)
        noise = s // 12 - 42
        return np.clip(centre + noise, 0, 218).astype(np.float32)
    raise ValueError(f"unknown law {law!r}")


def make(law: str, seed: int, nrows: int, dim: int = 128, centre_seed: int = 7, chunk: int = 1 << 16) -> np.ndarray:
    out = np.empty((nrows, dim), dtype=np.float32)
    for r0 in range(0, nrows, chunk):
        n = min(chunk, nrows - r0)
        out[r0 : r0 + n] = rows(law, seed, r0, n, dim, centre_seed)
    return out


def write_fvecs(path: str, a: np.ndarray) -> None:
    """[int32 d][d x float32] records (cpu_baseline.cpp:31-58 reads exactly this)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    n, d = a.shape
    out = np.empty((n, d + 1), dtype=np.float32)
    out[:, 0] = np.array([d], dtype=np.int32).view(np.float32)[0]
    out[:, 1:] = a
    out.tofile(path)


def read_fvecs(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.int32)
    if raw.size == 0:
        return np.zeros((0, 0), dtype=np.float32)
    d = int(raw[0])
    if raw.size % (d + 1) != 0:
        raise ValueError("File seems truncated.")
    m = raw.reshape(-1, d + 1)
    if not (m[:, 0] == d).all():
        raise ValueError("Inconsistent dimension.")
    return m[:, 1:].copy().view(np.float32)


def write_ivecs(path: str, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a, dtype=np.int32)
    n, d = a.shape
    out = np.empty((n, d + 1), dtype=np.int32)
    out[:, 0] = d
    out[:, 1:] = a
    out.tofile(path)
