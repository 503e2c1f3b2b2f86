"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/vsb200.h declares; the
entry points fail loudly (VS_ERR_CUDA) instead of falling back when there is no device."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vsb200.h")).read()
    return sorted(set(re.findall(r"VSB_API\s+[\w\s\*]+?\b(vs_\w+)\s*\(", src)))


def test_header_symbols_exported(vsb):
    names = declared_symbols()
    assert len(names) >= 15
    L = vsb.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/vsb200.h but not exported: {missing}"
    assert L.vs_abi_version() == 1


def test_no_cpu_fallback_and_argument_checks(vsb):
    L = vsb.lib()
    h = C.c_void_p()
    base = np.zeros((4, 128), dtype=np.float32)
    # argument validation happens before any device work
    assert L.vs_exact_create(C.byref(h), None, C.c_int64(4), 128, 0, C.c_int64(0)) == 1
    assert L.vs_exact_create(C.byref(h), base.ctypes.data_as(C.c_void_p), C.c_int64(4), 64, 0, C.c_int64(0)) == 5
    rc = L.vs_exact_create(C.byref(h), base.ctypes.data_as(C.c_void_p), C.c_int64(4), 128, 0, C.c_int64(0))
    err = L.vs_last_error()
    if vsb.device_count() == 0:
        assert rc == 2 and b"no CPU fallback" in err  # VS_ERR_CUDA, loudly
        assert not h
    else:
        assert rc == 0
        L.vs_exact_destroy(h)


def test_product_does_not_reference_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "hai-25-rag-on-edge_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), f"{f} mentions the oracle"
