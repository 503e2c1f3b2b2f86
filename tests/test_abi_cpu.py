"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/vsb200.h declares; the
entry points fail loudly (VS_ERR_CUDA) instead of falling back when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

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


def test_header_is_plain_c_and_cxx(tmp_path):
    """include/vsb200.h is the drop-in boundary: it must compile on its own as C99 and as C++17 (no torch / CUDA types)."""
    import subprocess

    inc = os.path.join(ROOT, "include")
    c = tmp_path / "t.c"
    c.write_text('#include "vsb200.h"\nint main(void) { vs_exact_t* h = 0; (void)h; return VS_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-c", str(c), "-o", str(tmp_path / "t_c.o")],
                   check=True)
    cc = tmp_path / "t.cpp"
    cc.write_text('#include "vsb200.h"\nint main() { vs_ivf_t* h = nullptr; (void)h; return VS_OK; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-c", str(cc), "-o", str(tmp_path / "t_cc.o")],
                   check=True)


def test_host_programs_fail_loudly_without_a_gpu(tmp_path, vsb):
    """The cpu_baseline-compatible executable has no CPU fallback: on a machine without a CUDA device it must exit
    non-zero with a message instead of computing anything (skipped where a device is visible)."""
    import subprocess

    exe = os.path.join(ROOT, "hai-25-rag-on-edge_b200", "bin", "cpu_baseline")
    if not os.path.exists(exe):
        pytest.skip("host programs not built")
    if vsb.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    base, qry = tmp_path / "b.fvecs", tmp_path / "q.fvecs"
    vsb.synth.write_fvecs(str(base), vsb.synth.make("sift", 1, 64))
    vsb.synth.write_fvecs(str(qry), vsb.synth.make("sift", 2, 4))
    r = subprocess.run([exe, str(base), str(qry), "5", str(tmp_path / "out.txt")], capture_output=True, text=True)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout) or "device" in (r.stderr + r.stdout).lower()
    assert not os.path.exists(tmp_path / "out.txt") or os.path.getsize(tmp_path / "out.txt") == 0
