"""CPU-side checks of the C ABI: the product library builds/loads, exports every symbol the header
declares, its ctypes mirror has the same struct layout, it fails loudly without a GPU, and the
product package never touches the oracle.  No compute calls that need a device."""
import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "lj_b200.h")
LIB = os.path.join(ROOT, "lj_gpu_b200", "liblj_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    from lj_gpu_b200 import _capi
    return _capi.load()


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^LJ_API\s+[\w\s\*]+?\b(lj_\w+)\s*\(", text, flags=re.M)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from lj_gpu_b200 import _capi
    names = declared_symbols()
    assert len(names) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lj_\w+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(_capi.PROTOTYPES), set(names) ^ set(_capi.PROTOTYPES)
    # nothing but the C ABI leaks out of the library (C++ symbols stay hidden)
    leaked = [s for s in re.findall(r" T (\S+)", out) if not s.startswith("lj_")]
    assert not leaked, leaked[:5]


def test_enum_values_match_the_header(tmp_path):
    """Every LJ_* enumerator of include/lj_b200.h has the same value in the ctypes mirror."""
    from lj_gpu_b200 import _capi
    text = open(HEADER).read()
    names = sorted(set(re.findall(r"^\s*(LJ_[A-Z0-9_]+)\s*=\s*\d+", text, flags=re.M)))
    assert {"LJ_LIST_TILES", "LJ_LIST_TILES_WIDE", "LJ_VARIANT_CELLTILE", "LJ_PREC_MIXED", "LJ_SOA_D"} <= set(names)
    src = tmp_path / "en.c"
    src.write_text('#include <stdio.h>\n#include "lj_b200.h"\nint main(void){'
                   + "".join('printf("%s %%d\\n", (int)%s);' % (n, n) for n in names) + 'return 0;}\n')
    exe = tmp_path / "en"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        assert hasattr(_capi, n), n
        assert int(got[n]) == getattr(_capi, n), n


def test_struct_layout_matches_the_header(tmp_path):
    from lj_gpu_b200 import _capi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lj_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(lj_force_args),'
                   'sizeof(lj_list_args), sizeof(lj_measure_args), sizeof(lj_buf),'
                   'offsetof(lj_force_args, ell_width), offsetof(lj_list_args, row_end),'
                   'offsetof(lj_measure_args, d2h_bytes));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(_capi.LjForceArgs), C.sizeof(_capi.LjListArgs), C.sizeof(_capi.LjMeasureArgs),
            C.sizeof(_capi.LjBuf), _capi.LjForceArgs.ell_width.offset, _capi.LjListArgs.row_end.offset,
            _capi.LjMeasureArgs.d2h_bytes.offset]
    assert got == want


def test_no_device_means_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lj_gpu_b200 import LJContext, LJError, _capi
    h = C.c_void_p()
    assert lib.lj_ctx_create(C.byref(h), 0) == _capi.LJ_ERR_NO_DEVICE
    assert h.value is None
    assert b"no CPU fallback" in lib.lj_status_string(_capi.LJ_ERR_NO_DEVICE)
    with pytest.raises(LJError):
        LJContext(0)


def test_host_generator_is_bit_exact(lib, oracle, golden):
    from lj_gpu_b200 import init_fcc, print_results
    for rho in (0.5, 1.0):
        q = init_fcc(rho, 50.0)
        assert hashlib.sha256(q.tobytes()).hexdigest() == str(golden(rho)["q_sha256"])
    assert np.array_equal(init_fcc(0.8, 21.7), oracle.init_fcc(0.8, 21.7))
    # capacity protocol of lj_init_fcc: -(needed) when the buffer is too small
    cells = C.c_int32(0)
    assert lib.lj_init_fcc(0.5, 50.0, None, 0, C.byref(cells)) == -62500 and cells.value == 25
    p = np.arange(36, dtype=np.float64).reshape(12, 3)
    lines = print_results(p).splitlines()
    assert len(lines) == 10 and lines[0] == "0.0000000000 1.0000000000 2.0000000000"
    assert lines[5] == "21.0000000000 22.0000000000 23.0000000000"


def test_product_never_reaches_into_the_oracle():
    pkg = os.path.join(ROOT, "lj_gpu_b200")
    for base, _, files in os.walk(pkg):
        if "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="replace").read()
                # comments may cite the oracle; code may not import, include, link or dlopen it
                bad = re.search(r"(from|import)\s+oracle|#include\s*[\"<][^\n]*oracle|liblj_oracle|"
                                r"CDLL\([^\n]*oracle|ljoracle", text)
                assert not bad, os.path.join(base, f)
    # and the library does not link against it
    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_oracle_header_declares_itself_test_infrastructure():
    for f in ("lj_oracle.c", "ref_harness.cpp", "ljoracle.py"):
        head = open(os.path.join(ROOT, "oracle", f)).read(1500)
        assert "TEST INFRASTRUCTURE ONLY" in head
    assert "Parity status: PINNED" in open(os.path.join(ROOT, "oracle", "lj_oracle.c")).read(2500)
