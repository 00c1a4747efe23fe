"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/pepquery_b200.h declares, the
ctypes mirrors match the C struct layouts, host-only entry points work, and there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pepquery_b200.h")


@pytest.fixture(scope="module")
def lib():
    from pepquery_b200 import host
    if not os.path.exists(host.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    return host.load_library()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pq_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from pepquery_b200 import host
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libpepquery_b200.so does not export %s" % n
    assert sorted(host.EXPORTS) == names


def test_struct_layouts_match_the_header(lib, tmp_path):
    from pepquery_b200 import _abi
    structs = ["pq_mod", "pq_params", "pq_form", "pq_psm", "pq_competitor", "pq_ptm", "pq_counters"]
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){%s printf("%%zu %%zu %%zu\\n", offsetof(pq_params, var), '
                   'offsetof(pq_psm, p_value), offsetof(pq_form, seq)); return 0;}\n' %
                   (HEADER, "".join('printf("%%zu\\n", sizeof(%s));' % s for s in structs)))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)]).decode().split()
    for s, sz in zip(structs, out):
        assert C.sizeof(getattr(_abi, s)) == int(sz), s
    assert [int(x) for x in out[len(structs):]] == [_abi.pq_params.var.offset, _abi.pq_psm.p_value.offset, _abi.pq_form.seq.offset]


def test_default_params_match(lib):
    from pepquery_b200 import _abi
    p = _abi.pq_params()
    lib.pq_default_params(C.byref(p))
    q = _abi.default_params()
    assert bytes(p) == bytes(q)
    assert lib.pq_version() == 1


def test_no_cpu_fallback(lib):
    """Without a CUDA device pq_create must fail with PQ_ERR_NO_DEVICE (and say so), never run on the host."""
    from conftest import have_gpu
    from pepquery_b200 import PepQueryError, Search, _abi
    if have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(PepQueryError) as e:
        Search()
    assert e.value.code == _abi.PQ_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_bad_arguments_are_errors_not_exits(lib):
    from pepquery_b200 import _abi
    h = C.c_void_p()
    assert lib.pq_create(None, C.byref(h)) == _abi.PQ_ERR_BAD_ARG
    assert b"null" in lib.pq_last_error(None)
    assert lib.pq_get_counters(None, None) == _abi.PQ_ERR_BAD_ARG
    assert lib.pq_destroy(None) == _abi.PQ_OK


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pepquery_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pepquery_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    s = line.strip()
                    if s.startswith(("//", "#", "*", '"""')) and "include" not in s and "import" not in s:
                        continue
                    if re.search(r"(import|from|include|dlopen|CDLL).*\boracle\b", s) or "pq_oracle" in s and "import" in s:
                        bad.append((f, s))
    assert not bad, bad


MGF = """BEGIN IONS
TITLE=1
PEPMASS=500.2505 12345.0
CHARGE=2+
100.1000 10.00
200.2000 20.50
300.3000 5.25
END IONS

BEGIN IONS
TITLE=second spectrum
PEPMASS=700.33
CHARGE=3+
150.0 1
END IONS
"""


def test_parse_mgf(lib):
    from pepquery_b200 import parse_mgf
    sp = parse_mgf(MGF)
    assert sp["prec_mz"].tolist() == [500.2505, 700.33] and sp["charge"].tolist() == [2, 3]
    assert sp["peak_off"].tolist() == [0, 3, 4]
    assert sp["mz"].tolist() == [100.1, 200.2, 300.3, 150.0] and sp["inten"].tolist() == [10.0, 20.5, 5.25, 1.0]
    assert len(parse_mgf("")["prec_mz"]) == 0
    # capacity is checked (two-call pattern)
    raw = MGF.encode()
    ns, npk = C.c_int64(1), C.c_uint64(1)
    z = np.zeros(8)
    zi = np.zeros(8, dtype=np.int32)
    zo = np.zeros(8, dtype=np.uint64)
    from pepquery_b200 import _abi
    rc = lib.pq_parse_mgf(raw, C.c_uint64(len(raw)), C.byref(ns), C.byref(npk), z.ctypes.data_as(C.c_void_p), zi.ctypes.data_as(C.c_void_p),
                          zo.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p))
    assert rc == _abi.PQ_ERR_CAPACITY
