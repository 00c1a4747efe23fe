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


def test_parse_mgf_numbers_equal_python_float(lib):
    """Every accepted number text must become the correctly rounded double (strtod / Double.parseDouble): the fast path
    (integer mantissa times an exact power of ten) and the strtod fallback, against Python's float()."""
    import random
    from pepquery_b200 import parse_mgf
    rng = random.Random(5)
    toks = ["0", "0.0", "5.", ".5", "123.4567", "1234567.89", "0.000001", "1e5", "1E-5", "2.5e+3", "9007199254740993", "9007199254740992",
            "123456789012345678901234567890", "0.1234567890123456789012345", "1e22", "1e23", "1e-22", "1e-23", "4.35", "1.7976931348623157e308",
            "4.9e-324", "100.", "00012.5000", "1234.5678e-2", "3.14159265358979323846"]
    for _ in range(400):
        k = rng.randint(1, 17)
        digits = "".join(rng.choice("0123456789") for _ in range(k))
        cut = rng.randint(0, k)
        t = (digits[:cut] or "0") + "." + digits[cut:]
        if rng.random() < 0.3:
            t += "e%+d" % rng.randint(-30, 30)
        toks.append(t)
    text = "BEGIN IONS\nPEPMASS=%s\nCHARGE=2+\n" % toks[4] + "".join("%s %s\n" % (a, b) for a, b in zip(toks, reversed(toks))) + "END IONS\n"
    sp = parse_mgf(text)
    assert len(sp["mz"]) == len(toks)
    want_a = np.array([float(t) for t in toks]); want_b = want_a[::-1]
    assert np.array_equal(sp["mz"], want_a) and np.array_equal(sp["inten"], want_b)
    assert sp["prec_mz"][0] == float(toks[4])


def test_parse_mgf_threads_agree(lib, monkeypatch):
    """The parallel parse (chunks cut at BEGIN IONS lines) must equal the one-thread parse, also with blank-prefixed and
    CRLF lines, comment / header lines between spectra, a spectrum without peaks and intensity-less peak lines."""
    import random
    from pepquery_b200 import parse_mgf
    rng = random.Random(9)
    parts = ["# produced by a test\nCOM=x\n"]
    for i in range(30000):
        eol = "\r\n" if i % 7 == 0 else "\n"
        ind = "  " if i % 5 == 0 else ""
        n = 0 if i % 97 == 0 else rng.randint(1, 40)
        parts.append("%sBEGIN IONS%sTITLE=s%d%sPEPMASS=%.5f 1234.5%sCHARGE=%d+%s" % (ind, eol, i, eol, 300 + rng.random() * 900, eol, rng.randint(1, 4), eol))
        for k in range(n):
            parts.append("%s%.4f%s%s" % (ind, 100 + rng.random() * 1800, "" if k % 11 == 0 else " %.2f" % (rng.random() * 1e5), eol))
        parts.append("END IONS%s%s" % (eol, "\n" if i % 3 == 0 else ""))
    text = "".join(parts)
    assert len(text) > (1 << 20)
    monkeypatch.setenv("PQ_MGF_THREADS", "1")
    one = parse_mgf(text)
    for th in ("2", "7", "16"):
        monkeypatch.setenv("PQ_MGF_THREADS", th)
        many = parse_mgf(text)
        for k in one:
            assert np.array_equal(one[k], many[k]), (th, k)
    assert len(one["prec_mz"]) == 30000 and int(one["peak_off"][-1]) == len(one["mz"])
