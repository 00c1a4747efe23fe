"""ctypes binding of the CPU oracle (oracle/pq_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  The product package (pepquery_b200/) never imports this module.
Parity unpinned at the compomics boundary — see the header of pq_oracle.cpp.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from pepquery_b200 import _abi  # noqa: E402  (POD struct mirrors only)

_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return os.path.join(_HERE, "libpq_oracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpq_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        for name in ("orc_log10", "orc_log", "orc_log10_factorial", "orc_factorial_log", "orc_residue_mass", "orc_proton"):
            getattr(L, name).restype = C.c_double
        L.orc_log10.argtypes = [C.c_double]
        L.orc_log.argtypes = [C.c_double]
        L.orc_residue_mass.argtypes = [C.c_char]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def find_in_proteome(toff, tres, poff, pres):
    """Step-1 novelty filter (FM-index semantics): target offsets u32[n+1] + residues, protein offsets u64[n+1] + residues."""
    toff = np.ascontiguousarray(toff, dtype=np.uint32); poff = np.ascontiguousarray(poff, dtype=np.uint64)
    tres = np.ascontiguousarray(tres, dtype=np.uint8); pres = np.ascontiguousarray(pres, dtype=np.uint8)
    out = np.zeros(max(1, len(toff) - 1), dtype=np.int32)
    lib().orc_find_in_proteome(C.c_int32(len(toff) - 1), _p(toff), _p(tres), C.c_int32(len(poff) - 1), _p(poff), _p(pres), _p(out))
    return out[:len(toff) - 1]


def java_random(seed, bound, n):
    out = np.zeros(n, dtype=np.int32)
    lib().orc_java_random_ints(C.c_int64(seed), C.c_int32(bound), C.c_int32(n), _p(out))
    return out


def shuffles(seq, num=1000, seed=12457):
    L = len(seq) + 1
    cap = num * L + 16
    buf = C.create_string_buffer(cap)
    n = C.c_int32()
    rc = lib().orc_shuffles(seq.encode(), C.c_int32(num), C.c_int64(seed), buf, C.c_int64(cap), C.byref(n))
    assert rc == 0
    return [buf.raw[i * L:i * L + L - 1].decode() for i in range(n.value)]


class Oracle:
    """Same call sequence as pepquery_b200.host.Search, executed by the CPU restatement."""

    def __init__(self, params, threads=1):
        self.L = lib()
        self.params = params
        self.h = C.c_void_p()
        rc = self.L.orc_create(C.byref(params), C.c_int32(threads), C.byref(self.h))
        assert rc == 0

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def load_spectra(self, sp):
        self._sp = sp
        return self.L.orc_load_spectra(self.h, C.c_int64(len(sp["prec_mz"])), _p(sp["prec_mz"]), _p(sp["charge"]),
                                       _p(sp["peak_off"]), _p(sp["mz"]), _p(sp["inten"]))

    def set_targets(self, off, res):
        return self.L.orc_set_targets(self.h, C.c_int32(len(off) - 1), _p(off), _p(res))

    def _forms(self, fn):
        n = C.c_int64()
        fn(self.h, None, C.c_int64(0), C.byref(n))
        arr = (_abi.pq_form * max(1, n.value))()
        rc = fn(self.h, arr, C.c_int64(n.value), C.byref(n))
        assert rc == 0
        return arr, n.value

    def target_forms(self):
        return self._forms(self.L.orc_get_target_forms)

    def reference_forms(self):
        return self._forms(self.L.orc_get_reference_forms)

    def reference_sequences(self):
        ns, nb = C.c_int64(), C.c_int64()
        self.L.orc_get_reference_sequences(self.h, None, C.c_int64(0), C.byref(ns), C.byref(nb))
        buf = C.create_string_buffer(max(1, nb.value))
        self.L.orc_get_reference_sequences(self.h, buf, C.c_int64(nb.value), C.byref(ns), C.byref(nb))
        return [s.decode() for s in buf.raw[:nb.value].split(b"\0")[:ns.value]]

    def build_reference(self, off, res):
        return self.L.orc_build_reference(self.h, C.c_int32(len(off) - 1), _p(off), _p(res))

    def step2(self):
        n = C.c_int64()
        self.L.orc_step2_match_targets(self.h, C.byref(n))
        return n.value

    def step3(self):
        return self.L.orc_step3_competitors(self.h)

    def step4(self):
        return self.L.orc_step4_shuffles(self.h)

    def rank(self):
        return self.L.orc_rank_and_validate(self.h)

    def step5(self, ptms):
        arr = (_abi.pq_ptm * max(1, len(ptms)))(*ptms)
        return self.L.orc_step5_ptm(self.h, C.c_int32(len(ptms)), arr)

    def psms(self):
        n = C.c_int64()
        self.L.orc_get_psms(self.h, None, C.c_int64(0), C.byref(n))
        arr = (_abi.pq_psm * max(1, n.value))()
        self.L.orc_get_psms(self.h, arr, C.c_int64(n.value), C.byref(n))
        return np.frombuffer(arr, dtype=psm_dtype(), count=n.value).copy()

    def competitors(self, step):
        n = C.c_int64()
        self.L.orc_get_competitors(self.h, C.c_int32(step), None, C.c_int64(0), C.byref(n))
        arr = (_abi.pq_competitor * max(1, n.value))()
        self.L.orc_get_competitors(self.h, C.c_int32(step), arr, C.c_int64(n.value), C.byref(n))
        return np.frombuffer(arr, dtype=competitor_dtype(), count=n.value).copy()

    def delta_scores(self):
        """(ref_delta_score, mod_delta_score) per PSM, psms() order."""
        n = C.c_int64()
        self.L.orc_get_delta_scores(self.h, None, None, C.c_int64(0), C.byref(n))
        a = np.zeros(max(1, n.value)); b = np.zeros(max(1, n.value))
        self.L.orc_get_delta_scores(self.h, _p(a), _p(b), C.c_int64(n.value), C.byref(n))
        return a[:n.value], b[:n.value]

    def set_cost_model(self, hoisted):
        """False: as written (spectrum pre-processing redone per pair, like scorePeptide2Spectrum); True: hoisted (once per spectrum)."""
        self.L.orc_set_cost_model(self.h, C.c_int32(1 if hoisted else 0))

    def set_reference_window(self, lo, hi):
        """Forces the (open) mass range of the reference table instead of taking it from the matched spectra."""
        self.L.orc_set_reference_window(self.h, C.c_int32(1), C.c_double(lo), C.c_double(hi))

    def set_targeted_psms(self, target_seq, spectrum):
        """SpectraInput.pep2targeted_spectra: (target sequence index, spectrum index) pairs; call after set_targets."""
        a = np.ascontiguousarray(target_seq, dtype=np.int32); b = np.ascontiguousarray(spectrum, dtype=np.int32)
        rc = self.L.orc_set_targeted_psms(self.h, C.c_int64(len(a)), _p(a), _p(b))
        assert rc == 0

    def targeted_status(self):
        n = C.c_int64()
        self.L.orc_get_targeted_status(self.h, None, C.c_int64(0), C.byref(n))
        out = np.zeros(max(1, n.value), dtype=np.int32)
        self.L.orc_get_targeted_status(self.h, _p(out), C.c_int64(n.value), C.byref(n))
        return out[:n.value]

    def pairs(self):
        out = np.zeros(4, dtype=np.uint64)
        self.L.orc_get_pairs(self.h, _p(out))
        return out

    def score_pairs(self, spectrum_index, forms, n=None):
        n = len(spectrum_index) if n is None else n
        si = np.ascontiguousarray(spectrum_index, dtype=np.int32)
        score = np.zeros(n, dtype=np.float64)
        ca = np.zeros(n, dtype=np.int32)
        cb = np.zeros(n, dtype=np.int32)
        self.L.orc_score_pairs(self.h, C.c_int64(n), _p(si), forms, _p(score), _p(ca), _p(cb))
        return score, ca, cb

    def shuffle_form(self, target_form, seq):
        f = _abi.pq_form()
        self.L.orc_shuffle_form(self.h, C.c_int32(target_form), seq.encode(), C.byref(f))
        return f

    def preprocess(self, si, method):
        cap = 4096
        mz = np.zeros(cap)
        val = np.zeros(cap)
        n = C.c_int32()
        lim = C.c_double()
        rc = self.L.orc_preprocess(self.h, C.c_int32(si), C.c_int32(method), _p(mz), _p(val), C.c_int32(cap), C.byref(n), C.byref(lim))
        assert rc == 0
        k = max(0, n.value)
        return mz[:k].copy(), val[:k].copy(), n.value, lim.value


def psm_dtype():
    return np.dtype([(n, _np_type(t)) for n, t in _abi.pq_psm._fields_])


def competitor_dtype():
    return np.dtype([(n, _np_type(t)) for n, t in _abi.pq_competitor._fields_])


def _np_type(t):
    return {C.c_int32: np.int32, C.c_double: np.float64, C.c_int64: np.int64, C.c_uint64: np.uint64}[t]
