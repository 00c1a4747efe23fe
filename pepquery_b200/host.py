"""Host-side mirror of the reference's entry points for the pair-scoring path, on top of the C ABI.

Reference (Java) call site                                   -> here
  pg/CParameter.java statics                                 -> CParameter (builds a pq_params)
  SpectraInput.readMSMSfast        (pg/SpectraInput.java:91)  -> Search.readMSMSfast(spectra)   [load + Step 2]
  DatabaseInput.db_search          (pg/DatabaseInput.java:672)-> Search.db_search(proteins)     [index + Step 3]
  StatisticScoreCalc.run           (pg/StatisticScoreCalc.java:19) -> Search.statistic_score()  [Step 4]
  RankPSM.rankPSMs + passed_psm_validation                   -> Search.rank_psms()
  ModificationDB.doPTMValidation   (ModificationDB.java:1320) -> Search.do_ptm_validation(ptms) [Step 5]
  PeptideSearchMT.scorePeptide2Spectrum (PSMT:1144)           -> scorePeptide2Spectrum / Search.score_pairs

Error behaviour: the reference calls System.exit on bad input; here every failure raises PepQueryError carrying the
pq_status code and pq_last_error() text.  A missing CUDA library or GPU is an error, never a fallback.
"""
import ctypes as C
import os

import numpy as np

from . import _abi

_ROOT = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class PepQueryError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pepquery_b200 error %d: %s" % (code, msg))
        self.code = code


def library_path():
    return os.path.join(_ROOT, "libpepquery_b200.so")


def load_library():
    """Loads libpepquery_b200.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise PepQueryError(_abi.PQ_ERR_NO_DEVICE, "CUDA library %s is not built; there is no CPU fallback "
                            "(run `python -c 'import __graft_entry__ as g; g.build()'`)" % path)
    L = C.CDLL(path)
    L.pq_last_error.restype = C.c_char_p
    L.pq_last_error.argtypes = [C.c_void_p]
    L.pq_version.restype = C.c_int32
    for name in EXPORTS:
        if name not in ("pq_last_error", "pq_version", "pq_default_params"):
            getattr(L, name).restype = C.c_int32
    _LIB = L
    return L


# every symbol include/pepquery_b200.h declares
EXPORTS = ["pq_version", "pq_last_error", "pq_default_params", "pq_create", "pq_destroy", "pq_load_spectra", "pq_load_mgf", "pq_parse_mgf_device",
           "pq_set_targets", "pq_set_targeted_psms", "pq_get_targeted_psm_status", "pq_get_target_forms", "pq_build_reference", "pq_stage_reference", "pq_get_reference_forms",
           "pq_step2_match_targets", "pq_step3_competitors", "pq_step4_shuffles", "pq_rank_and_validate",
           "pq_step5_ptm", "pq_get_psms", "pq_get_delta_scores", "pq_export_psms_device", "pq_get_competitors", "pq_get_shuffles", "pq_score_pairs",
           "pq_find_in_proteome", "pq_get_counters", "pq_reset_counters", "pq_event_record", "pq_event_elapsed_ms", "pq_parse_mgf",
           "pq_read_ms_library", "pq_selftest"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def CParameter(**kw):
    """pq_params with the defaults of pg/CParameter.java:84-463 (and -n 1000); keyword overrides by field name."""
    return _abi.default_params(**kw)


def psm_dtype():
    m = {C.c_int32: np.int32, C.c_double: np.float64}
    return np.dtype([(n, m[t]) for n, t in _abi.pq_psm._fields_])


def competitor_dtype():
    m = {C.c_int32: np.int32, C.c_double: np.float64}
    return np.dtype([(n, m[t]) for n, t in _abi.pq_competitor._fields_])


class Search:
    """One pq_ctx = one search on one GPU."""

    def __init__(self, params=None, device=None):
        self.L = load_library()
        self.params = params if params is not None else CParameter()
        if device is not None:
            self.params.device = int(device)
        self.h = C.c_void_p()
        rc = self.L.pq_create(C.byref(self.params), C.byref(self.h))
        if rc != 0:
            raise PepQueryError(rc, (self.L.pq_last_error(None) or b"").decode())

    # -- plumbing ----------------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise PepQueryError(rc, (self.L.pq_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.pq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- raw ABI calls -------------------------------------------------------------------------------------------
    def load_spectra(self, sp):
        n = len(sp["prec_mz"])
        self._ck(self.L.pq_load_spectra(self.h, C.c_int64(n), _p(sp["prec_mz"]), _p(sp["charge"]), _p(sp["peak_off"]),
                                        _p(sp["mz"]), _p(sp["inten"])))

    def load_mgf(self, text):
        """MGF text (bytes / uint8 array, e.g. in pinned memory) parsed on the device and loaded; -> number of spectra."""
        n = C.c_int64()
        if isinstance(text, (bytes, bytearray)):
            self._ck(self.L.pq_load_mgf(self.h, text, C.c_uint64(len(text)), C.byref(n)))
        else:
            self._ck(self.L.pq_load_mgf(self.h, _p(text), C.c_uint64(text.nbytes), C.byref(n)))
        return n.value

    def parse_mgf_device(self, text):
        """pq_parse_mgf_device: MGF text parsed on the GPU -> CSR dict on the host."""
        raw = text if isinstance(text, (bytes, bytearray)) else None
        ptr = raw if raw is not None else _p(text)
        nb = len(raw) if raw is not None else text.nbytes
        ns, npk = C.c_int64(0), C.c_uint64(0)
        self._ck(self.L.pq_parse_mgf_device(self.h, ptr, C.c_uint64(nb), C.byref(ns), C.byref(npk), None, None, None, None, None))
        sp = dict(prec_mz=np.zeros(ns.value), charge=np.zeros(ns.value, dtype=np.int32), peak_off=np.zeros(ns.value + 1, dtype=np.uint64),
                  mz=np.zeros(npk.value), inten=np.zeros(npk.value))
        self._ck(self.L.pq_parse_mgf_device(self.h, ptr, C.c_uint64(nb), C.byref(ns), C.byref(npk), _p(sp["prec_mz"]), _p(sp["charge"]), _p(sp["peak_off"]),
                                            _p(sp["mz"]), _p(sp["inten"])))
        return sp

    def set_targets(self, off, res):
        self._ck(self.L.pq_set_targets(self.h, C.c_int32(len(off) - 1), _p(off), _p(res)))

    def set_targeted_psms(self, target_seq, spectrum):
        """SpectraInput.pep2targeted_spectra: (target sequence index, spectrum index) pairs; after set_targets, before step2."""
        a = np.ascontiguousarray(target_seq, dtype=np.int32); b = np.ascontiguousarray(spectrum, dtype=np.int32)
        self._ck(self.L.pq_set_targeted_psms(self.h, C.c_int64(len(a)), _p(a), _p(b)))

    def targeted_psm_status(self):
        """One pq_psm_type per targeted pair (PeptideSearchMT.export_targeted_psm_status)."""
        n = C.c_int64()
        self._ck(self.L.pq_get_targeted_psm_status(self.h, None, C.c_int64(0), C.byref(n)))
        out = np.zeros(max(1, n.value), dtype=np.int32)
        self._ck(self.L.pq_get_targeted_psm_status(self.h, _p(out), C.c_int64(n.value), C.byref(n)))
        return out[:n.value]

    def _forms(self, fn):
        n = C.c_int64()
        self._ck(fn(self.h, None, C.c_int64(0), C.byref(n)))
        arr = (_abi.pq_form * max(1, n.value))()
        self._ck(fn(self.h, arr, C.c_int64(n.value), C.byref(n)))
        return arr, n.value

    def target_forms(self):
        return self._forms(self.L.pq_get_target_forms)

    def reference_forms(self):
        return self._forms(self.L.pq_get_reference_forms)

    def build_reference(self, off=None, res=None):
        """off/res = None: the database handed over earlier by stage_reference()."""
        if off is None:
            self._ck(self.L.pq_build_reference(self.h, C.c_int32(0), None, None))
        else:
            self._ck(self.L.pq_build_reference(self.h, C.c_int32(len(off) - 1), _p(off), _p(res)))

    def stage_reference(self, off, res):
        self._ck(self.L.pq_stage_reference(self.h, C.c_int32(len(off) - 1), _p(off), _p(res)))

    def step2(self):
        n = C.c_int64()
        self._ck(self.L.pq_step2_match_targets(self.h, C.byref(n)))
        return n.value

    def step3(self):
        self._ck(self.L.pq_step3_competitors(self.h))

    def step4(self):
        self._ck(self.L.pq_step4_shuffles(self.h))

    def rank(self):
        self._ck(self.L.pq_rank_and_validate(self.h))

    def step5(self, ptms):
        arr = (_abi.pq_ptm * max(1, len(ptms)))(*ptms)
        self._ck(self.L.pq_step5_ptm(self.h, C.c_int32(len(ptms)), arr))

    def psms(self, out=None):
        """PSM records in the caller's spectrum order.  `out`: optional uint8 buffer (e.g. pinned memory) to fill."""
        n = C.c_int64()
        self._ck(self.L.pq_get_psms(self.h, None, C.c_int64(0), C.byref(n)))
        dt = psm_dtype()
        if out is None:
            arr = np.empty(max(1, n.value), dtype=dt)
        else:
            arr = np.frombuffer(out, dtype=dt, count=max(1, n.value))
        self._ck(self.L.pq_get_psms(self.h, _p(arr), C.c_int64(n.value), C.byref(n)))
        return arr[:n.value]

    def psms_device(self):
        """(device address, count) of the exported PSM records, for a device-side gather (NCCL)."""
        addr = C.c_uint64(); n = C.c_int64()
        self._ck(self.L.pq_export_psms_device(self.h, C.byref(addr), C.byref(n)))
        return addr.value, n.value

    def psm_count(self):
        n = C.c_int64()
        self._ck(self.L.pq_get_psms(self.h, None, C.c_int64(0), C.byref(n)))
        return n.value

    def find_in_proteome(self, toff, tres, poff, pres):
        """Step-1 novelty filter: 1 per target sequence that occurs in a protein (I = L), else 0."""
        toff = np.ascontiguousarray(toff, dtype=np.uint32); poff = np.ascontiguousarray(poff, dtype=np.uint64)
        tres = np.ascontiguousarray(tres, dtype=np.uint8); pres = np.ascontiguousarray(pres, dtype=np.uint8)
        out = np.zeros(max(1, len(toff) - 1), dtype=np.int32)
        self._ck(self.L.pq_find_in_proteome(self.h, C.c_int32(len(toff) - 1), _p(toff), _p(tres), C.c_int32(len(poff) - 1), _p(poff), _p(pres), _p(out)))
        return out[:len(toff) - 1]

    def delta_scores(self):
        """(ref_delta_score, mod_delta_score) per PSM in psms() order (PeptideSearchMT.add_delta_score)."""
        n = C.c_int64()
        self._ck(self.L.pq_get_delta_scores(self.h, None, None, C.c_int64(0), C.byref(n)))
        a = np.zeros(max(1, n.value)); b = np.zeros(max(1, n.value))
        self._ck(self.L.pq_get_delta_scores(self.h, _p(a), _p(b), C.c_int64(n.value), C.byref(n)))
        return a[:n.value], b[:n.value]

    def competitors(self, step=3, out=None):
        """detail.txt (step 3) / ptm.txt (step 5) rows.  `out`: optional uint8 buffer (e.g. pinned memory) to fill."""
        n = C.c_int64()
        self._ck(self.L.pq_get_competitors(self.h, C.c_int32(step), None, C.c_int64(0), C.byref(n)))
        dt = competitor_dtype()
        if out is None or n.value * dt.itemsize > out.nbytes:
            arr = np.empty(max(1, n.value), dtype=dt)
        else:
            arr = np.frombuffer(out, dtype=dt, count=max(1, n.value))
        self._ck(self.L.pq_get_competitors(self.h, C.c_int32(step), _p(arr), C.c_int64(n.value), C.byref(n)))
        return arr[:n.value]

    def shuffles(self, target_seq, length):
        n = C.c_int32()
        self._ck(self.L.pq_get_shuffles(self.h, C.c_int32(target_seq), None, C.c_int64(0), C.byref(n)))
        cap = max(1, n.value) * (length + 1)
        buf = C.create_string_buffer(cap)
        self._ck(self.L.pq_get_shuffles(self.h, C.c_int32(target_seq), buf, C.c_int64(cap), C.byref(n)))
        L1 = length + 1
        return [buf.raw[i * L1:i * L1 + length].decode() for i in range(n.value)]

    def score_pairs(self, spectrum_index, forms, n=None):
        n = len(spectrum_index) if n is None else n
        si = np.ascontiguousarray(spectrum_index, dtype=np.int32)
        score = np.zeros(n, dtype=np.float64)
        ca = np.zeros(n, dtype=np.int32)
        cb = np.zeros(n, dtype=np.int32)
        self._ck(self.L.pq_score_pairs(self.h, C.c_int64(n), _p(si), forms, _p(score), _p(ca), _p(cb)))
        return score, ca, cb

    def counters(self):
        c = _abi.pq_counters()
        self._ck(self.L.pq_get_counters(self.h, C.byref(c)))
        return c

    def reset_counters(self):
        self._ck(self.L.pq_reset_counters(self.h))

    def event_record(self, which):
        self._ck(self.L.pq_event_record(self.h, C.c_int32(which)))

    def event_elapsed_ms(self):
        ms = C.c_double()
        self._ck(self.L.pq_event_elapsed_ms(self.h, C.byref(ms)))
        return ms.value

    def selftest(self, which, n, seed=1):
        """Device self-test of an arithmetic shortcut (pq_selftest): number of mismatches, must be 0."""
        bad = C.c_uint64()
        self._ck(self.L.pq_selftest(self.h, C.c_int32(which), C.c_uint64(n), C.c_uint64(seed), C.byref(bad)))
        return bad.value

    # -- the reference's names -------------------------------------------------------------------------------------
    def readMSMSfast(self, spectra, target_off, target_res):
        """SpectraInput.readMSMSfast (pg/SpectraInput.java:91-279): load spectra, score targets in window."""
        self.load_spectra(spectra)
        self.set_targets(target_off, target_res)
        return self.step2()

    def db_search(self, prot_off, prot_res):
        """DatabaseInput.db_search (pg/DatabaseInput.java:672-758)."""
        self.build_reference(prot_off, prot_res)
        self.step3()

    def statistic_score(self):
        """StatisticScoreCalc.run (pg/StatisticScoreCalc.java:19-63)."""
        self.step4()

    def rank_psms(self, out=None):
        """RankPSM.rankPSMs (pg/RankPSM.java:36) + CParameter.passed_psm_validation (pg/CParameter.java:913)."""
        self.rank()
        return self.psms(out)

    def do_ptm_validation(self, ptms):
        """ModificationDB.doPTMValidation (OpenModificationSearch/ModificationDB.java:1320-1466)."""
        self.step5(ptms)
        return self.psms()


def scorePeptide2Spectrum(search, form, spectrum_index):
    """PeptideSearchMT.scorePeptide2Spectrum (pg/PeptideSearchMT.java:1144-1251) for one pair -> (score, a, b)."""
    arr = (_abi.pq_form * 1)(form)
    s, a, b = search.score_pairs([spectrum_index], arr, 1)
    return float(s[0]), int(a[0]), int(b[0])


def parse_mgf(text):
    """MGF text -> CSR dict (pq_parse_mgf; replaces compomics MgfFileIterator, SpectraInput.java:150)."""
    L = load_library()
    raw = text if isinstance(text, (bytes, bytearray)) else text.encode()
    ns = C.c_int64(0)
    npk = C.c_uint64(0)
    rc = L.pq_parse_mgf(raw, C.c_uint64(len(raw)), C.byref(ns), C.byref(npk), None, None, None, None, None)
    if rc != 0:
        raise PepQueryError(rc, "pq_parse_mgf")
    sp = dict(prec_mz=np.zeros(ns.value), charge=np.zeros(ns.value, dtype=np.int32),
              peak_off=np.zeros(ns.value + 1, dtype=np.uint64), mz=np.zeros(npk.value), inten=np.zeros(npk.value))
    rc = L.pq_parse_mgf(raw, C.c_uint64(len(raw)), C.byref(ns), C.byref(npk), _p(sp["prec_mz"]), _p(sp["charge"]),
                        _p(sp["peak_off"]), _p(sp["mz"]), _p(sp["inten"]))
    if rc != 0:
        raise PepQueryError(rc, "pq_parse_mgf")
    return sp


def write_mgf(sp, titles=None, charge_line=True):
    """CSR dict -> MGF text in the reference's own form (index/IndexWorker.asMgf :422-457): TITLE / PEPMASS / CHARGE, peaks as
    "%.4f %.2f".  Test and tool helper (the reference's writers stay Java)."""
    out = []
    off = sp["peak_off"]
    for i in range(len(sp["prec_mz"])):
        out.append("BEGIN IONS\nTITLE=%s\nPEPMASS=%r 1.0\n" % (titles[i] if titles is not None else "s:%d:%d" % (i + 1, sp["charge"][i]), float(sp["prec_mz"][i])))
        if charge_line and sp["charge"][i] > 0:
            out.append("CHARGE=%d+\n" % sp["charge"][i])
        a, b = int(off[i]), int(off[i + 1])
        out.append("".join("%.4f %.2f\n" % (m, v) for m, v in zip(sp["mz"][a:b], sp["inten"][a:b])))
        out.append("END IONS\n")
    return "".join(out)


def build_ms_library(sp, out_dir, titles=None):
    """index/BuildMSlibrary + IndexWorker.save2file (:474-486): one MGF per 0.1-Da bin of the neutral precursor mass,
    <round(10 * mass)>.mgf, spectra appended in input order.  Returns the titles it wrote."""
    from .shard import precursor_mass
    os.makedirs(out_dir, exist_ok=True)
    mass = precursor_mass(sp["prec_mz"], sp["charge"].astype(np.float64))
    bins = np.floor(mass * 10.0 + 0.5).astype(np.int64)                 # Math.round
    if titles is None:
        titles = ["run1:%d:%d" % (i + 1, sp["charge"][i]) for i in range(len(mass))]
    files = {}
    off = sp["peak_off"]
    for i in range(len(mass)):
        one = dict(prec_mz=sp["prec_mz"][i:i + 1], charge=sp["charge"][i:i + 1], peak_off=np.array([0, off[i + 1] - off[i]], dtype=np.uint64),
                   mz=sp["mz"][int(off[i]):int(off[i + 1])], inten=sp["inten"][int(off[i]):int(off[i + 1])])
        files.setdefault(int(bins[i]), []).append(write_mgf(one, [titles[i]]) + "\n")
    for b, chunks in files.items():
        with open(os.path.join(out_dir, "%d.mgf" % b), "w") as f:
            f.write("".join(chunks))
    return titles


def read_ms_library(path, mass_lo, mass_hi):
    """pq_read_ms_library: the spectra of an indexed MS/MS library with mass_lo <= neutral precursor mass <= mass_hi
    (msio/MsLibrarySearchWorker.loadSpectra :154-252) -> (CSR dict, titles)."""
    L = load_library()
    ns, npk, tb = C.c_int64(0), C.c_uint64(0), C.c_uint64(0)
    d = path.encode()
    rc = L.pq_read_ms_library(d, C.c_double(mass_lo), C.c_double(mass_hi), C.byref(ns), C.byref(npk), C.byref(tb), None, None, None, None, None, None, None)
    if rc != 0:
        raise PepQueryError(rc, "pq_read_ms_library")
    sp = dict(prec_mz=np.zeros(ns.value), charge=np.zeros(ns.value, dtype=np.int32), peak_off=np.zeros(ns.value + 1, dtype=np.uint64),
              mz=np.zeros(npk.value), inten=np.zeros(npk.value))
    toff = np.zeros(ns.value + 1, dtype=np.uint64); tbuf = C.create_string_buffer(max(1, tb.value))
    rc = L.pq_read_ms_library(d, C.c_double(mass_lo), C.c_double(mass_hi), C.byref(ns), C.byref(npk), C.byref(tb), _p(sp["prec_mz"]), _p(sp["charge"]),
                              _p(sp["peak_off"]), _p(sp["mz"]), _p(sp["inten"]), _p(toff), tbuf)
    if rc != 0:
        raise PepQueryError(rc, "pq_read_ms_library")
    raw = tbuf.raw
    titles = [raw[int(toff[i]):int(toff[i + 1])].decode() for i in range(ns.value)]
    return sp, titles
