"""Synthetic proteins / target peptides / spectra of the shape SURVEY.md §8d names.

Thin numpy wrapper over tools/libpq_synth.so (C++, deterministic per-index RNG streams).  Test and bench
infrastructure; not part of the scoring path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    src = os.path.join(_ROOT, "tools", "pq_synth.cpp")
    out = os.path.join(_ROOT, "tools", "libpq_synth.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", out, src])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "tools", "libpq_synth.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.syn_proteins.restype = C.c_int64
        L.syn_proteins.argtypes = [C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64]
        L.syn_targets.restype = C.c_int64
        L.syn_targets.argtypes = [C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64]
        L.syn_mass_bounds.restype = C.c_int32
        L.syn_mass_bounds.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.syn_spectra.restype = C.c_int64
        L.syn_spectra.argtypes = [C.c_int64, C.c_uint64, C.c_int64, C.c_double, C.c_double,
                                  C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def proteins(n, seed=20240101):
    """-> (offsets uint64[n+1], residues uint8[total])"""
    L = _lib()
    off = np.zeros(n + 1, dtype=np.uint64)
    total = L.syn_proteins(n, seed, _p(off), None, 0)
    res = np.zeros(total, dtype=np.uint8)
    L.syn_proteins(n, seed, _p(off), _p(res), total)
    return off, res


def targets(n, seed=7):
    """-> (offsets uint32[n+1], residues uint8[total])"""
    L = _lib()
    off = np.zeros(n + 1, dtype=np.uint32)
    total = L.syn_targets(n, seed, _p(off), None, 0)
    res = np.zeros(total, dtype=np.uint8)
    L.syn_targets(n, seed, _p(off), _p(res), total)
    return off, res


def mass_bounds(shards, tgt, prot, seed=99):
    toff, tres = tgt
    poff, pres = prot
    b = np.zeros(shards + 1, dtype=np.float64)
    _lib().syn_mass_bounds(seed, shards, len(toff) - 1, _p(toff), _p(tres), len(poff) - 1, _p(poff), _p(pres), _p(b))
    return b


def set_strips(bounds, world, rank):
    """Shard rule for the next spectra() calls: shard `rank` owns every world-th strip of the given mass boundaries
    (None switches back to the plain [mass_lo, mass_hi) rule)."""
    L = _lib()
    L.syn_set_strips.restype = None
    if bounds is None:
        L.syn_set_strips(C.c_int32(0), None, C.c_int32(1), C.c_int32(0))
    else:
        b = np.ascontiguousarray(bounds, dtype=np.float64)
        L.syn_set_strips(C.c_int32(len(b)), _p(b), C.c_int32(world), C.c_int32(rank))


def spectra(n, tgt, prot, seed=99, itol=0.6, first_index=0, mass_lo=0.0, mass_hi=1e9, threads=0, pinned=None):
    """-> dict(prec_mz f64[n], charge i32[n], peak_off u64[n+1], mz f64[P], inten f64[P]).

    pinned: optional callable (nbytes) -> numpy float64 array backed by pinned memory, for the peak arrays."""
    L = _lib()
    toff, tres = tgt
    poff, pres = prot
    prec = np.zeros(n, dtype=np.float64)
    chg = np.zeros(n, dtype=np.int32)
    off = np.zeros(n + 1, dtype=np.uint64)
    args = (n, seed, first_index, mass_lo, mass_hi, len(toff) - 1, _p(toff), _p(tres), len(poff) - 1, _p(poff), _p(pres), itol)
    total = L.syn_spectra(*args, _p(prec), _p(chg), _p(off), None, None, threads)
    if pinned is not None:
        mz = pinned(total)
        inten = pinned(total)
    else:
        mz = np.zeros(total, dtype=np.float64)
        inten = np.zeros(total, dtype=np.float64)
    L.syn_spectra(*args, _p(prec), _p(chg), _p(off), _p(mz), _p(inten), threads)
    return dict(prec_mz=prec, charge=chg, peak_off=off, mz=mz, inten=inten)


def planted_targets(prot, rng):
    """Substrings of the proteome (some with I/L swapped, lower case, at protein ends), a sequence that only exists across a
    protein boundary, and random sequences; -> offsets, residues, expected nhit (Python `in` on I->L folded strings)."""
    poff, pres = prot
    prots = [bytes(pres[int(poff[i]):int(poff[i + 1])]).decode() for i in range(len(poff) - 1)]
    fold = lambda s: s.upper().replace("I", "L")
    seqs = []
    for k in range(60):
        p = prots[int(rng.integers(len(prots)))]
        L = int(rng.integers(1 if k < 5 else 6, 30))
        a = int(rng.integers(0, max(1, len(p) - L)))
        if k % 7 == 0: a = 0
        if k % 7 == 1: a = max(0, len(p) - L)
        s = p[a:a + L]
        if k % 3 == 0: s = s.replace("L", "I")
        if k % 3 == 1: s = "".join("L" if (c == "I" and i % 2) else c for i, c in enumerate(s))
        if k % 5 == 0: s = s.lower()
        seqs.append(s)
    seqs.append(prots[0][-5:] + prots[1][:5])                       # crosses a protein boundary
    letters = "ACDEFGHKLMNPQRSTVWY"
    for k in range(40):
        seqs.append("".join(letters[int(x)] for x in rng.integers(0, len(letters), int(rng.integers(7, 26)))))
    seqs.append(prots[3][:-1] + "W" if not prots[3].endswith("W") else prots[3][:-1] + "A")     # a whole protein with its last residue changed
    seqs.append(prots[4])                                            # a whole protein
    folded = [fold(p) for p in prots]
    want = np.array([1 if any(fold(s) in p for p in folded) else 0 for s in seqs], dtype=np.int32)
    off = np.zeros(len(seqs) + 1, dtype=np.uint32); off[1:] = np.cumsum([len(s) for s in seqs])
    res = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()
    return off, res, want


def mgf_text(sp, out=None, threads=0):
    """The spectra as MGF text in the reference's own form (IndexWorker.asMgf); the synthetic precursor m/z carry five decimals and
    the peaks four / two, so parsing the text gives the arrays back bit for bit.  out: optional callable (nbytes) -> uint8 array
    (e.g. pinned memory) to receive the text.  -> uint8 array."""
    L = _lib()
    L.syn_mgf_build.restype = C.c_int64; L.syn_mgf_copy.restype = C.c_int64
    n = len(sp["prec_mz"])
    size = L.syn_mgf_build(C.c_int64(n), _p(sp["prec_mz"]), _p(sp["charge"]), _p(sp["peak_off"]), _p(sp["mz"]), _p(sp["inten"]), C.c_int32(threads))
    buf = out(size) if out is not None else np.empty(size, dtype=np.uint8)
    got = L.syn_mgf_copy(_p(buf), C.c_int64(size))
    assert got == size
    return buf[:size]
