"""ctypes mirror of include/pepquery_b200.h (POD structs only).

Kept byte-for-byte in step with the header; tests/test_abi.py checks the sizes against a C probe.
"""
import ctypes as C

PQ_MAX_ISOTOPE = 8
PQ_MAX_MODS = 8
PQ_MAX_PEP_LEN = 63
PQ_MAX_FORM_MODS = 16

PQ_OK = 0
PQ_ERR_BAD_ARG = -1
PQ_ERR_CAPACITY = -2
PQ_ERR_CUDA = -3
PQ_ERR_NO_DEVICE = -4
PQ_ERR_STATE = -5
PQ_ERR_UNSUPPORTED = -6
PQ_ERR_NOMEM = -7

PQ_MOD_ANYWHERE = 0
PQ_MOD_PEP_NTERM = 1
PQ_MOD_PEP_CTERM = 2
PQ_MOD_PEP_NTERM_AA = 3
PQ_MOD_PEP_CTERM_AA = 4
PQ_MOD_PROT_NTERM = 5
PQ_MOD_PROT_CTERM = 6
PQ_MOD_PROT_NTERM_AA = 7
PQ_MOD_PROT_CTERM_AA = 8


class pq_mod(C.Structure):
    _fields_ = [("id", C.c_int32), ("position", C.c_int32), ("residues", C.c_char * 24),
                ("delta", C.c_double), ("neutral_loss", C.c_double)]


class pq_params(C.Structure):
    _fields_ = [("tol", C.c_double), ("tol_ppm", C.c_int32), ("score_method", C.c_int32),
                ("itol", C.c_double), ("min_score", C.c_double), ("min_peaks", C.c_int32),
                ("n_isotope", C.c_int32), ("isotope", C.c_int32 * PQ_MAX_ISOTOPE),
                ("max_var_mods", C.c_int32), ("max_mods_per_aa", C.c_int32), ("fast", C.c_int32),
                ("hc", C.c_int32), ("n_random", C.c_int32), ("enzyme", C.c_int32),
                ("max_missed", C.c_int32), ("min_len", C.c_int32), ("max_len", C.c_int32),
                ("min_mass", C.c_double), ("max_mass", C.c_double),
                ("shuffle_seed", C.c_int64), ("mod_seed", C.c_int64),
                ("n_fixed", C.c_int32), ("n_var", C.c_int32),
                ("fixed", pq_mod * PQ_MAX_MODS), ("var", pq_mod * PQ_MAX_MODS),
                ("device", C.c_int32), ("score_all_shuffles", C.c_int32),
                ("prepare_at_load", C.c_int32), ("candidate_first", C.c_int32),
                ("search_at_load", C.c_int32), ("library_mode", C.c_int32)]


class pq_form(C.Structure):
    _fields_ = [("seq_index", C.c_int32), ("length", C.c_int32), ("mass", C.c_double),
                ("n_mods", C.c_int32), ("mod_site", C.c_uint8 * PQ_MAX_FORM_MODS),
                ("mod_index", C.c_int8 * PQ_MAX_FORM_MODS), ("seq", C.c_char * (PQ_MAX_PEP_LEN + 1))]


class pq_psm(C.Structure):
    _fields_ = [("spectrum", C.c_int32), ("form", C.c_int32), ("score", C.c_double),
                ("exp_mass", C.c_double), ("pep_mass", C.c_double), ("tol_ppm", C.c_double),
                ("isotope_error", C.c_int32), ("count_b", C.c_int32), ("count_y", C.c_int32),
                ("n_db", C.c_int32), ("total_db", C.c_int32), ("n_random", C.c_int32),
                ("total_random", C.c_int32), ("valid", C.c_int32), ("p_value", C.c_double),
                ("rank", C.c_int32), ("n_ptm", C.c_int32), ("confident", C.c_int32),
                ("charge", C.c_int32)]


class pq_competitor(C.Structure):
    _fields_ = [("spectrum", C.c_int32), ("ref_form", C.c_int32), ("ptm", C.c_int32),
                ("ptm_site", C.c_int32), ("score", C.c_double), ("pep_mass", C.c_double)]


class pq_ptm(C.Structure):
    _fields_ = [("id", C.c_int32), ("position", C.c_int32), ("residue", C.c_char),
                ("pad", C.c_char * 7), ("delta", C.c_double)]


class pq_counters(C.Structure):
    _fields_ = [("pairs_step2", C.c_uint64), ("pairs_step3", C.c_uint64), ("pairs_step4", C.c_uint64),
                ("pairs_step5", C.c_uint64), ("spectra_loaded", C.c_uint64), ("spectra_usable", C.c_uint64),
                ("spectra_prepared", C.c_uint64), ("target_forms", C.c_uint64),
                ("reference_sequences", C.c_uint64), ("reference_forms", C.c_uint64), ("psms", C.c_uint64),
                ("score_kernel_launches", C.c_uint64), ("other_kernel_launches", C.c_uint64),
                ("score_algorithmic_bytes", C.c_uint64),
                ("ms_prepare", C.c_double), ("ms_window", C.c_double), ("ms_score", C.c_double),
                ("ms_shuffle", C.c_double), ("ms_total", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("ms_score_step", C.c_double * 4), ("bytes_score_step", C.c_uint64 * 4),
                ("ms_host_reference", C.c_double), ("pairs_step4_bounded", C.c_uint64), ("pairs_step5_bounded", C.c_uint64),
                ("ms_ladder", C.c_double)]


# Modification constants pinned by the reference's resources/top_modifications.tsv (ids = -fixMod/-varMod numbers)
CARBAMIDOMETHYL_C = dict(id=1, position=PQ_MOD_ANYWHERE, residues=b"C", delta=57.02146372057, neutral_loss=0.0)
OXIDATION_M = dict(id=2, position=PQ_MOD_ANYWHERE, residues=b"M", delta=15.99491461956,
                   neutral_loss=63.99828574784)   # CH4OS
ACETYL_NTERM = dict(id=5, position=PQ_MOD_PEP_NTERM, residues=b"", delta=42.0105646837, neutral_loss=0.0)
PHOSPHO_S = dict(id=7, position=PQ_MOD_ANYWHERE, residues=b"S", delta=79.96633052074999, neutral_loss=97.97689520445)


def make_mod(**kw):
    m = pq_mod()
    m.id = kw["id"]; m.position = kw["position"]; m.residues = kw.get("residues", b"")
    m.delta = kw["delta"]; m.neutral_loss = kw.get("neutral_loss", 0.0)
    return m


def default_params(fixed=(CARBAMIDOMETHYL_C,), var=(OXIDATION_M,), **over):
    """Defaults of pg/CParameter.java:84-463, with -n 1000."""
    p = pq_params()
    p.tol = 10.0; p.tol_ppm = 1; p.score_method = 1; p.itol = 0.6; p.min_score = 12.0; p.min_peaks = 10
    p.n_isotope = 1; p.isotope[0] = 0
    p.max_var_mods = 3; p.max_mods_per_aa = 1; p.fast = 0; p.hc = 0; p.n_random = 1000
    p.enzyme = 1; p.max_missed = 2; p.min_len = 7; p.max_len = 45; p.min_mass = 500.0; p.max_mass = 10000.0
    p.shuffle_seed = 12457; p.mod_seed = 2022; p.prepare_at_load = 1; p.candidate_first = 1; p.search_at_load = 1
    p.n_fixed = len(fixed); p.n_var = len(var)
    for i, m in enumerate(fixed):
        p.fixed[i] = make_mod(**m)
    for i, m in enumerate(var):
        p.var[i] = make_mod(**m)
    p.device = 0
    for k, v in over.items():
        setattr(p, k, v)
    return p
