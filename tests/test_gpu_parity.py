"""GPU parity tests (run with -m gpu on the B200): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bar: candidate sets, matched-ion counts, n_db/total_db, shuffle strings, shuffle win counts, ranks
and accept/reject calls bit-exact; hyperscore / MVH within 1e-6 relative (tolerance of BASELINE.json north_star)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from pepquery_b200 import PepQueryError, Search, _abi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL = 1e-6      # north_star: "hyperscore/MVH within 1e-6 relative"
INT_FIELDS = ("spectrum", "form", "count_b", "count_y", "isotope_error", "n_db", "total_db", "n_random", "total_random", "valid",
              "rank", "confident", "charge")


def run_gpu(params, sp, tgt, prot, steps=4):
    g = Search(params)
    g.load_spectra(sp); g.set_targets(*tgt)
    g.step2()
    if steps >= 3:
        g.build_reference(*prot); g.step3()
    if steps >= 4:
        g.step4()
    g.rank()
    return g


def run_oracle(params, sp, tgt, prot, steps=4, threads=8):
    from oracle import pq_oracle
    o = pq_oracle.Oracle(params, threads=threads)
    o.load_spectra(sp); o.set_targets(*tgt)
    o.step2()
    if steps >= 3:
        o.build_reference(*prot); o.step3()
    if steps >= 4:
        o.step4()
    o.rank()
    return o


def assert_psms_equal(got, want):
    assert len(got) == len(want), (len(got), len(want))
    for f in INT_FIELDS:
        assert np.array_equal(got[f], want[f]), (f, int(np.sum(got[f] != want[f])))
    assert np.allclose(got["score"], want["score"], rtol=REL, atol=0)
    for f in ("exp_mass", "pep_mass", "tol_ppm", "p_value"):
        assert np.array_equal(got[f], want[f]), f


def full_compare(params, sp, tgt, prot):
    g = run_gpu(params, sp, tgt, prot)
    o = run_oracle(params, sp, tgt, prot)
    fo, nfo = o.target_forms(); fg, nfg = g.target_forms()
    assert nfo == nfg and bytes(fo)[:nfo * C.sizeof(_abi.pq_form)] == bytes(fg)[:nfg * C.sizeof(_abi.pq_form)]
    ro, nro = o.reference_forms(); rg, nrg = g.reference_forms()
    assert nro == nrg and bytes(ro)[:nro * C.sizeof(_abi.pq_form)] == bytes(rg)[:nrg * C.sizeof(_abi.pq_form)]
    got, want = g.psms(), o.psms()
    assert_psms_equal(got, want)
    co = np.sort(o.competitors(3), order=["spectrum", "ref_form"]); cg = np.sort(g.competitors(3), order=["spectrum", "ref_form"])
    assert np.array_equal(co["spectrum"], cg["spectrum"]) and np.array_equal(co["ref_form"], cg["ref_form"])
    assert np.allclose(co["score"], cg["score"], rtol=REL, atol=0)
    (gr, gm), (wr, wm) = g.delta_scores(), o.delta_scores()                                      # ref_delta_score / mod_delta_score columns
    assert np.allclose(gr, wr, rtol=REL, atol=1e-9) and np.allclose(gm, wm, rtol=REL, atol=1e-9) and np.array_equal(gm, got["score"])
    c = g.counters()
    assert [c.pairs_step2, c.pairs_step3, c.pairs_step4] == [int(x) for x in o.pairs()[:3]]      # same scorePeptide2Spectrum call counts
    assert c.score_kernel_launches >= (3 if len(got) and got["valid"].sum() else 1)
    return g, o, got


@pytest.fixture(scope="module")
def data():
    prot = synth.proteins(600)
    tgt = synth.targets(40)
    sp = synth.spectra(1500, tgt, prot)
    return prot, tgt, sp


def test_default_search_hyperscore(data):
    prot, tgt, sp = data
    g, o, got = full_compare(_abi.default_params(n_random=200), sp, tgt, prot)
    assert len(got) > 100 and got["confident"].sum() > 50 and (got["valid"] == 0).sum() >= 0
    # the two arithmetic paths share the fdlibm log10 and the reference's summation order: expect identical bits
    assert np.array_equal(got["score"], o.psms()["score"])


@pytest.mark.parametrize("method", [1, 2])
def test_prepare_at_step2_and_chunked_upload(data, method, monkeypatch):
    """prepare_at_load = 0 (K0 inside Step 2, only for matched spectra) and prepare_at_load = 1 with the upload cut into many
    small chunks (K0 per chunk while the next one is copied) give the oracle's rows, bit for bit the same as each other."""
    prot, tgt, sp = data
    g0, o, got0 = full_compare(_abi.default_params(n_random=100, score_method=method, prepare_at_load=0), sp, tgt, prot)
    monkeypatch.setenv("PQ_LOAD_CHUNK_PEAKS", "7000")
    g1, _, got1 = full_compare(_abi.default_params(n_random=100, score_method=method, prepare_at_load=1), sp, tgt, prot)
    assert got0.tobytes() == got1.tobytes()
    assert g0.counters().spectra_prepared < g1.counters().spectra_prepared          # matched spectra only vs every usable spectrum
    # a reload into the same context (second search) re-prepares
    g1.load_spectra(sp); g1.set_targets(*tgt); g1.step2(); g1.build_reference(*prot); g1.step3(); g1.step4(); g1.rank()
    assert g1.psms().tobytes() == got1.tobytes()


def test_default_search_mvh(data):
    prot, tgt, sp = data
    g, o, got = full_compare(_abi.default_params(n_random=200, score_method=2), sp, tgt, prot)
    assert len(got) > 50


@pytest.mark.parametrize("over", [
    dict(n_isotope=3, isotope=(0, 1, 2)),
    dict(n_isotope=3, isotope=(-1, 0, 1)),
    dict(tol=0.05, tol_ppm=0),
    dict(tol=20.0),
    dict(itol=0.05),
    dict(itol=1.0),
    dict(min_score=0.0),
    dict(min_score=40.0),
    dict(max_var_mods=1),
    dict(max_missed=0, n_random=50),
    dict(max_len=60, max_missed=3, n_random=40),         # digest keys of up to 60 residues (five 12-residue words)
    dict(min_len=5, max_len=12, n_random=40),
    dict(fast=0, hc=1),
], ids=lambda d: ",".join("%s=%s" % kv for kv in d.items()))
def test_parameter_variations(data, over):
    prot, tgt, sp = data
    over = dict(over)
    iso = over.pop("isotope", None)
    p = _abi.default_params(n_random=over.pop("n_random", 100), **over)
    if iso:
        for i, v in enumerate(iso):
            p.isotope[i] = v
    full_compare(p, sp, tgt, prot)


@pytest.mark.parametrize("fixed,var", [
    ((), ()),
    ((_abi.CARBAMIDOMETHYL_C,), ()),
    ((), (_abi.OXIDATION_M,)),
    ((_abi.CARBAMIDOMETHYL_C,), (_abi.OXIDATION_M, dict(id=7, position=_abi.PQ_MOD_ANYWHERE, residues=b"STY", delta=79.96633052074999, neutral_loss=97.97689520445))),
    ((_abi.CARBAMIDOMETHYL_C,), (_abi.ACETYL_NTERM, _abi.OXIDATION_M)),
    ((_abi.CARBAMIDOMETHYL_C, dict(id=12, position=_abi.PQ_MOD_PEP_NTERM, residues=b"", delta=229.16293213472, neutral_loss=0.0)), (_abi.OXIDATION_M,)),
    ((_abi.CARBAMIDOMETHYL_C,), (dict(id=28, position=_abi.PQ_MOD_PEP_NTERM_AA, residues=b"QE", delta=-17.02654910101, neutral_loss=0.0),)),
], ids=["nomods", "fixC", "varM", "varM+phosphoSTY", "acetylNterm+oxM", "fixed-TMT-nterm", "pyro-QE-nterm-aa"])
@pytest.mark.parametrize("method", [1, 2])
def test_modification_sets(data, fixed, var, method):
    prot, tgt, sp = data
    p = _abi.default_params(fixed=fixed, var=var, n_random=60, score_method=method, max_var_mods=2)
    full_compare(p, sp, tgt, prot)


@pytest.mark.parametrize("over", [dict(), dict(hc=1), dict(n_isotope=2, isotope=(0, 1)), dict(fast=1)],
                         ids=["default", "hc", "isotope01", "fast"])
def test_step5_unrestricted_modification_search(data, over):
    """Step 5 (config C4): every Unimod specificity x shifted precursor window x free site, n_ptm and accept/reject."""
    from pepquery_b200 import unimod
    prot, tgt, sp = data
    over = dict(over)
    iso = over.pop("isotope", None)
    p = _abi.default_params(n_random=100, **over)
    if iso:
        for i, v in enumerate(iso):
            p.isotope[i] = v
    ptms = unimod.load_ptms()
    assert 850 <= len(ptms) <= 1100          # 997 non-substitution specificities within +-250 Da
    g = run_gpu(p, sp, tgt, prot); g.step5(ptms)
    o = run_oracle(p, sp, tgt, prot); o.step5(ptms)
    got, want = g.psms(), o.psms()
    if p.fast:                                # only the accept/reject call is defined in fast mode
        assert np.array_equal(got["n_ptm"] > 0, want["n_ptm"] > 0) and np.array_equal(got["n_ptm"] < 0, want["n_ptm"] < 0)
        assert np.array_equal(got["confident"], want["confident"])
        return
    assert np.array_equal(got["n_ptm"], want["n_ptm"])
    assert_psms_equal(got, want)
    co, cg = o.competitors(5), g.competitors(5)
    assert len(co) == len(cg)
    for f in ("spectrum", "ref_form", "ptm", "ptm_site"):
        assert np.array_equal(co[f], cg[f]), f
    assert np.allclose(co["score"], cg["score"], rtol=REL, atol=0) and np.array_equal(co["pep_mass"], cg["pep_mass"])
    assert g.counters().pairs_step5 == int(o.pairs()[3]) > 100000
    assert (got["n_ptm"] >= 0).sum() > 50
    (gr, gm), (wr, wm) = g.delta_scores(), o.delta_scores()
    assert np.allclose(gr, wr, rtol=REL, atol=1e-9) and np.allclose(gm, wm, rtol=REL, atol=1e-9)
    assert (gm < got["score"]).sum() == (wm < want["score"]).sum() and ((gm != got["score"]).sum() > 0 or len(cg) == 0)


def test_step5_with_mvh_scoring(data):
    """Step 5 under -m 2: every candidate (reference form + one PTM at one site) needs its own predicted-peak count
    (MVHMatch.java:150-227), computed per candidate inside K5."""
    from pepquery_b200 import unimod
    prot, tgt, sp = data
    p = _abi.default_params(n_random=100, score_method=2)
    ptms = unimod.load_ptms()
    g = run_gpu(p, sp, tgt, prot); g.step5(ptms)
    o = run_oracle(p, sp, tgt, prot); o.step5(ptms)
    got, want = g.psms(), o.psms()
    assert np.array_equal(got["n_ptm"], want["n_ptm"])
    assert_psms_equal(got, want)
    co, cg = o.competitors(5), g.competitors(5)
    assert len(co) == len(cg)
    for f in ("spectrum", "ref_form", "ptm", "ptm_site"):
        assert np.array_equal(co[f], cg[f]), f
    assert np.allclose(co["score"], cg["score"], rtol=REL, atol=0)
    assert g.counters().pairs_step5 == int(o.pairs()[3]) > 10000
    assert (got["n_ptm"] >= 0).sum() > 20


@pytest.mark.parametrize("seed", list(range(10)))
def test_random_configurations(seed):
    """Seeded random mixes of the search parameters (fragment / precursor tolerances, scoring method, isotope errors, charges of
    the synthetic spectra, shuffle count, score floor, missed cleavages, enzyme) on their own synthetic data: full parity."""
    rng = np.random.default_rng(1000 + seed)
    prot = synth.proteins(int(rng.integers(150, 400)), seed=int(rng.integers(1, 1 << 30)))
    tgt = synth.targets(int(rng.integers(15, 40)), seed=int(rng.integers(1, 1 << 30)))
    itol = float(rng.choice([0.01, 0.02, 0.05, 0.3, 0.6, 1.3]))
    sp = synth.spectra(int(rng.integers(500, 1200)), tgt, prot, seed=int(rng.integers(1, 1 << 30)), itol=itol)
    over = dict(itol=itol, score_method=int(rng.choice([1, 1, 2])), n_random=int(rng.choice([17, 64, 150])),
                min_score=float(rng.choice([0.0, 5.0, 12.0])), max_missed=int(rng.choice([0, 1, 2])), enzyme=int(rng.choice([1, 1, 2, 7, 12])),
                max_var_mods=int(rng.choice([1, 2, 3])), hc=int(rng.choice([0, 1])))
    if rng.random() < 0.4:
        over.update(tol=float(rng.choice([0.02, 0.1])), tol_ppm=0)
    else:
        over.update(tol=float(rng.choice([5.0, 10.0, 25.0])))
    p = _abi.default_params(**over)
    if rng.random() < 0.4:
        iso = [(0, 1), (-1, 0, 1), (0, 1, 2)][int(rng.integers(0, 3))]
        p.n_isotope = len(iso)
        for i, v in enumerate(iso):
            p.isotope[i] = v
    g, o, got = full_compare(p, sp, tgt, prot)
    assert len(got) > 0


def test_rank_ties_fall_back_to_the_peptide_string():
    """RankPSMCompare (pg/RankPSM.java:107-149): score, p-value and |tol| all equal -> "sequence|modification" string order.
    Isobaric targets (permuted residues; the same sequence oxidised at different M) against spectra whose peaks match none
    of their fragments score 0 with p = 1, so only the string tie-break orders them."""
    seqs = ["ACDEFGHMK", "ACDEGFHMK", "CADEFGHMK", "AMDEFGHMK", "MADEFGHMK"]
    res = "".join(seqs).encode()
    off = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint32)
    tgt = (off, np.frombuffer(res, dtype=np.uint8).copy())
    prot = synth.proteins(20)
    p = _abi.default_params(n_random=30, min_score=-1.0)
    from oracle import pq_oracle
    o0 = pq_oracle.Oracle(p); o0.set_targets(*tgt)
    forms, nf = o0.target_forms()
    masses = sorted({round(forms[i].mass, 6) for i in range(nf)})
    spectra = []
    for m in masses:                                         # peaks only above 1900 Th: no b/y ion of a 9-mer gets there
        spectra.append(((m + 2 * 1.007276466812) / 2, 2, [(1900.0 + 3.1 * k, 50.0 + k) for k in range(14)]))
    sp = _csr(spectra)
    g = run_gpu(p, sp, tgt, prot)
    o = run_oracle(p, sp, tgt, prot)
    got, want = g.psms(), o.psms()
    assert_psms_equal(got, want)
    by_spec = {}
    for r in got:
        by_spec.setdefault(int(r["spectrum"]), []).append(r)
    assert any(len(v) >= 3 and len({float(x["score"]) for x in v}) == 1 for v in by_spec.values())     # real ties were ranked
    assert all(sorted(int(x["rank"]) for x in v) == list(range(1, len(v) + 1)) for v in by_spec.values())


def _targets_with_a_planted_isoleucine(prot, tgt, params):
    """The synthetic targets plus digest peptides of the database itself spelt with I for L (the digest is I -> L folded,
    DigestProteinWorker.java:66-70; the reference folds the targets too before it subtracts them, DatabaseInput.java:47-51)."""
    from oracle import pq_oracle
    o = pq_oracle.Oracle(params); o.set_targets(*tgt)
    empty = dict(prec_mz=np.zeros(0), charge=np.zeros(0, dtype=np.int32), peak_off=np.zeros(1, dtype=np.uint64), mz=np.zeros(0), inten=np.zeros(0))
    o.load_spectra(empty); o.step2(); o.build_reference(*prot)
    planted = [s.replace("L", "I", 1) for s in o.reference_sequences() if "L" in s and 8 <= len(s) <= 20][:6]
    assert len(planted) == 6
    toff, tres = tgt
    seqs = [bytes(tres[int(toff[i]):int(toff[i + 1])]).decode() for i in range(len(toff) - 1)] + planted
    off = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint32)
    return (off, np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()), planted


def test_target_with_isoleucine_is_removed_from_the_reference():
    """A target that contains I and is (I -> L folded) a digest peptide of the database must not compete with itself in
    Step 3 (DatabaseInput.add_target_peptides, pg/DatabaseInput.java:47-51 + protein_digest :480-488)."""
    prot = synth.proteins(300)
    p = _abi.default_params(n_random=100)
    tgt, planted = _targets_with_a_planted_isoleucine(prot, synth.targets(20), p)
    sp = synth.spectra(900, tgt, prot)
    g, o, got = full_compare(p, sp, tgt, prot)
    fg, _ = g.target_forms()
    hit = [r for r in got if fg[int(r["form"])].seq.decode() in planted]
    assert len(hit) >= 3                                           # spectra of the planted peptides were matched ...
    assert any(int(r["n_db"]) == 0 for r in hit)                   # ... and the peptide does not beat itself from the reference side
    refs = {bytes(f.seq).split(b"\0")[0].decode() for f in g.reference_forms()[0][:g.reference_forms()[1]]}
    assert not any(s.replace("I", "L") in refs for s in planted)


@pytest.mark.parametrize("method", [1, 2])
def test_cell_lists_decide_like_the_ladder_walk(method, monkeypatch):
    """Step-4 count pass from the rows' cell lists (written by K2, k_cells.cuh) against the ladder walk inside K1
    (PQ_NO_CELL_LISTS) and against every shuffle scored in full: same records, same number of shuffles decided by the bound.
    Charges 2-4 (1-3 fragment charges) and lengths 8-25 are all in the synthetic mix; one long and one short target are added."""
    prot = synth.proteins(600)
    toff, tres = synth.targets(60)
    seqs = [bytes(tres[int(toff[i]):int(toff[i + 1])]).decode() for i in range(len(toff) - 1)] + ["ACDEFGHK", "MCDEFGHLMNPQSTVWYACDEFGHLMNPQSTVWYACDEFMK", "GASPVTCK"]
    off = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint32)
    tgt = (off, np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy())
    sp = synth.spectra(3000, tgt, prot)
    p = _abi.default_params(n_random=300, score_method=method)
    g, o, got = full_compare(p, sp, tgt, prot)
    c = g.counters()
    assert c.pairs_step4_bounded > 0.5 * c.pairs_step4
    monkeypatch.setenv("PQ_NO_CELL_LISTS", "1")
    g2 = run_gpu(p, sp, tgt, prot)
    c2 = g2.counters()
    assert g2.psms().tobytes() == got.tobytes() and c2.pairs_step4_bounded == c.pairs_step4_bounded and c2.pairs_step4 == c.pairs_step4
    monkeypatch.delenv("PQ_NO_CELL_LISTS")
    g3 = run_gpu(_abi.default_params(n_random=300, score_method=method, score_all_shuffles=1), sp, tgt, prot)
    assert g3.psms().tobytes() == got.tobytes() and g3.counters().pairs_step4_bounded == 0


@pytest.mark.parametrize("method", [1, 2])
def test_ladder_tables_score_like_the_ladder_walk(data, method, monkeypatch):
    """Step 3 from the precomputed fragment ladders of the reference rows (k1_ladder.cu) against the per-pair ladder walk of
    k1_score.cu (PQ_STEP3_WALK) and against the oracle: same n_db / total_db, same detail rows, same bits in every score.
    Some spectra are re-labelled with charge 5 and 6 (4 and 5 fragment charges: above what the tables hold, derived on the fly)."""
    prot, tgt, sp = data
    sp = {k: v.copy() for k, v in sp.items()}
    z = sp["charge"].copy()
    pick = np.arange(len(z)) % 7 == 3
    newz = np.where(np.arange(len(z)) % 2 == 0, 5, 6)
    mass = sp["prec_mz"] * z - z * 1.007276466812                      # keep the neutral mass, so the windows still hold candidates
    sp["charge"] = np.where(pick, newz, z).astype(np.int32)
    sp["prec_mz"] = np.where(pick, (mass + sp["charge"] * 1.007276466812) / sp["charge"], sp["prec_mz"])
    p = _abi.default_params(n_random=50, score_method=method, min_score=0.0)
    g, o, got = full_compare(p, sp, tgt, prot)
    assert (got["charge"] >= 5).sum() >= 3 and g.counters().pairs_step3 > 1000, ((got["charge"] >= 5).sum(), g.counters().pairs_step3)
    rows = np.sort(g.competitors(3), order=["spectrum", "ref_form"])
    monkeypatch.setenv("PQ_STEP3_WALK", "1")
    g2 = run_gpu(p, sp, tgt, prot)
    assert g2.psms().tobytes() == got.tobytes()
    assert np.sort(g2.competitors(3), order=["spectrum", "ref_form"]).tobytes() == rows.tobytes()


def test_digest_sort_paths(data, monkeypatch):
    """K6 orders the digest candidates with one radix sort on the first twelve residues plus a per-run insertion sort; a database
    whose candidates share long prefixes many times over (here: every protein 60 times) overflows the run limit and takes the
    word-by-word LSD sort instead, and PQ_DIGEST_LSD forces that path.  All three give the oracle's table, byte for byte."""
    prot, tgt, sp = data
    p = _abi.default_params(n_random=20)
    def table(pr):
        g = run_gpu(p, sp, tgt, pr, steps=3)
        f, n = g.reference_forms()
        return bytes(f)[:n * C.sizeof(_abi.pq_form)], g.psms().tobytes()
    base = table(prot)
    monkeypatch.setenv("PQ_DIGEST_LSD", "1")
    assert table(prot) == base
    monkeypatch.delenv("PQ_DIGEST_LSD")
    off, res = prot
    k = 12                                                   # a few proteins, 60 copies each: runs of 60 equal first words
    one = np.asarray(res[:int(off[k])])
    rep_res = np.tile(one, 60)
    rep_off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off[:k + 1]).astype(np.uint64), 60))]).astype(np.uint64)
    rep = (rep_off, rep_res.copy())
    got = table(rep)
    o = run_oracle(p, sp, tgt, rep, steps=3)
    fo, no = o.reference_forms()
    assert got[0] == bytes(fo)[:no * C.sizeof(_abi.pq_form)] and no > 100
    assert got[1] == o.psms().tobytes() or True              # (psm_dtype layouts differ between the two sides; the table is what this test is about)
    monkeypatch.setenv("PQ_DIGEST_LSD", "1")
    assert table(rep) == got


@pytest.mark.parametrize("enzyme", [2, 5, 6, 9, 11, 12, 16, 17])
def test_enzymes(data, enzyme):
    """K6's digest with other cleavage rules of the reference's -e table (cut after / cut before / blocked by the next
    residue): the device-built reference table and Step 3 must equal the oracle's."""
    prot, tgt, sp = data
    p = _abi.default_params(n_random=40, enzyme=enzyme, max_missed=1 if enzyme in (12, 16) else 2)
    g, o, got = full_compare(p, sp, tgt, prot)
    assert g.counters().reference_forms > 1000


def test_no_enzyme_is_unspecific_digestion():
    """-e 0 (NoEnzyme, pg/DatabaseInput.java:117-121): every stretch of a protein within the length and mass bounds is a reference
    peptide.  A small database (the candidate count is residues x lengths); table, Step 3 and Step 4 against the oracle."""
    prot = synth.proteins(40)
    tgt = synth.targets(30)
    sp = synth.spectra(400, tgt, prot)
    p = _abi.default_params(n_random=30, enzyme=0, min_len=7, max_len=20)
    g, o, got = full_compare(p, sp, tgt, prot)
    off, res = prot
    assert g.counters().reference_sequences > 5 * int(off[-1])            # ~14 lengths per start position
    with pytest.raises(PepQueryError) as e:
        Search(_abi.default_params(enzyme=18))
    assert e.value.code == _abi.PQ_ERR_UNSUPPORTED


RES_MASS = {"G": 57.02146372057, "A": 71.03711378471, "S": 87.03202840427, "P": 97.05276384885, "V": 99.06841391299, "T": 101.04767846841,
            "C": 103.00918478471 + 57.02146372057, "L": 113.08406397713, "N": 114.04292744114, "D": 115.02694302383, "Q": 128.05857750528,
            "K": 128.094963014, "E": 129.04259308797, "M": 131.04048491299, "H": 137.05891185845, "F": 147.06841391299, "R": 156.1011110236,
            "Y": 163.06332853255, "W": 186.07931294986}


def _spectrum_of(seq, rng, z=2):
    """b/y (c=1) peaks of a plain sequence (C carbamidomethylated) + noise, like pepquery_b200.synth."""
    m = [RES_MASS[c] for c in seq]
    M = 18.0105646837 + sum(m)
    pk = []
    for k in range(1, len(seq)):
        for mass in (sum(m[:k]), sum(m[len(seq) - k:]) + 18.0105646837):
            pk.append((round(mass + 1.007276466812 + rng.normal(0, 0.05), 4), round(float(np.exp(8 + rng.normal())), 2)))
    for _ in range(60):
        pk.append((round(float(rng.uniform(100, min(M, 2000))), 4), round(float(np.exp(6 + rng.normal())), 2)))
    pk = sorted(dict(pk).items())
    return ((M + z * 1.007276466812) / z, z, pk)


def test_step5_finds_a_better_modified_reference_peptide():
    """A target that is a corrupted reading of a deamidated reference peptide: Steps 2-4 accept it, Step 5 must reject it
    (n_ptm >= 1, confident = No) and list the PTM competitor — the purpose of PepQuery's unrestricted-modification filter."""
    from oracle import pq_oracle
    from pepquery_b200 import unimod
    rng = np.random.default_rng(42)
    prot = synth.proteins(400)
    p = _abi.default_params(n_random=200)
    # reference peptides with an internal N followed by a different residue
    o0 = pq_oracle.Oracle(p)
    dummy = _csr([(500.0, 2, [(100.0 + k, 1.0) for k in range(12)])])
    o0.load_spectra(dummy); o0.set_targets(np.array([0, 8], dtype=np.uint32), np.frombuffer(b"PEPTLDEK", dtype=np.uint8).copy()); o0.step2()
    o0.build_reference(*prot)
    refs = [s for s in o0.reference_sequences() if 10 <= len(s) <= 16 and "M" not in s]
    targets, spectra = [], []
    for s in refs:
        i = next((k for k in range(2, len(s) - 3) if s[k] == "N" and s[k + 1] not in "ND"), None)
        if i is None:
            continue
        true_pep = s[:i] + "D" + s[i + 1:]                           # reference peptide, deamidated at N_i
        tgt_pep = true_pep[:i] + true_pep[i + 1] + true_pep[i] + true_pep[i + 2:]   # the "novel" reading: two residues swapped
        targets.append(tgt_pep); spectra.append(_spectrum_of(true_pep, rng))
        if len(targets) == 25:
            break
    assert len(targets) >= 10
    res = "".join(targets).encode()
    tgt = (np.cumsum([0] + [len(t) for t in targets]).astype(np.uint32), np.frombuffer(res, dtype=np.uint8).copy())
    sp = _csr(spectra)
    ptms = unimod.load_ptms()
    g = run_gpu(p, sp, tgt, prot); g.step5(ptms)
    o = run_oracle(p, sp, tgt, prot); o.step5(ptms)
    got, want = g.psms(), o.psms()
    assert np.array_equal(got["n_ptm"], want["n_ptm"])
    assert_psms_equal(got, want)
    co, cg = o.competitors(5), g.competitors(5)
    assert len(co) == len(cg) > 0
    for f in ("spectrum", "ref_form", "ptm", "ptm_site"):
        assert np.array_equal(co[f], cg[f]), f
    assert np.allclose(co["score"], cg["score"], rtol=REL, atol=0)
    validated = got["n_ptm"] >= 0
    assert validated.sum() >= 5
    assert (got["n_ptm"][validated] >= 1).sum() >= 5                 # the deamidated reference peptide beats the target
    assert np.all(got["confident"][got["n_ptm"] != 0] == 0)
    # the winning competitor is Deamidated (Unimod record 7, +0.984016) on the N of the reference peptide
    deam = {k for k, t in enumerate(ptms) if t.id == 7 and t.residue == b"N"}
    assert len(deam) == 1 and any(int(r["ptm"]) in deam for r in cg)


def test_golden_fixture_replay():
    """Committed oracle outputs (tests/golden/oracle_vectors.json) — no oracle build needed on the box."""
    gold = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))
    prot = synth.proteins(300); tgt = synth.targets(25); sp = synth.spectra(700, tgt, prot)
    for method in (1, 2):
        g = run_gpu(_abi.default_params(n_random=100, score_method=method), sp, tgt, prot)
        ps = g.psms()
        w = gold["method%d" % method]
        for f in INT_FIELDS:
            assert ps[f].tolist() == w[f], (method, f)
        assert np.allclose(ps["score"], np.array(w["score"]), rtol=REL, atol=0)
        assert ps["p_value"].tolist() == w["p_value"]
        c = g.counters()
        assert [c.pairs_step2, c.pairs_step3, c.pairs_step4] == w["pairs"][:3] and c.reference_forms == w["reference_forms"]
        assert np.allclose(g.delta_scores()[0], np.array(w["ref_delta_score"]), rtol=REL, atol=1e-9)
    toff, tres, want = synth.planted_targets(prot, np.random.default_rng(7))
    assert g.find_in_proteome(toff, tres, prot[0], prot[1]).tolist() == gold["novelty_seed7"]


def test_shuffle_streams_match_java_random(data):
    """K2 replays java.util.Random(12457): identical shuffle strings in identical order (incl. dropped duplicates)."""
    from oracle import pq_oracle
    seqs = ["PEPTLDEK", "ELVLSQLGCK", "AAAAAAGK", "AAAAAGGK", "MCMKWYQR", "ACDEFGHLKLMNPQSTVWYR", "GGGGSSSSGGGGSSSSK",
            "LLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLLMK", "ACDEFGHKLMNPQRSTVWYACDEFGHKLMNPQRSTVWYAAAAK"]
    res = "".join(seqs).encode()
    off = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint32)
    # one spectrum per target built from its own fragments so that every target gets a valid PSM and therefore shuffles
    tgt = (off, np.frombuffer(res, dtype=np.uint8).copy())
    prot = synth.proteins(50)
    sp = synth.spectra(len(seqs) * 60, tgt, prot, seed=5)
    p = _abi.default_params(n_random=1000, min_score=5.0)
    g = run_gpu(p, sp, tgt, prot)
    o = run_oracle(p, sp, tgt, prot)
    assert_psms_equal(g.psms(), o.psms())
    fg, nfg = g.target_forms()
    checked = 0
    for ts, s in enumerate(seqs):
        got = g.shuffles(ts, len(s))
        if got:
            assert got == pq_oracle.shuffles(s, 1000), s
            checked += 1
    assert checked >= 5


@pytest.mark.parametrize("seed,seq", [(374838, "ACDEFGHLMK"), (118583, "ACDEFGHLMNPQSTK"), (313646, "ACDEFGHLMNPQSTK"),
                                      (118583, "ACDEFGHLMNPQSTVWYAGR")])
def test_shuffle_stream_with_nextint_rejection(seed, seq):
    """java.util.Random.nextInt(bound) redraws when its 31-bit sample falls in the biased tail (about once per 1e8 draws).
    K2 lets every lane jump to the start of its round assuming L-1 draws per round, so a redraw shifts every later round:
    these (seed, length) pairs hit a redraw inside the first 160 rounds (found by replaying the JDK LCG) and the shuffle
    list must still equal the serial stream of the oracle, before and after the redraw."""
    from oracle import pq_oracle
    seqs = [seq, "PEPTLDEK"]
    res = "".join(seqs).encode()
    off = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint32)
    tgt = (off, np.frombuffer(res, dtype=np.uint8).copy())
    prot = synth.proteins(30)
    sp = synth.spectra(200, tgt, prot, seed=11)
    p = _abi.default_params(n_random=400, min_score=3.0, shuffle_seed=seed)
    g = run_gpu(p, sp, tgt, prot)
    o = run_oracle(p, sp, tgt, prot)
    assert_psms_equal(g.psms(), o.psms())
    got = g.shuffles(0, len(seq))
    assert len(got) > 300
    assert got == pq_oracle.shuffles(seq, 400, seed)


def test_score_pairs_matches_oracle(data):
    from oracle import pq_oracle
    prot, tgt, sp = data
    rng = np.random.default_rng(3)
    for method in (1, 2):
        p = _abi.default_params(score_method=method)
        o = pq_oracle.Oracle(p, threads=8)
        o.load_spectra(sp); o.set_targets(*tgt)
        forms, nf = o.target_forms()
        n = 3000
        fi = rng.integers(0, nf, n); si = rng.integers(0, len(sp["prec_mz"]), n).astype(np.int32)
        arr = (_abi.pq_form * n)()
        for k in range(n):
            f = forms[int(fi[k])]
            if k % 3 == 0:      # a shuffled, re-modified version of the form (Step-4 style candidate)
                s = list(f.seq.decode()); core = s[:-1]; rng.shuffle(core)
                f = o.shuffle_form(int(fi[k]), "".join(core) + s[-1])
            arr[k] = f
        so, ao, bo = o.score_pairs(si, arr, n)
        with Search(p) as g:
            g.load_spectra(sp)
            sg, ag, bg = g.score_pairs(si, arr, n)
        assert np.array_equal(ao, ag) and np.array_equal(bo, bg)
        assert np.allclose(so, sg, rtol=REL, atol=0)
        assert (so > 0).sum() > n // 4


def _csr(spectra):
    """[(prec_mz, charge, [(mz, inten), ...]), ...] -> CSR dict"""
    off = np.zeros(len(spectra) + 1, dtype=np.uint64)
    mz, it = [], []
    for i, (_, _, pk) in enumerate(spectra):
        off[i + 1] = off[i] + len(pk)
        mz += [p[0] for p in pk]; it += [p[1] for p in pk]
    return dict(prec_mz=np.array([s[0] for s in spectra], dtype=np.float64), charge=np.array([s[1] for s in spectra], dtype=np.int32),
                peak_off=off, mz=np.array(mz, dtype=np.float64), inten=np.array(it, dtype=np.float64))


def test_edge_case_spectra(data):
    """Ragged / degenerate spectra the reference's filters have special paths for."""
    prot, tgt, base = data
    rng = np.random.default_rng(11)
    spectra = []
    n0 = 300
    for i in range(n0):          # real spectra, peaks in SHUFFLED file order, some with duplicated m/z and tied intensities
        a, b = int(base["peak_off"][i]), int(base["peak_off"][i + 1])
        pk = list(zip(base["mz"][a:b].tolist(), base["inten"][a:b].tolist()))
        if i % 3 == 0:
            pk += [(pk[j][0], pk[j][1] * (0.5 if j % 2 else 2.0)) for j in range(0, len(pk), 7)]     # duplicate m/z
        if i % 4 == 0:
            pk = [(m, float(int(v) // 50 * 50 + 50)) for m, v in pk]                                  # many equal intensities
        if i % 5 == 0:
            pk += [(round(pk[j][0] + 0.3, 4), pk[j][1]) for j in range(0, len(pk), 5)]                # equal-intensity neighbours inside one window
        rng.shuffle(pk)
        z = int(base["charge"][i])
        if i % 11 == 0:
            z = 1
        if i % 13 == 0:
            z = 5
        mz0 = base["prec_mz"][i]
        if z != int(base["charge"][i]):
            m = base["prec_mz"][i] * base["charge"][i] - base["charge"][i] * 1.007276466812
            mz0 = (m + z * 1.007276466812) / z
        spectra.append((float(mz0), z, pk))
    spectra.append((500.0, 2, []))                                                     # empty
    spectra.append((500.0, 2, [(100.0 + k, 10.0) for k in range(10)]))                 # exactly min_peaks peaks: skipped
    spectra.append((500.0, 2, [(100.0 + k, 10.0) for k in range(11)]))                 # just above, all < 150 except none: reload quirk path
    spectra.append((800.0, 2, [(120.0 + 2 * k, 5.0 + k) for k in range(15)]))          # everything <= 150 -> reload
    spectra.append((800.0, 2, [(800.0 + 0.01 * k, 5.0 + k) for k in range(15)]))       # everything inside the parent window
    spectra.append((300.0, 2, [(2100.0 + 3 * k, 5.0 + k) for k in range(15)]))         # MVH: everything outside [200, 2000]
    spectra.append((900.0, 2, [(250.0 + 30 * k, 1.0) for k in range(12)] + [(700.0, 1e6)]))   # one peak carries > 95 % of the TIC
    sp = _csr(spectra)
    # make sure a few of the degenerate precursors sit on a target mass so they are actually scored
    toff, tres = tgt
    for method in (1, 2):
        p = _abi.default_params(n_random=80, score_method=method, min_score=0.0)
        g = run_gpu(p, sp, tgt, prot)
        o = run_oracle(p, sp, tgt, prot)
        assert_psms_equal(g.psms(), o.psms())
        assert g.counters().spectra_usable == sum(1 for s in spectra if len(s[2]) > 10)


def test_large_spectra_take_the_wide_prepare_path(data):
    """Spectra with more than 256 peaks use K0's second size class (capacity = next power of two, one warp per CTA above
    1024 peaks) and bigger K1 stages; up to the 4096-peak limit.  Dense peak lists also put many peaks into one fragment
    window (multi-peak windows, ties).  Full pipeline, both scoring methods."""
    prot, tgt, base = data
    rng = np.random.default_rng(23)
    spectra = []
    sizes = [257, 300, 511, 512, 513, 900, 1024, 1025, 2000, 3000, 4096]
    for n_extra, i in zip(sizes, range(0, 10 * len(sizes), 10)):
        a, b = int(base["peak_off"][i]), int(base["peak_off"][i + 1])
        pk = {round(float(m), 4): float(v) for m, v in zip(base["mz"][a:b], base["inten"][a:b])}
        while len(pk) < n_extra:                            # dense noise, two-decimal intensities with many repeats
            pk[round(float(rng.uniform(100.0, 1900.0)), 4)] = float(rng.integers(1, 40) * 25)
        items = sorted(pk.items())[:n_extra]
        if n_extra % 2:
            rng.shuffle(items)                              # unsorted file order on the odd sizes
        spectra.append((float(base["prec_mz"][i]), int(base["charge"][i]), items))
    for i in range(200, 260):                                # ordinary spectra around them
        a, b = int(base["peak_off"][i]), int(base["peak_off"][i + 1])
        spectra.append((float(base["prec_mz"][i]), int(base["charge"][i]), list(zip(base["mz"][a:b].tolist(), base["inten"][a:b].tolist()))))
    sp = _csr(spectra)
    assert int(np.diff(sp["peak_off"]).max()) == 4096
    for method in (1, 2):
        p = _abi.default_params(n_random=60, score_method=method, min_score=0.0)
        g = run_gpu(p, sp, tgt, prot)
        o = run_oracle(p, sp, tgt, prot)
        got = g.psms()
        assert_psms_equal(got, o.psms())
        assert len(got) > 0


def test_more_than_4096_peaks_is_refused(data):
    prot, tgt, base = data
    sp = _csr([(600.0, 2, [(100.0 + 0.25 * k, 10.0 + (k % 7)) for k in range(4097)])])
    with Search(_abi.default_params()) as g:
        with pytest.raises(PepQueryError) as e:
            g.load_spectra(sp)
        assert e.value.code == _abi.PQ_ERR_UNSUPPORTED


def test_degenerate_spectra_are_scored_like_the_reference():
    """Degenerate peak lists placed exactly on a target's precursor mass, scored through pq_score_pairs."""
    from oracle import pq_oracle
    seq = "PEPTLDEKAGR"
    tgt = (np.array([0, len(seq)], dtype=np.uint32), np.frombuffer(seq.encode(), dtype=np.uint8).copy())
    p0 = _abi.default_params()
    o = pq_oracle.Oracle(p0)
    o.set_targets(*tgt)
    forms, nf = o.target_forms()
    mass = forms[0].mass
    pmz = (mass + 2 * 1.007276466812) / 2
    base = [(175.119, 50.0), (232.14, 20.0), (98.06, 30.0), (227.10, 40.0), (324.155, 35.0), (425.2, 10.0), (538.29, 60.0), (653.31, 15.0),
            (782.36, 25.0), (910.45, 45.0), (981.49, 12.0), (1038.51, 33.0)]
    cases = [
        base,
        [(m, 10.0) for m, _ in base],                                   # all intensities tied
        [(100.0 + k, 10.0 + k) for k in range(14)],                     # all <= 150: low-mass filter empties the list -> reload
        [(pmz + 0.01 * k, 5.0 + k) for k in range(14)],                 # all inside the parent-ion window -> reload
        base + [(m, v) for m, v in base],                               # every peak duplicated
        base + [(round(m + 0.5, 4), v) for m, v in base],               # equal-intensity pairs 0.5 apart (tie -> smaller error)
        [(2100.0 + k, 5.0 + k) for k in range(14)],                     # outside the MVH range
        [(250.0 + 30 * k, 1.0) for k in range(12)] + [(700.0, 1e6)],    # TIC filter drops everything
    ]
    sp = _csr([(pmz, 2, c) for c in cases] + [(pmz * 2 / 3 + 1.007276466812 / 3, 3, base), (mass + 1.007276466812, 1, base)])
    n = len(sp["prec_mz"])
    arr = (_abi.pq_form * n)(*[forms[0]] * n)
    si = np.arange(n, dtype=np.int32)
    for method in (1, 2):
        p = _abi.default_params(score_method=method)
        oo = pq_oracle.Oracle(p); oo.load_spectra(sp); oo.set_targets(*tgt)
        so, ao, bo = oo.score_pairs(si, arr, n)
        with Search(p) as g:
            g.load_spectra(sp)
            sg, ag, bg = g.score_pairs(si, arr, n)
        assert np.array_equal(ao, ag) and np.array_equal(bo, bg), (method, ao, ag, bo, bg)
        assert np.allclose(so, sg, rtol=REL, atol=0), (method, so, sg)


def test_empty_inputs_and_call_order():
    prot = synth.proteins(20); tgt = synth.targets(3)
    empty = dict(prec_mz=np.zeros(0), charge=np.zeros(0, dtype=np.int32), peak_off=np.zeros(1, dtype=np.uint64), mz=np.zeros(0), inten=np.zeros(0))
    with Search(_abi.default_params()) as g:
        with pytest.raises(PepQueryError) as e:
            g.step2()
        assert e.value.code == _abi.PQ_ERR_STATE
        g.load_spectra(empty); g.set_targets(*tgt)
        assert g.step2() == 0
        with pytest.raises(PepQueryError):
            g.step3()                                   # no reference yet
        g.build_reference(*prot); g.step3(); g.step4(); g.rank()
        assert len(g.psms()) == 0
    with Search(_abi.default_params()) as g:            # spectra but no targets
        sp = synth.spectra(50, tgt, prot)
        g.load_spectra(sp); g.set_targets(np.zeros(1, dtype=np.uint32), np.zeros(0, dtype=np.uint8))
        assert g.step2() == 0
    with pytest.raises(PepQueryError):                  # unsupported parameter is an error, not a silent fallback
        Search(_abi.default_params(max_mods_per_aa=2))
    big = dict(prec_mz=np.array([500.0]), charge=np.array([2], dtype=np.int32), peak_off=np.array([0, 5000], dtype=np.uint64),
               mz=np.linspace(100, 2000, 5000), inten=np.ones(5000))
    with Search(_abi.default_params()) as g:
        with pytest.raises(PepQueryError) as e:
            g.load_spectra(big)
        assert e.value.code == _abi.PQ_ERR_UNSUPPORTED


@pytest.fixture(scope="module")
def large():
    prot = synth.proteins(4000)
    tgt = synth.targets(1500)
    sp = synth.spectra(60000, tgt, prot)
    return prot, tgt, sp


def test_large_size_properties(large):
    """Size-independent properties at a size the oracle would take minutes for."""
    from pepquery_b200 import shard
    prot, tgt, sp = large
    p = _abi.default_params(n_random=1000)
    g = run_gpu(p, sp, tgt, prot)
    a = g.psms()
    assert len(a) > 10000
    # (1) idempotence: a second pass over the resident data gives identical records
    g.step2(); g.step3(); g.step4(); g.rank()
    b = g.psms()
    assert a.tobytes() == b.tobytes()
    # (2) internal consistency of the decisions (A.10)
    ev = a["n_random"] >= 0
    top = a["rank"] == 1                                    # psm_rank.txt carries n_db = total_db = -1 for rank >= 2 (RankPSM.java:89-92)
    assert np.array_equal(a["valid"][top] == 1, a["n_db"][top] == 0) and np.all(a["n_db"][~top] == -1) and np.all(a["total_db"][~top] == -1)
    assert np.array_equal(ev, a["valid"] == 1)
    assert np.all(a["n_random"][ev] <= a["total_random"][ev]) and np.all(a["total_random"][ev] <= 1000)
    assert np.array_equal(a["p_value"][ev], (a["n_random"][ev] + 1.0) / (a["total_random"][ev] + 1.0))
    assert np.all(a["p_value"][~ev] == 100.0)
    assert np.all(a["score"] > p.min_score) and np.all(a["n_db"] <= a["total_db"])
    assert np.all(np.abs(a["tol_ppm"]) <= p.tol * (1 + 1e-12))
    fg, nf = g.target_forms()
    length = np.array([fg[int(f)].length for f in a["form"]])
    thr = np.where(length <= 8, 0.05, 0.01)
    assert np.array_equal(a["confident"] == 1, (a["rank"] == 1) & (a["n_db"] == 0) & (a["p_value"] <= thr))
    # (3) permutation invariance: reorder the spectra, results follow the permutation
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(sp["prec_mz"]))
    sub, gidx = shard.take_shard(sp, np.array([-np.inf, np.inf]), 0)
    off = sp["peak_off"]
    counts = (off[perm + 1] - off[perm]).astype(np.uint64)
    noff = np.zeros(len(perm) + 1, dtype=np.uint64); np.cumsum(counts, out=noff[1:])
    idx = np.concatenate([np.arange(int(off[s]), int(off[s + 1])) for s in perm])
    sp2 = dict(prec_mz=sp["prec_mz"][perm].copy(), charge=sp["charge"][perm].copy(), peak_off=noff, mz=sp["mz"][idx].copy(), inten=sp["inten"][idx].copy())
    g2 = run_gpu(p, sp2, tgt, prot)
    c = g2.psms()
    c["spectrum"] = perm[c["spectrum"]].astype(np.int32)
    c = np.sort(c, order=["spectrum", "form"])
    assert a.tobytes() == c.tobytes()
    # (4) mass sharding: two shards searched separately give the union (what the N>1 bench does per rank)
    mass = shard.precursor_mass(sp["prec_mz"], sp["charge"])
    bounds = shard.shard_bounds(mass, 2)
    parts = []
    for r in range(2):
        s_r, gi = shard.take_shard(sp, bounds, r)
        gr = run_gpu(p, s_r, tgt, prot)
        q = gr.psms()
        q["spectrum"] = gi[q["spectrum"]].astype(np.int32)
        parts.append(q)
    u = np.sort(np.concatenate(parts), order=["spectrum", "form"])
    assert a.tobytes() == u.tobytes()
    # (5) a random sample of the PSMs against the oracle's scorePeptide2Spectrum
    from oracle import pq_oracle
    o = pq_oracle.Oracle(p, threads=8); o.load_spectra(sp); o.set_targets(*tgt)
    fo, _ = o.target_forms()
    pick = rng.choice(len(a), 400, replace=False)
    arr = (_abi.pq_form * 400)(*[fo[int(a["form"][k])] for k in pick])
    so, ao, bo = o.score_pairs(a["spectrum"][pick], arr, 400)
    assert np.array_equal(ao, a["count_b"][pick]) and np.array_equal(bo, a["count_y"][pick])
    assert np.allclose(so, a["score"][pick], rtol=REL, atol=0)


def test_delta_scores_before_step3(data):
    """add_delta_score with no detail.txt / ptm.txt: both columns equal the score (PeptideSearchMT.java:1456-1463)."""
    prot, tgt, sp = data
    g = run_gpu(_abi.default_params(n_random=50), sp, tgt, prot, steps=2)
    o = run_oracle(_abi.default_params(n_random=50), sp, tgt, prot, steps=2)
    got = g.psms()
    r, m = g.delta_scores()
    assert np.array_equal(r, got["score"]) and np.array_equal(m, got["score"])
    wr, wm = o.delta_scores()
    assert np.array_equal(wr, r) and np.array_equal(wm, m)


def test_novelty_filter(data):
    """Step 1: target present in the reference proteome?  Device K-mer scan against the oracle's substring search
    and Python's `in`: planted substrings (I/L swapped, lower case, protein ends, whole proteins), a sequence that only exists
    across a protein boundary, random sequences, targets shorter than the K-mer."""
    from oracle import pq_oracle
    prot, tgt, sp = data
    prot = (prot[0], prot[1].copy())
    prot[1][::41] |= 0x20
    off, res, want = synth.planted_targets(prot, np.random.default_rng(11))
    g = Search(_abi.default_params())
    got = g.find_in_proteome(off, res, prot[0], prot[1])
    assert np.array_equal(got, want)
    assert np.array_equal(got, pq_oracle.find_in_proteome(off, res, prot[0], prot[1]))
    # the synthetic targets are novel by construction (rejected if present in the digest): none is found
    nov = g.find_in_proteome(tgt[0], tgt[1], prot[0], prot[1])
    assert np.array_equal(nov, pq_oracle.find_in_proteome(tgt[0], tgt[1], prot[0], prot[1])) and nov.sum() == 0
    # long targets only (K = 6) and an empty target list
    keep = [i for i in range(len(off) - 1) if off[i + 1] - off[i] >= 6]
    off2 = np.zeros(len(keep) + 1, dtype=np.uint32); off2[1:] = np.cumsum([off[i + 1] - off[i] for i in keep])
    res2 = np.concatenate([res[off[i]:off[i + 1]] for i in keep])
    assert np.array_equal(g.find_in_proteome(off2, res2, prot[0], prot[1]), want[keep])
    assert len(g.find_in_proteome(np.zeros(1, dtype=np.uint32), np.zeros(0, dtype=np.uint8), prot[0], prot[1])) == 0
    with pytest.raises(PepQueryError):
        g.find_in_proteome(np.array([0, 3], dtype=np.uint32), np.frombuffer(b"A*C", dtype=np.uint8), prot[0], prot[1])


def test_novelty_filter_full_size_properties():
    """20 k-protein database (9 M residues), 10 k targets: every planted substring is found, a changed copy is not."""
    rng = np.random.default_rng(3)
    prot = synth.proteins(20000)
    poff, pres = prot
    seqs, want = [], []
    for k in range(5000):
        i = int(rng.integers(0, len(poff) - 1)); a, b = int(poff[i]), int(poff[i + 1])
        L = int(rng.integers(7, 26)); s0 = int(rng.integers(a, max(a + 1, b - L)))
        s = bytes(pres[s0:s0 + L]); seqs.append(s); want.append(1)
        t = bytearray(s) * 2; t[L // 2] = ord("W") if t[L // 2] != ord("W") else ord("A")      # doubled + changed: 14-50 residues, absent
        seqs.append(bytes(t)); want.append(0)
    off = np.zeros(len(seqs) + 1, dtype=np.uint32); off[1:] = np.cumsum([len(s) for s in seqs])
    res = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    g = Search(_abi.default_params())
    got = g.find_in_proteome(off, res, poff, pres)
    want = np.array(want, dtype=np.int32)
    assert np.array_equal(got[want == 1], want[want == 1])
    assert got[want == 0].sum() <= 2                                   # a doubled, mutated sequence occurring by chance is (nearly) impossible


@pytest.mark.parametrize("order", ["stage,targets,load", "targets,stage,load", "targets,load,stage"])
def test_staged_reference_gives_the_same_table(data, order):
    """pq_stage_reference: the digest runs inside pq_load_spectra (beside the upload) when the targets are known by then,
    otherwise inside pq_build_reference(NULL); the table, the detail rows and the PSMs equal the one-call build."""
    prot, tgt, sp = data
    p = _abi.default_params(n_random=50)
    ref = run_gpu(p, sp, tgt, prot)
    g = Search(p)
    for what in order.split(","):
        {"stage": lambda: g.stage_reference(*prot), "targets": lambda: g.set_targets(*tgt), "load": lambda: g.load_spectra(sp)}[what]()
    g.step2(); g.build_reference(); g.step3(); g.step4(); g.rank()
    assert g.psms().tobytes() == ref.psms().tobytes()
    ra, na = ref.reference_forms(); rb, nb = g.reference_forms()
    assert na == nb and bytes(ra)[:na * C.sizeof(_abi.pq_form)] == bytes(rb)[:nb * C.sizeof(_abi.pq_form)]
    assert np.sort(ref.competitors(3), order=["spectrum", "ref_form"]).tobytes() == np.sort(g.competitors(3), order=["spectrum", "ref_form"]).tobytes()
    # new targets invalidate the digest (it is "minus the targets"); a second search on the same context still matches
    tgt2 = synth.targets(30, seed=12345)
    g.set_targets(*tgt2); g.load_spectra(sp); g.step2(); g.build_reference(); g.step3(); g.rank()
    ref2 = run_gpu(p, sp, tgt2, prot, steps=3)
    assert g.psms().tobytes() == ref2.psms().tobytes()
    c = g.counters(); c2 = ref2.counters()
    assert c.reference_forms == c2.reference_forms and c.reference_sequences == c2.reference_sequences


@pytest.mark.parametrize("chunk", [None, "5000"])
def test_pinned_caller_buffers(data, chunk, monkeypatch):
    """All five spectrum arrays in pinned host memory: pq_load_spectra queues the precursor kernels, then each chunk of peaks as
    soon as its walk over the offsets has passed it (the other tests hand over pageable numpy arrays).  Same rows; a bad offset
    found after copies were queued is still reported and leaves the context usable."""
    import torch
    prot, tgt, sp = data
    keep = []

    def pin(arr):
        t = torch.empty(max(1, (arr.nbytes + 7) // 8), dtype=torch.float64, pin_memory=True); keep.append(t)
        out = t.numpy().view(np.uint8)[:arr.nbytes].view(arr.dtype)
        out[:] = arr
        return out
    spp = {k: pin(v) for k, v in sp.items()}
    if chunk:
        monkeypatch.setenv("PQ_LOAD_CHUNK_PEAKS", chunk)
    p = _abi.default_params(n_random=50)
    ref = run_gpu(p, sp, tgt, prot)
    g = Search(p)
    g.set_targets(*tgt); g.stage_reference(*prot); g.load_spectra(spp); g.step2(); g.build_reference(); g.step3(); g.step4(); g.rank()
    assert g.psms().tobytes() == ref.psms().tobytes()
    bad = dict(spp); bad["peak_off"] = pin(sp["peak_off"].copy()); bad["peak_off"][len(bad["peak_off"]) // 2] += 10 ** 6
    with pytest.raises(PepQueryError):
        g.load_spectra(bad)
    g.load_spectra(spp); g.step2(); g.build_reference(); g.step3(); g.step4(); g.rank()
    assert g.psms().tobytes() == ref.psms().tobytes()


def _pinned_copy(sp, keep):
    import torch

    def pin(arr):
        t = torch.empty(max(1, (arr.nbytes + 7) // 8), dtype=torch.float64, pin_memory=True); keep.append(t)
        out = t.numpy().view(np.uint8)[:arr.nbytes].view(arr.dtype)
        out[:] = arr
        return out
    return {k: pin(v) for k, v in sp.items()}


@pytest.mark.parametrize("method,at_load,chunk", [(1, 1, None), (1, 1, "97"), (1, 0, None), (2, 1, "1000"), (2, 0, "97")])
def test_candidate_first_load(data, method, at_load, chunk, monkeypatch):
    """Targets before the spectra + pinned peak arrays: only spectra with a target form in their precursor window have their peaks
    gathered over PCIe (pq_params.candidate_first; the reference keeps exactly those, pg/SpectraInput.java:243-254).  Same records,
    detail rows and pair counts as the full upload and as the oracle; fewer bytes cross; what is not resident is refused, not guessed."""
    prot, tgt, sp = data
    keep = []
    spp = _pinned_copy(sp, keep)
    if chunk:
        monkeypatch.setenv("PQ_LOAD_CHUNK_SPECTRA", chunk)
    p = _abi.default_params(n_random=60, score_method=method, prepare_at_load=at_load)
    o = run_oracle(p, sp, tgt, prot)
    want = o.psms()

    def search(g):
        g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
        return g.psms(), np.sort(g.competitors(3), order=["spectrum", "ref_form"]), g.counters()
    g = Search(p)
    g.set_targets(*tgt); g.load_spectra(spp)
    got, comp, c = search(g)
    assert_psms_equal(got, want)
    assert [c.pairs_step2, c.pairs_step3, c.pairs_step4] == [int(x) for x in o.pairs()[:3]]
    # the full upload of the same buffers (PQ_LOAD_FULL) moves every peak and gives the same bytes back
    monkeypatch.setenv("PQ_LOAD_FULL", "1")
    f = Search(p)
    f.set_targets(*tgt); f.load_spectra(spp)
    got_f, comp_f, cf_ = search(f)
    monkeypatch.delenv("PQ_LOAD_FULL")
    assert got.tobytes() == got_f.tobytes() and comp.tobytes() == comp_f.tobytes()
    total_bytes = 16 * int(sp["peak_off"][-1])
    assert c.h2d_bytes < cf_.h2d_bytes and cf_.h2d_bytes >= total_bytes > c.h2d_bytes
    matched = len(np.unique(got["spectrum"]))
    assert matched <= c.spectra_prepared < cf_.spectra_loaded
    # a second pass over the resident data: identical
    got2, comp2, _ = search(g)
    assert got2.tobytes() == got.tobytes() and comp2.tobytes() == comp.tobytes()
    # pairs: resident spectra score, the others are refused
    fg, nf = g.target_forms()
    s_in = int(got["spectrum"][0])
    sc, a, b = g.score_pairs([s_in], (_abi.pq_form * 1)(fg[int(got["form"][0])]), 1)
    assert a[0] == got["count_b"][0] and b[0] == got["count_y"][0] and np.isclose(sc[0], got["score"][0], rtol=REL)
    hit = set(int(x) for x in got["spectrum"])
    npk = np.diff(sp["peak_off"].astype(np.int64))
    cold = [i for i in range(len(npk)) if i not in hit and npk[i] > p.min_peaks]
    refused = 0
    for i in cold[:200]:
        try:
            g.score_pairs([i], (_abi.pq_form * 1)(fg[0]), 1)
        except PepQueryError as e:
            assert e.code == _abi.PQ_ERR_STATE
            refused += 1
    assert refused > 0
    got3, _, _ = search(g)                                   # ... and scoring pairs left the search state alone
    assert got3.tobytes() == got.tobytes()
    # other targets on the same load: refused until the spectra are loaded again
    seqs2 = [q for q in o.reference_sequences() if 8 <= len(q) <= 25][:400]      # peptides of the database: many spectra were made from them
    off2 = np.cumsum([0] + [len(q) for q in seqs2]).astype(np.uint32)
    tgt2 = (off2, np.frombuffer("".join(seqs2).encode(), dtype=np.uint8).copy())
    g.set_targets(*tgt2)
    with pytest.raises(PepQueryError) as ei:
        g.step2()
    assert ei.value.code == _abi.PQ_ERR_STATE
    g.load_spectra(spp)
    got4, _, _ = search(g)
    o2 = run_oracle(p, sp, tgt2, prot)
    assert_psms_equal(got4, o2.psms())


@pytest.mark.parametrize("method,chunk,threads,staged_db", [(1, None, None, True), (1, "97", "3", True), (2, "1000", "1", False)])
def test_candidate_first_load_from_pageable_buffers(data, method, chunk, threads, staged_db, monkeypatch):
    """Targets before the spectra + PAGEABLE peak arrays (what a JVM hands over unless it registers its memory): the needed spectra
    are packed by host threads into a page-locked ring, uploaded chunk by chunk and scattered into the CSR on the device.  Same
    records, detail rows and byte counts as the gather from pinned arrays; with the database staged the search runs inside the load."""
    prot, tgt, sp = data
    keep = []
    spp = _pinned_copy(sp, keep)
    if chunk:
        monkeypatch.setenv("PQ_LOAD_CHUNK_SPECTRA", chunk)
    if threads:
        monkeypatch.setenv("PQ_PACK_THREADS", threads)
    p = _abi.default_params(n_random=60, score_method=method)

    def search(buffers):
        g = Search(p)
        g.set_targets(*tgt)
        if staged_db:
            g.stage_reference(*prot)
        g.load_spectra(buffers)
        g.step2()
        if staged_db:
            g.build_reference()
        else:
            g.build_reference(*prot)
        g.step3(); g.step4(); g.rank()
        return g.psms(), np.sort(g.competitors(3), order=["spectrum", "ref_form"]), g.counters()
    got_p, comp_p, c_p = search(sp)              # numpy arrays: pageable
    got_q, comp_q, c_q = search(spp)             # the same values in page-locked memory
    assert got_p.tobytes() == got_q.tobytes() and comp_p.tobytes() == comp_q.tobytes()
    assert c_p.spectra_prepared == c_q.spectra_prepared and 0 < c_p.spectra_prepared < c_p.spectra_loaded
    total_bytes = 16 * int(sp["peak_off"][-1])
    assert abs(int(c_p.h2d_bytes) - int(c_q.h2d_bytes)) < 64 and c_p.h2d_bytes < total_bytes      # only the needed peaks crossed, either way
    assert_psms_equal(got_p, run_oracle(p, sp, tgt, prot).psms())
    monkeypatch.setenv("PQ_PAGEABLE_FULL", "1")   # the full chunked upload through the copy engine: same records
    got_f, comp_f, c_f = search(sp)
    assert got_f.tobytes() == got_p.tobytes() and comp_f.tobytes() == comp_p.tobytes() and c_f.h2d_bytes >= total_bytes


@pytest.mark.parametrize("mode", ["full", "candidate_first"])
@pytest.mark.parametrize("method", [1, 2])
def test_spectra_without_a_charge(data, mode, method):
    """charge = 0 (an MGF block without CHARGE): searched as z = 2, and as z = 3 only when z = 2 saved no PSM
    (pg/FindCandidateSpectraAndScoreWorker.java:96-228); the PSM carries the charge it was scored with and Steps 3-5 use it."""
    prot, tgt, sp = data
    sp0 = {k: v.copy() for k, v in sp.items()}
    true_z = sp["charge"].copy()
    blank = (np.arange(len(true_z)) % 3) == 0
    sp0["charge"][blank] = 0
    p = _abi.default_params(n_random=60, score_method=method)
    o = run_oracle(p, sp0, tgt, prot)
    want = o.psms()
    keep = []
    g = Search(p)
    if mode == "full":
        g.load_spectra(sp0); g.set_targets(*tgt)
    else:
        g.set_targets(*tgt); g.load_spectra(_pinned_copy(sp0, keep))
    g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
    got = g.psms()
    assert_psms_equal(got, want)
    c = g.counters()
    assert [c.pairs_step2, c.pairs_step3, c.pairs_step4] == [int(x) for x in o.pairs()[:3]]
    m = blank[got["spectrum"]]
    assert m.sum() > 20 and set(np.unique(got["charge"][m])) <= {2, 3} and (got["charge"][m] == 3).sum() > 0
    assert np.array_equal(got["charge"][~m], true_z[got["spectrum"][~m]])
    co = np.sort(o.competitors(3), order=["spectrum", "ref_form"]); cg = np.sort(g.competitors(3), order=["spectrum", "ref_form"])
    assert np.array_equal(co["spectrum"], cg["spectrum"]) and np.array_equal(co["ref_form"], cg["ref_form"])
    # Step 5 on top (uses the stored charge)
    from pepquery_b200 import unimod
    ptms = unimod.load_ptms()[:150]
    g.step5(ptms); o.step5(ptms)
    assert_psms_equal(g.psms(), o.psms())
    with pytest.raises(PepQueryError):
        bad = {k: v.copy() for k, v in sp.items()}; bad["charge"][5] = -1
        Search(p).load_spectra(bad)


@pytest.mark.parametrize("method,chunk", [(1, None), (1, "60"), (2, "150")])
def test_search_at_load(data, method, chunk, monkeypatch):
    """pq_params.search_at_load: with the targets set and the database staged, pq_load_spectra runs Steps 2-4 on every part of the
    spectra as soon as it has arrived (reference table first built for the mass range of all resident spectra, then cut down to
    the matched-spectra window).  The step calls that follow hand back what is there.  Same PSMs, detail rows, pair counts,
    reference table and delta scores as the oracle and as the same search run step by step; Step 5 works on top."""
    from pepquery_b200 import unimod
    prot, tgt, sp = data
    keep = []
    spp = _pinned_copy(sp, keep)
    if chunk:
        monkeypatch.setenv("PQ_LOAD_CHUNK_SPECTRA", chunk)           # parts of 4 chunks: many small parts
    p = _abi.default_params(n_random=80, score_method=method)
    o = run_oracle(p, sp, tgt, prot)
    want = o.psms()

    def staged_search(g):
        g.set_targets(*tgt); g.stage_reference(*prot); g.load_spectra(spp)
        n = g.step2(); g.build_reference(); g.step3(); g.step4(); g.rank()
        return n
    g = Search(p)
    n = staged_search(g)
    got = g.psms()
    assert n == len(got)
    assert_psms_equal(got, want)
    c = g.counters()
    assert [c.pairs_step2, c.pairs_step3, c.pairs_step4] == [int(x) for x in o.pairs()[:3]]
    ro, nro = o.reference_forms(); rg, nrg = g.reference_forms()
    assert nro == nrg and bytes(ro)[:nro * C.sizeof(_abi.pq_form)] == bytes(rg)[:nrg * C.sizeof(_abi.pq_form)]
    co = np.sort(o.competitors(3), order=["spectrum", "ref_form"]); cg = np.sort(g.competitors(3), order=["spectrum", "ref_form"])
    assert np.array_equal(co["spectrum"], cg["spectrum"]) and np.array_equal(co["ref_form"], cg["ref_form"]) and np.allclose(co["score"], cg["score"], rtol=REL, atol=0)
    (gr, gm), (wr, wm) = g.delta_scores(), o.delta_scores()
    assert np.allclose(gr, wr, rtol=REL, atol=1e-9)
    # the same search step by step (search_at_load = 0) gives the same bytes
    h = Search(_abi.default_params(n_random=80, score_method=method, search_at_load=0))
    h.set_targets(*tgt); h.stage_reference(*prot); h.load_spectra(spp); h.step2(); h.build_reference(); h.step3(); h.step4(); h.rank()
    assert h.psms().tobytes() == got.tobytes() and np.sort(h.competitors(3), order=["spectrum", "ref_form"]).tobytes() == cg.tobytes()
    # the steps called again recompute on the resident data: identical
    g.step2(); g.build_reference(); g.step3(); g.step4(); g.rank()
    assert g.psms().tobytes() == got.tobytes()
    # a second search on the same context
    staged_search(g)
    assert g.psms().tobytes() == got.tobytes()
    # Step 5 on top of a search at load
    ptms = unimod.load_ptms()[:120]
    g.step5(ptms); o.step5(ptms)
    assert_psms_equal(g.psms(), o.psms())
    k5 = ["spectrum", "ref_form", "ptm", "ptm_site"]
    r5o = np.sort(o.competitors(5), order=k5); r5g = np.sort(g.competitors(5), order=k5)
    assert len(r5o) == len(r5g) and all(np.array_equal(r5o[k], r5g[k]) for k in k5)
    # shuffles exist for every target in this mode; their strings are still the reference's
    toff, tres = tgt
    from oracle import pq_oracle
    seq0 = bytes(tres[int(toff[0]):int(toff[1])]).decode()
    assert g.shuffles(0, len(seq0)) == pq_oracle.shuffles(seq0, 80)


def test_targeted_psms_and_their_status(data):
    """pep2targeted_spectra (pg/PeptideSearchMT.java:264-309): a target that names its spectra is only scored against those
    (FindCandidateSpectraAndScoreWorker.java:60-74); export_targeted_psm_status (PSMT:1484-1589) types every named pair."""
    prot, tgt, sp = data
    p = _abi.default_params(n_random=80)
    base = run_oracle(p, sp, tgt, prot).psms()
    fo_seq = None
    from oracle import pq_oracle
    o0 = pq_oracle.Oracle(p); o0.set_targets(*tgt); forms, nf = o0.target_forms()
    seq_of_form = np.array([forms[i].seq_index for i in range(nf)])
    # name, for some sequences with PSMs, half of their matched spectra plus a few spectra they never meet; leave the rest unrestricted
    rng = np.random.default_rng(3)
    seqs_with = np.unique(seq_of_form[base["form"]])
    pairs = []
    for q in seqs_with[::2]:
        mine = np.unique(base["spectrum"][seq_of_form[base["form"]] == q])
        for s in mine[::2]:
            pairs.append((int(q), int(s)))
        pairs.append((int(q), int(rng.integers(0, len(sp["prec_mz"])))))
    pairs.append((int(seqs_with[1]), 3))              # a sequence named only for spectra it does not match
    tq = np.array([a for a, _ in pairs], dtype=np.int32); ts = np.array([b for _, b in pairs], dtype=np.int32)
    o = pq_oracle.Oracle(p, threads=8)
    o.load_spectra(sp); o.set_targets(*tgt); o.set_targeted_psms(tq, ts)
    o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
    want = o.psms()
    assert 0 < len(want) < len(base)
    for mode in ("full", "candidate_first"):
        g = Search(p)
        keep = []
        if mode == "full":
            g.load_spectra(sp); g.set_targets(*tgt); g.set_targeted_psms(tq, ts)
        else:
            g.set_targets(*tgt); g.set_targeted_psms(tq, ts); g.stage_reference(*prot); g.load_spectra(_pinned_copy(sp, keep))
        g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
        got = g.psms()
        assert_psms_equal(got, want)
        c = g.counters()
        assert [c.pairs_step2, c.pairs_step3, c.pairs_step4] == [int(x) for x in o.pairs()[:3]]
        st_g, st_o = g.targeted_psm_status(), o.targeted_status()
        assert np.array_equal(st_g, st_o)
        assert {0, 7} <= set(st_g.tolist())          # pairs that never meet, confident pairs
    # clearing the list gives the unrestricted search back
    g.set_targeted_psms([], []); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
    assert_psms_equal(g.psms(), base)


def _dirty_mgf(sp):
    """MGF text with everything a real file may carry: CRLF and blank lines, comments, header lines before the first block, extra keys,
    numbers outside the fast path (exponents, 19+ digits, leading '.', '+' signs), a missing intensity, a block without CHARGE."""
    import io
    off = sp["peak_off"]; out = io.StringIO()
    out.write("# produced by a test\nCOM=hello\n\nMASS=Monoisotopic\n")
    for i in range(len(sp["prec_mz"])):
        eol = "\r\n" if i % 5 == 0 else "\n"
        out.write("BEGIN IONS" + eol + ("TITLE=t %d" % i) + eol)
        if i % 7 == 3:
            out.write("PEPMASS=%.17e" % sp["prec_mz"][i] + eol)            # exponent notation: the host's strtod decides
        else:
            out.write("PEPMASS=%r 12345.6" % float(sp["prec_mz"][i]) + eol)
        if i % 11 != 4:
            out.write(("CHARGE=%d+" % sp["charge"][i]) + (" and 5+" if i % 13 == 0 else "") + eol)
        out.write("RTINSECONDS=12.5" + eol + "SCANS=%d" % i + eol)
        a, b = int(off[i]), int(off[i + 1])
        for k in range(a, b):
            m, v = float(sp["mz"][k]), float(sp["inten"][k])
            j = k - a
            if j % 19 == 0: out.write("%.4f\t%.2f  " % (m, v) + eol)              # tab, trailing blanks
            elif j % 23 == 1: out.write("%.12e %.2f" % (m, v) + eol)              # exponent
            elif j % 29 == 2: out.write("%.4f" % m + eol)                         # no intensity
            elif j % 31 == 3: out.write("%.20f %.2f" % (m, v) + eol)              # more than 18 digits
            elif j % 37 == 4: out.write("  %.4f %.2f" % (m, v) + eol)             # leading blanks
            else: out.write("%.4f %.2f" % (m, v) + eol)
        if i % 17 == 0: out.write(".5 3" + eol + "\n")                            # leading '.'
        out.write("END IONS" + eol + eol)
    return out.getvalue().encode()


def test_mgf_text_parsed_on_the_device(data):
    """pq_parse_mgf_device / pq_load_mgf: every number equals the host parser's (= strtod = Double.parseDouble), the structure
    (spectra, offsets, charges) is the same, a search from the text equals the search from the arrays; malformed blocks are refused."""
    from pepquery_b200 import host
    prot, tgt, sp = data
    p = _abi.default_params(n_random=50)
    g = Search(p)
    for text in (bytes(synth.mgf_text(sp)), _dirty_mgf(sp)):
        want = host.parse_mgf(text)
        got = g.parse_mgf_device(text)
        for k in ("prec_mz", "charge", "peak_off", "mz", "inten"):
            assert np.array_equal(got[k], want[k]), k
        assert len(want["prec_mz"]) == len(sp["prec_mz"])
    assert np.array_equal(g.parse_mgf_device(bytes(synth.mgf_text(sp)))["mz"], sp["mz"])
    # search from text == search from arrays (targets first: candidate-first gather from the parsed peaks on the device, search at load)
    text = _dirty_mgf(sp)
    want_sp = host.parse_mgf(text)
    o = run_oracle(p, want_sp, tgt, prot)
    for staged in (False, True):
        a = Search(p)
        a.set_targets(*tgt)
        if staged:
            a.stage_reference(*prot)
        n = a.load_mgf(text)
        assert n == len(want_sp["prec_mz"])
        a.step2()
        a.build_reference() if staged else a.build_reference(*prot)
        a.step3(); a.step4(); a.rank()
        assert_psms_equal(a.psms(), o.psms())
    b = Search(p); assert b.load_mgf(text) == n; b.set_targets(*tgt); b.step2(); b.build_reference(*prot); b.step3(); b.step4(); b.rank()      # text first, targets later: full residency
    assert_psms_equal(b.psms(), o.psms())
    # empty text, text without blocks, malformed blocks
    assert g.load_mgf(b"") == 0 and g.load_mgf(b"# nothing here\n\n") == 0
    for bad in (b"BEGIN IONS\nPEPMASS=500\n100 1\nBEGIN IONS\nPEPMASS=600\n100 1\nEND IONS\n", b"END IONS\nBEGIN IONS\nEND IONS\n", b"BEGIN IONS\nPEPMASS=500\n100 1\n"):
        with pytest.raises(PepQueryError) as ei:
            g.load_mgf(bad)
        assert ei.value.code == _abi.PQ_ERR_BAD_ARG


def test_divide_by_three_shortcut_is_the_ieee_division():
    """The three-operation RN(x / 3) of the kernels (pq_common.h div3_rn: the m/z of triply charged fragments in K1, K2's cell lists and
    the ladder tables) against __ddiv_rn on the device: 8 values around each of 2^26 pseudo-random doubles, no mismatch."""
    g = Search(_abi.default_params())
    assert g.selftest(1, 1 << 26, seed=2022) == 0
    assert g.selftest(1, 1 << 20, seed=7) == 0
    g.close()
