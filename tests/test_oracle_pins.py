"""Pins the CPU oracle to everything the reference tree (and the JDK spec) does pin — SURVEY.md §8c.

The reference ships no tests or golden vectors and its inner arithmetic lives in an un-vendored jar, so parity at the
compomics boundary is UNPINNED; these are the anchors that exist: java.util.Random (JDK spec), the worked shuffle
examples, the modification masses of resources/top_modifications.tsv (-> tests/golden/mod_masses.json), the formulae in
the Java source, and a regression fixture of the oracle's own outputs."""
import ctypes as C
import json
import math
import os
import random

import numpy as np
import pytest

from oracle import pq_oracle
from pepquery_b200 import _abi

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_java_random_known_answers():
    # SURVEY.md §8c "Derived known answers (from the JDK spec)"
    assert pq_oracle.java_random(12457, 0, 3).tolist() == [798512189, 819482266, 2027269868]
    assert pq_oracle.java_random(12457, 10, 8).tolist() == [9, 6, 8, 2, 9, 3, 0, 9]
    assert pq_oracle.java_random(2022, 3, 8).tolist() == [0, 1, 2, 0, 1, 0, 0, 1]


def test_java_random_against_python_restatement():
    def ref(seed, bound, n):
        s = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)
        out = []

        def nxt():
            nonlocal s
            s = (s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
            return s >> 17
        for _ in range(n):
            r = nxt()
            m = bound - 1
            if bound & m == 0:
                out.append((bound * r) >> 31)
            else:
                u = r
                while True:
                    r = u % bound
                    if u - r + m < (1 << 31):
                        break
                    u = nxt()
                out.append(r)
        return out
    for seed, bound in [(1, 7), (12457, 44), (2022, 2), (99, 1000003), (5, 1 << 30), (7, (1 << 30) + 1)]:
        assert pq_oracle.java_random(seed, bound, 200).tolist() == ref(seed, bound, 200)


def test_shuffle_worked_examples():
    # pg/PeptideInput.java:376-476; expected strings from SURVEY.md §8c
    assert pq_oracle.shuffles("PEPTLDEK", 5) == ["PEELDPTK", "TEPLEDPK", "EPTDEPLK", "PPTEDELK", "DEPLPTEK"]
    assert pq_oracle.shuffles("ELVLSQLGCK", 5) == ["ECLSLQGLVK", "LCQELGLVSK", "VLQLSLECGK", "LLGESLQCVK", "LEQGLCSVLK"]


def test_shuffle_properties():
    for seq in ["PEPTLDEK", "AAAAAAAK", "AAAAAABK", "MCMKWYQR", "ACDEFGHLKLMNPQSTVWYR"]:
        sh = pq_oracle.shuffles(seq, 1000)
        assert len(set(sh)) == len(sh) and seq not in sh            # de-duplicated, target excluded
        for s in sh:
            assert s[-1] == seq[-1] and sorted(s) == sorted(seq)    # C-terminal residue fixed, a permutation
    assert pq_oracle.shuffles("AAAAAAAK", 1000) == []               # only one arrangement: num is capped by nComb = 1
    # 7 arrangements -> num = min(1000, nComb) = 7 rounds only; duplicates and the target are dropped, not redrawn
    assert 1 <= len(pq_oracle.shuffles("AAAAAABK", 1000)) <= 6


def test_fdlibm_log_within_one_ulp_of_libm():
    L = pq_oracle.lib()
    rnd = random.Random(1)
    for _ in range(20000):
        x = math.exp(rnd.uniform(-40, 40))
        for f, g in ((L.orc_log, math.log), (L.orc_log10, math.log10)):
            a, b = f(x), g(x)
            assert a == b or abs(a - b) <= 2 * math.ulp(b)
    assert L.orc_log10(1.0) == 0.0 and L.orc_log10(100.0) == 2.0 and L.orc_log10(1e15) == 15.0
    assert L.orc_log10(float("inf")) == float("inf")


def test_hyperscore_spot_value_and_factorials():
    L = pq_oracle.lib()
    # b=3, y=5, sum=250 -> 4*(log10 250 + log10 3! + log10 5!)
    v = 4.0 * (L.orc_log10(250.0) + L.orc_log10_factorial(3) + L.orc_log10_factorial(5))
    assert v == 21.021090020413226
    L.orc_factorial_log.restype = C.c_double
    for n in range(0, 200):
        assert abs(L.orc_factorial_log(n) - math.lgamma(n + 1)) <= 1e-12 * max(1.0, math.lgamma(n + 1))
    for n in range(1, 61):
        assert abs(L.orc_log10_factorial(n) - math.lgamma(n + 1) / math.log(10)) < 1e-12 * max(1.0, n)


def test_masses_pinned_by_top_modifications():
    """Residue masses must be sums of the atom masses that resources/top_modifications.tsv pins."""
    gold = json.load(open(os.path.join(GOLD, "mod_masses.json")))
    carb = gold["Carbamidomethylation of C"]["mass"]        # C2 H3 N O
    ox = gold["Oxidation of M"]["mass"]                     # O
    acetyl = gold["Acetylation of K"]["mass"]               # C2 H2 O
    phospho = gold["Phosphorylation of S"]["mass"]          # H O3 P
    pyro_e = gold["Pyrolidone from E"]["mass"]              # -H2O
    pyro_q = gold["Pyrolidone from Q"]["mass"]              # -NH3
    O = ox
    h2o = -pyro_e
    H = (h2o - O) / 2
    Cc = 12.0
    N = -pyro_q - 3 * H
    assert abs((2 * Cc + 3 * H + N + O) - carb) < 1e-9 and abs((2 * Cc + 2 * H + O) - acetyl) < 1e-9
    P = phospho - H - 3 * O
    assert abs(P - 30.97376163) < 1e-7
    S = 31.97207100
    comp = {  # residue compositions (C, H, N, O, S)
        "G": (2, 3, 1, 1, 0), "A": (3, 5, 1, 1, 0), "S": (3, 5, 1, 2, 0), "P": (5, 7, 1, 1, 0), "V": (5, 9, 1, 1, 0),
        "T": (4, 7, 1, 2, 0), "C": (3, 5, 1, 1, 1), "L": (6, 11, 1, 1, 0), "I": (6, 11, 1, 1, 0), "N": (4, 6, 2, 2, 0),
        "D": (4, 5, 1, 3, 0), "Q": (5, 8, 2, 2, 0), "K": (6, 12, 2, 1, 0), "E": (5, 7, 1, 3, 0), "M": (5, 9, 1, 1, 1),
        "H": (6, 7, 3, 1, 0), "F": (9, 9, 1, 1, 0), "R": (6, 12, 4, 1, 0), "Y": (9, 9, 1, 2, 0), "W": (11, 10, 2, 1, 0)}
    L = pq_oracle.lib()
    for aa, (c, h, n, o, s) in comp.items():
        want = c * Cc + h * H + n * N + o * O + s * S
        assert abs(L.orc_residue_mass(aa.encode()) - want) < 2e-8, aa
    assert abs(_abi.CARBAMIDOMETHYL_C["delta"] - carb) == 0 and abs(_abi.OXIDATION_M["delta"] - ox) == 0
    assert abs(_abi.OXIDATION_M["neutral_loss"] - (Cc + 4 * H + O + S)) < 1e-8      # CH4OS
    assert abs(_abi.PHOSPHO_S["neutral_loss"] - (3 * H + P + 4 * O)) < 1e-7         # H3PO4


def _forms(o):
    arr, n = o.target_forms()
    return [(arr[i].seq.decode(), [(int(arr[i].mod_index[k]), int(arr[i].mod_site[k])) for k in range(arr[i].n_mods)], arr[i].mass) for i in range(n)]


def test_isoform_expansion_order_and_masses():
    # pg/PeptideInput.java:114-237: k-subsets in lexicographic order, fixed mods appended, fixed-only form last
    o = pq_oracle.Oracle(_abi.default_params())
    off = np.array([0, 6], dtype=np.uint32)
    res = np.frombuffer(b"MCMAMK", dtype=np.uint8).copy()
    o.set_targets(off, res)
    f = _forms(o)
    var = [[s for (m, s) in mods if m >= 0] for _, mods, _ in f]
    assert var == [[1], [3], [5], [1, 3], [1, 5], [3, 5], [1, 3, 5], []]
    assert all([(m, s) for (m, s) in mods if m < 0] == [(-1, 2)] for _, mods, _ in f)     # C2 carbamidomethyl everywhere
    L = pq_oracle.lib()
    base = 18.0105646837 + sum(L.orc_residue_mass(c.encode()) for c in "MCMAMK") + 57.02146372057
    for (_, mods, mass), v in zip(f, var):
        assert abs(mass - (base + 15.99491461956 * len(v))) < 1e-9
    # max_var_mods caps the subset size
    o2 = pq_oracle.Oracle(_abi.default_params(max_var_mods=1))
    o2.set_targets(off, res)
    assert [[s for (m, s) in mods if m >= 0] for _, mods, _ in _forms(o2)] == [[1], [3], [5], []]


def test_mod_replacement_on_shuffles():
    # pg/PeptideInput.java:485-557: fresh Random(2022) per call; nextInt(2) -> 0 on the first draw, so the first free M
    o = pq_oracle.Oracle(_abi.default_params())
    o.set_targets(np.array([0, 8], dtype=np.uint32), np.frombuffer(b"AMCMDEFK", dtype=np.uint8).copy())
    forms = _forms(o)                      # [M2], [M4], [M2,M4], []
    f = o.shuffle_form(0, "MACDMEFK")      # one oxidation to re-place on M at sites 1 and 5
    mods = [(int(f.mod_index[k]), int(f.mod_site[k])) for k in range(f.n_mods)]
    first = int(pq_oracle.java_random(2022, 2, 1)[0])
    # the C carbamidomethyl entry of the target form draws too (one free site -> nextInt(1)), then the fixed fill adds nothing
    assert mods == [(0, [1, 5][first]), (-1, 3)]
    f2 = o.shuffle_form(2, "MACDMEFK")     # two oxidations: both M sites end up modified
    assert sorted(s for (m, s) in [(int(f2.mod_index[k]), int(f2.mod_site[k])) for k in range(f2.n_mods)] if m == 0) == [1, 5]
    assert abs(f2.mass - forms[2][2]) < 1e-9


def test_digest_rules():
    # DigestProteinWorker.java:64-137 + A10: cleave after K/R unless P follows, <=2 missed, length 7..45, mass 500..10000, I->L
    o = pq_oracle.Oracle(_abi.default_params())
    sp = dict(prec_mz=np.array([500.0]), charge=np.array([2], dtype=np.int32), peak_off=np.array([0, 0], dtype=np.uint64),
              mz=np.zeros(0), inten=np.zeros(0))
    o.load_spectra(sp)
    o.set_targets(np.array([0, 8], dtype=np.uint32), np.frombuffer(b"AAAAAAAK", dtype=np.uint8).copy())
    prot = b"MAAAAAAKPEPTIDEKGGSGGSGR*AAAAAAAKWWWWWWWWR"
    o.build_reference(np.array([0, len(prot)], dtype=np.uint64), np.frombuffer(prot, dtype=np.uint8).copy())
    seqs = set(o.reference_sequences())
    assert "MAAAAAAKPEPTLDEK" in seqs                     # K before P is not a cleavage site; I -> L
    assert "GGSGGSGR" in seqs and "GGSGGSGRAAAAAAAK" in seqs and "MAAAAAAKPEPTLDEKGGSGGSGR" in seqs
    assert "AAAAAAAK" not in seqs                           # equals a target sequence -> removed (DatabaseInput.java:480-488)
    assert "AAAAAAAKWWWWWWWWR" in seqs and "WWWWWWWWR" in seqs
    assert all(7 <= len(s) <= 45 and "I" not in s and "*" not in s for s in seqs)


# the -e table of pg/DatabaseInput.java:145-267 restated independently of the oracle: (cleave after, cleave before, blocked by next)
ENZYMES = {1: ("RK", "", "P"), 2: ("RK", "", ""), 3: ("R", "", "P"), 4: ("R", "", ""), 5: ("", "R", ""), 6: ("E", "", ""), 7: ("K", "", "P"),
           8: ("K", "", ""), 9: ("", "K", ""), 10: ("", "D", ""), 11: ("", "DE", ""), 12: ("FYWL", "", "P"), 13: ("FYWL", "", ""),
           14: ("FL", "", ""), 15: ("M", "", ""), 16: ("", "AFILMV", ""), 17: ("", "RK", "")}


def python_digest(prot, enzyme, missed=2, min_len=7, max_len=45):
    after, before, block = ENZYMES[enzyme]
    prot = prot.upper().replace("*", "").replace("I", "L")
    cuts = [0] + [i + 1 for i in range(len(prot) - 1)
                  if (prot[i] in after and prot[i + 1] not in block) or prot[i + 1] in before] + [len(prot)]
    out = set()
    for a in range(len(cuts) - 1):
        for mc in range(missed + 1):
            if a + mc + 1 < len(cuts):
                pep = prot[cuts[a]:cuts[a + mc + 1]]
                if min_len <= len(pep) <= max_len:
                    out.add(pep)
    return out


@pytest.mark.parametrize("enzyme", sorted(ENZYMES))
def test_enzyme_table(enzyme):
    """Every enzyme of the reference's -e list (compomics Enzyme.isCleavageSite semantics, A10)."""
    from pepquery_b200 import synth
    poff, pres = synth.proteins(12)
    prots = [bytes(pres[poff[i]:poff[i + 1]]).decode() for i in range(len(poff) - 1)]
    o = pq_oracle.Oracle(_abi.default_params(enzyme=enzyme, min_mass=0.0, max_mass=1e9))
    o.load_spectra(dict(prec_mz=np.array([500.0]), charge=np.array([2], dtype=np.int32), peak_off=np.array([0, 0], dtype=np.uint64),
                        mz=np.zeros(0), inten=np.zeros(0)))
    o.set_targets(np.array([0, 8], dtype=np.uint32), np.frombuffer(b"AAAAAAAK", dtype=np.uint8).copy())
    o.build_reference(poff, pres)
    want = set()
    for pr in prots:
        want |= python_digest(pr, enzyme)
    want.discard("AAAAAAAK")
    assert set(o.reference_sequences()) == want and len(want) > 50


def test_no_enzyme_is_every_substring():
    """-e 0 (NoEnzyme, pg/DatabaseInput.java:117-121, CleavageParameter.unSpecific): every stretch of a protein within the length
    bounds (the mass window opened wide here), I -> L folded, '*' dropped, minus the targets."""
    from pepquery_b200 import synth
    poff, pres = synth.proteins(3)
    prots = [bytes(pres[poff[i]:poff[i + 1]]).decode() for i in range(len(poff) - 1)]
    o = pq_oracle.Oracle(_abi.default_params(enzyme=0, min_len=7, max_len=12, min_mass=0.0, max_mass=1e9))
    o.load_spectra(dict(prec_mz=np.array([500.0]), charge=np.array([2], dtype=np.int32), peak_off=np.array([0, 0], dtype=np.uint64),
                        mz=np.zeros(0), inten=np.zeros(0)))
    o.set_targets(np.array([0, 8], dtype=np.uint32), np.frombuffer(b"AAAAAAAK", dtype=np.uint8).copy())
    o.build_reference(poff, pres)
    want = set()
    for pr in prots:
        pr = pr.upper().replace("*", "").replace("I", "L")
        for a in range(len(pr)):
            for ln in range(7, 13):
                if a + ln <= len(pr):
                    want.add(pr[a:a + ln])
    want.discard("AAAAAAAK")
    assert set(o.reference_sequences()) == want and len(want) > 1000


def test_oracle_regression_fixture():
    """The oracle must keep reproducing its committed outputs (tests/golden/oracle_vectors.json, tools/make_golden.py)."""
    from pepquery_b200 import synth
    gold = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))
    prot = synth.proteins(300); tgt = synth.targets(25); sp = synth.spectra(700, tgt, prot)
    for method in (1, 2):
        p = _abi.default_params(n_random=100, score_method=method)
        o = pq_oracle.Oracle(p, threads=4)
        o.load_spectra(sp); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
        ps = o.psms()
        g = gold["method%d" % method]
        for f in ("spectrum", "form", "count_b", "count_y", "n_db", "total_db", "n_random", "total_random", "valid", "rank", "confident"):
            assert ps[f].tolist() == g[f], (method, f)
        assert ps["score"].tolist() == g["score"] and ps["p_value"].tolist() == g["p_value"]
        assert [int(x) for x in o.pairs()] == g["pairs"]
        assert o.delta_scores()[0].tolist() == g["ref_delta_score"]
    toff, tres, want = synth.planted_targets(prot, np.random.default_rng(7))
    assert pq_oracle.find_in_proteome(toff, tres, prot[0], prot[1]).tolist() == gold["novelty_seed7"] == want.tolist()


def test_preprocessing_invariants(small_inputs):
    prot, tgt, sp = small_inputs
    o = pq_oracle.Oracle(_abi.default_params())
    o.load_spectra(sp)
    for si in range(0, 700, 37):
        mz, val, n, lim = o.preprocess(si, 1)
        assert 1 <= n <= 50 and np.all(np.diff(mz) > 0) and np.all(mz > 150.0)
        assert val.max() == 100.0 and val.min() >= 5.0 - 1e-9         # normalised to 100; >= 5 % of the maximum survived
        raw = sp["inten"][int(sp["peak_off"][si]):int(sp["peak_off"][si + 1])]
        assert raw.min() <= lim <= np.percentile(raw, 6)
        mz2, cls, n2, _ = o.preprocess(si, 2)
        if n2 >= 7:
            k = n2 // 7
            assert (cls == 0).sum() == k and (cls == 1).sum() == 2 * k and np.all((mz2 >= 200) & (mz2 <= 2000))


@pytest.mark.parametrize("threads", [1, 4])
def test_oracle_thread_count_does_not_change_results(small_inputs, threads):
    prot, tgt, sp = small_inputs
    o = pq_oracle.Oracle(_abi.default_params(n_random=50), threads=threads)
    o.load_spectra(sp); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
    ps = o.psms()
    gold = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))["method1"]
    assert ps["spectrum"].tolist() == gold["spectrum"] and ps["score"].tolist() == gold["score"] and ps["n_db"].tolist() == gold["n_db"]


def test_delta_scores_follow_the_detail_rows(small_inputs):
    """ref_delta_score / mod_delta_score (PeptideSearchMT.add_delta_score, pg/PeptideSearchMT.java:1365-1475) restated in
    Python from the rows the oracle reports for detail.txt: score minus the spectrum's best row, the plain score without one."""
    prot, tgt, sp = small_inputs
    o = pq_oracle.Oracle(_abi.default_params(n_random=20), threads=4)
    o.load_spectra(sp); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
    ps, rows = o.psms(), o.competitors(3)
    best = {}
    for r in rows:
        s = int(r["spectrum"])
        if s not in best or best[s] < r["score"]:
            best[s] = float(r["score"])
    want = np.array([p["score"] - best[int(p["spectrum"])] if int(p["spectrum"]) in best else p["score"] for p in ps])
    ref, mod = o.delta_scores()
    assert np.array_equal(ref, want) and np.array_equal(mod, ps["score"]) and len(best) > 10
    assert (ref < 0).sum() == (ps["n_db"] > 0).sum() + ((ref < 0) & (ps["n_db"] == 0)).sum()   # a better reference match => negative delta
    assert np.all(ref[ps["n_db"] > 0] <= 0)


def test_novelty_filter_against_python_substring_search():
    """Step-1 novelty filter (pg/InputProcessor.java:416-427, pg/PepMapping.java:60-107): the oracle against Python's `in`."""
    from pepquery_b200 import synth
    prot = synth.proteins(120)
    prot = (prot[0], prot[1].copy())
    prot[1][::53] |= 0x20                                            # some lower-case residues in the database
    off, res, want = synth.planted_targets(prot, np.random.default_rng(5))
    got = pq_oracle.find_in_proteome(off, res, prot[0], prot[1])
    assert np.array_equal(got, want) and 40 < want.sum() < len(want)


def test_divide_by_three_shortcut_in_exact_arithmetic():
    """pq_common.h div3_rn — RN(x r), one exact remainder, one fused correction with r = RN(1/3) — returns RN(x / 3).  Replayed with
    exact rationals (float(Fraction) rounds to nearest-even once, which is what a fused multiply-add does) on random significands over
    the exponents fragment masses can have, on the neighbours of multiples of three, and on the extremes of the significand."""
    from fractions import Fraction
    import random
    import struct

    r = 0.33333333333333331
    assert r == float(Fraction(1, 3)) and Fraction(r) == Fraction(2 ** 54 - 1, 3 * 2 ** 54)

    def fma(a, b, c):
        return float(Fraction(a) * Fraction(b) + Fraction(c))

    def div3(x):
        q0 = float(Fraction(x) * Fraction(r))
        rem = fma(-3.0, q0, x)
        assert Fraction(rem) == Fraction(x) - 3 * Fraction(q0)          # the remainder is exact
        return fma(rem, r, q0)

    def nudge(x, d):
        return struct.unpack("<d", struct.pack("<q", struct.unpack("<q", struct.pack("<d", x))[0] + d))[0]

    rnd = random.Random(2022)
    xs = []
    for _ in range(30000):
        x = rnd.random() + 1.0
        xs.append(x * 2.0 ** rnd.randint(-8, 15))
    for _ in range(10000):
        y = 3.0 * float(rnd.randint(1, 2 ** 52))
        xs += [nudge(y, -1), y, nudge(y, 1)]
    for e in range(-8, 16):
        for m in (1.0, nudge(1.0, 1), nudge(2.0, -1), 1.5, nudge(1.5, -1), nudge(1.5, 1)):
            xs.append(m * 2.0 ** e)
    for x in xs:
        assert div3(x) == float(Fraction(x) / 3), x
