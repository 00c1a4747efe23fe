"""GPU parity at the sizes BASELINE.json names (run with -m gpu on the B200).

  C1  100 peptides x 10 000 spectra x 20 000-protein database, 1000 shuffles — compared with the CPU oracle in FULL: every PSM
      field, the detail rows, the pair counts; hyperscore and MVH; Step 5 with the complete Unimod-sized PTM table.
  C2  10 000 peptides x 1 000 000 spectra — the GPU searches everything; the oracle searches a random 2 000-spectrum subset with the
      same targets and database and the reference-table mass window of the big run; every PSM of those spectra must agree.
The oracle runs with its hoisted cost model (pre-processing once per spectrum; tests/test_oracle_pins.py checks that it gives the
as-written results) on all host cores, so the file stays within a couple of minutes."""
import os

import numpy as np
import pytest

from pepquery_b200 import Search, _abi, synth
from test_gpu_parity import INT_FIELDS, REL, assert_psms_equal

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 8


def _oracle(p):
    from oracle import pq_oracle
    o = pq_oracle.Oracle(p, threads=THREADS)
    o.set_cost_model(True)
    return o


def _sorted_rows(rows, keys):
    return np.sort(rows, order=keys)


@pytest.fixture(scope="module")
def c1():
    prot = synth.proteins(20000)
    tgt = synth.targets(100)
    sp = synth.spectra(10000, tgt, prot)
    return prot, tgt, sp


@pytest.mark.parametrize("method", [1, 2])
def test_c1_full_size_against_the_oracle(c1, method):
    prot, tgt, sp = c1
    p = _abi.default_params(n_random=1000, score_method=method)
    g = Search(p)
    g.load_spectra(sp); g.set_targets(*tgt); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
    o = _oracle(p)
    o.load_spectra(sp); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
    got, want = g.psms(), o.psms()
    assert len(got) > 2500
    assert_psms_equal(got, want)
    assert np.array_equal(got["score"], want["score"])                      # same arithmetic order, same fdlibm log: identical bits
    c = g.counters()
    assert [c.pairs_step2, c.pairs_step3, c.pairs_step4] == [int(x) for x in o.pairs()[:3]]
    assert c.reference_forms == o.reference_forms()[1] > 2000000
    co = _sorted_rows(o.competitors(3), ["spectrum", "ref_form"]); cg = _sorted_rows(g.competitors(3), ["spectrum", "ref_form"])
    assert len(cg) > 1000 and np.array_equal(co["spectrum"], cg["spectrum"]) and np.array_equal(co["ref_form"], cg["ref_form"])
    assert np.array_equal(co["score"], cg["score"]) and np.array_equal(co["pep_mass"], cg["pep_mass"])
    (gr, gm), (wr, wm) = g.delta_scores(), o.delta_scores()
    assert np.array_equal(gr, wr) and np.array_equal(gm, wm)
    if method == 1:
        # Step 5: every Unimod specificity x shifted precursor window x free site (config C4's search on the C1 inputs)
        from pepquery_b200 import unimod
        ptms = unimod.load_ptms()
        g.step5(ptms); o.step5(ptms)
        got5, want5 = g.psms(), o.psms()
        assert_psms_equal(got5, want5)
        assert np.array_equal(got5["n_ptm"], want5["n_ptm"]) and (got5["n_ptm"] >= 0).sum() > 500
        assert g.counters().pairs_step5 == int(o.pairs()[3]) > 100000000
        keys = ["spectrum", "ref_form", "ptm", "ptm_site"]
        r5o = _sorted_rows(o.competitors(5), keys); r5g = _sorted_rows(g.competitors(5), keys)
        assert len(r5g) == len(r5o) > 0
        for k in keys:
            assert np.array_equal(r5o[k], r5g[k]), k
        assert np.allclose(r5o["score"], r5g["score"], rtol=REL, atol=0)
        (gr, gm), (wr, wm) = g.delta_scores(), o.delta_scores()
        assert np.allclose(gm, wm, rtol=REL, atol=1e-9)


def test_c2_scale_sampled_against_the_oracle():
    prot = synth.proteins(20000)
    tgt = synth.targets(10000)
    sp = synth.spectra(1000000, tgt, prot)
    p = _abi.default_params(n_random=1000)
    g = Search(p)
    g.load_spectra(sp); g.set_targets(*tgt); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
    got = g.psms()
    assert len(got) > 300000
    cg = g.competitors(3)
    # the oracle on a random subset, with the reference-table mass range of the full run (GeneratePeptideformWorker.java:40-52)
    lo, hi = float(got["exp_mass"].min()) - 250.0 - 5.0, float(got["exp_mass"].max()) + 250.0 + 5.0
    rng = np.random.default_rng(20240517)
    sel = np.sort(rng.choice(len(sp["prec_mz"]), 2000, replace=False))
    off = sp["peak_off"]
    counts = (off[sel + 1] - off[sel]).astype(np.int64)
    noff = np.zeros(len(sel) + 1, dtype=np.uint64); np.cumsum(counts, out=noff[1:])
    idx = np.repeat(off[sel].astype(np.int64) - noff[:-1].astype(np.int64), counts) + np.arange(int(noff[-1]), dtype=np.int64)
    sub = dict(prec_mz=sp["prec_mz"][sel].copy(), charge=sp["charge"][sel].copy(), peak_off=noff, mz=sp["mz"][idx].copy(), inten=sp["inten"][idx].copy())
    o = _oracle(p)
    o.set_reference_window(lo, hi)
    o.load_spectra(sub); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
    want = o.psms()
    want["spectrum"] = sel[want["spectrum"]].astype(np.int32)
    mine = got[np.isin(got["spectrum"], sel)]
    assert len(want) > 500
    assert_psms_equal(mine, want)
    assert np.array_equal(mine["score"], want["score"])
    assert g.counters().reference_forms == o.reference_forms()[1]
    co = o.competitors(3); co["spectrum"] = sel[co["spectrum"]].astype(np.int32)
    co = _sorted_rows(co, ["spectrum", "ref_form"]); cm = _sorted_rows(cg[np.isin(cg["spectrum"], sel)], ["spectrum", "ref_form"])
    assert len(cm) == len(co) > 100 and np.array_equal(co["spectrum"], cm["spectrum"]) and np.array_equal(co["ref_form"], cm["ref_form"])
    assert np.array_equal(co["score"], cm["score"])
    # the whole run: the decisions are internally consistent (A.10)
    top = got["rank"] == 1
    assert np.array_equal(got["valid"][top] == 1, got["n_db"][top] == 0)
    ev = got["n_random"] >= 0
    assert np.array_equal(got["p_value"][ev], (got["n_random"][ev] + 1.0) / (got["total_random"][ev] + 1.0))
