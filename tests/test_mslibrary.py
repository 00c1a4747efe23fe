"""The on-disk MS/MS library path: writer in the reference's text form, pq_read_ms_library, library_mode search."""
import os

import numpy as np
import pytest

from pepquery_b200 import _abi, host, shard, synth


def _library(tmp_path, sp):
    d = str(tmp_path / "lib")
    titles = host.build_ms_library(sp, d)
    return d, titles


def test_library_round_trip(small_inputs, tmp_path):
    """One MGF per 0.1-Da bin (index/IndexWorker.java:422-486); reading every bin back gives every spectrum, bit for bit
    (the synthetic peaks are %.4f / %.2f numbers, the precursor m/z is written with the shortest round-tripping text)."""
    prot, tgt, sp = small_inputs
    d, titles = _library(tmp_path, sp)
    mass = shard.precursor_mass(sp["prec_mz"], sp["charge"].astype(np.float64))
    bins = np.floor(mass * 10 + 0.5).astype(np.int64)
    assert sorted(int(f[:-4]) for f in os.listdir(d)) == sorted(set(bins.tolist()))
    got, gt = host.read_ms_library(d, 0.0, 1e5)
    assert len(gt) == len(titles) and sorted(gt) == sorted(titles)
    idx = np.array([int(t.split(":")[1]) - 1 for t in gt])
    assert np.array_equal(bins[idx], np.sort(bins))                          # bin by bin
    assert np.array_equal(got["prec_mz"], sp["prec_mz"][idx]) and np.array_equal(got["charge"], sp["charge"][idx])
    off = sp["peak_off"]
    for k in range(0, len(idx), 7):
        a, b = int(got["peak_off"][k]), int(got["peak_off"][k + 1]); s = idx[k]
        assert np.array_equal(got["mz"][a:b], sp["mz"][int(off[s]):int(off[s + 1])]) and np.array_equal(got["inten"][a:b], sp["inten"][int(off[s]):int(off[s + 1])])
    # a mass range = one shard of a sweep: exactly the spectra inside it
    lo, hi = np.quantile(mass, 0.3), np.quantile(mass, 0.6)
    part, pt = host.read_ms_library(d, float(lo), float(hi))
    want = {titles[i] for i in np.nonzero((mass >= lo) & (mass <= hi))[0]}
    assert set(pt) == want and len(pt) == len(want)
    # empty range, missing directory
    none, nt = host.read_ms_library(d, 5.0, 6.0)
    assert len(nt) == 0 and len(none["mz"]) == 0
    nothing, _ = host.read_ms_library(str(tmp_path / "absent"), 0.0, 100.0)
    assert len(nothing["prec_mz"]) == 0


def test_library_ignores_chargeless_and_unterminated_blocks(tmp_path):
    d = tmp_path / "lib2"; d.mkdir()
    m = 1000.04; mz = (m + 2 * 1.007276466812) / 2
    text = ("BEGIN IONS\nTITLE=a:1:2\nPEPMASS=%r 9\nCHARGE=2+\n100.5 1\n200.25 2\nEND IONS\n"
            "BEGIN IONS\nTITLE=nocharge\nPEPMASS=%r\n100.5 1\nEND IONS\n"
            "BEGIN IONS\nTITLE=broken\nPEPMASS=%r\nCHARGE=2+\n111.0 5\n"
            "BEGIN IONS\nTITLE=b:2:2\nPEPMASS=%r\nCHARGE=2+\n300.125 3\nEND IONS\n") % (mz, mz, mz, mz)
    (d / ("%d.mgf" % round(m * 10))).write_text(text)
    sp, titles = host.read_ms_library(str(d), 999.0, 1001.0)
    assert titles == ["a:1:2", "b:2:2"] and sp["mz"].tolist() == [100.5, 200.25, 300.125] and sp["peak_off"].tolist() == [0, 2, 3]
    # the same text through pq_parse_mgf: the unterminated block's peaks belong to nobody
    p = host.parse_mgf(text)
    assert p["peak_off"].tolist() == [0, 2, 3, 4] and p["charge"].tolist() == [2, 0, 2] and p["mz"].tolist() == [100.5, 200.25, 100.5, 300.125]


@pytest.mark.gpu
@pytest.mark.parametrize("over", [dict(), dict(min_peaks=11), dict(tol=0.3, tol_ppm=0), dict(score_method=2)], ids=["default", "minpeaks11", "tol0.3Da", "mvh"])
def test_library_mode_search(tmp_path, over):
    """Spectra read from the library, searched with the rules of msio/MsLibrarySearchWorker.java:42-153 (library_mode): n_peaks >=
    min_peaks, |mass error| <= tol without the bucket rule.  GPU == oracle; and where those rules change nothing, == the plain search."""
    from oracle import pq_oracle
    from pepquery_b200 import Search
    from test_gpu_parity import assert_psms_equal
    prot = synth.proteins(400); tgt = synth.targets(30); sp0 = synth.spectra(1200, tgt, prot)
    d, _ = _library(tmp_path, sp0)
    sp, titles = host.read_ms_library(d, 0.0, 1e5)
    assert len(titles) == 1200
    p = _abi.default_params(n_random=60, library_mode=1, **over)

    def run(cls_factory, params, spectra):
        g = cls_factory(params)
        g.load_spectra(spectra); g.set_targets(*tgt); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
        return g
    g = run(Search, p, sp)
    o = run(lambda q: pq_oracle.Oracle(q, threads=8), p, sp)
    got, want = g.psms(), o.psms()
    assert len(got) > 100
    assert_psms_equal(got, want)
    c = g.counters()
    assert [c.pairs_step2, c.pairs_step3, c.pairs_step4] == [int(x) for x in o.pairs()[:3]]
    plain = run(Search, _abi.default_params(n_random=60, **over), sp).psms()
    if not over or over.get("score_method"):
        assert got.tobytes() == plain.tobytes()                   # 10 ppm, every spectrum has > min_peaks peaks: same search
    elif "tol" in over:
        assert len(got) > len(plain)                              # forms two buckets away (0.15 .. 0.3 Da) count in a library search
    else:
        assert len(got) >= len(plain)                             # spectra with exactly min_peaks peaks count
    with pytest.raises(Exception):
        bad = {k: v.copy() for k, v in sp.items()}; bad["charge"][3] = 0
        Search(p).load_spectra(bad)
