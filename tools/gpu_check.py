"""Development aid: run the oracle and the CUDA path on the same synthetic search and print every disagreement.
(The gated version of these checks lives in tests/test_gpu_parity.py.)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pepquery_b200 import Search, _abi, synth  # noqa: E402
from oracle import pq_oracle  # noqa: E402


def compare(n_prot=2000, n_tgt=100, n_spec=5000, n_random=200, method=1, verbose=True, **over):
    prot = synth.proteins(n_prot)
    tg = synth.targets(n_tgt)
    sp = synth.spectra(n_spec, tg, prot)
    p = _abi.default_params(n_random=n_random, score_method=method, **over)
    o = pq_oracle.Oracle(p, threads=8)
    o.load_spectra(sp); o.set_targets(*tg)
    g = Search(p)
    g.load_spectra(sp); g.set_targets(*tg)
    bad = 0
    fo, nfo = o.target_forms(); fg, nfg = g.target_forms()
    same_forms = nfo == nfg and all(bytes(fo[i]) == bytes(fg[i]) for i in range(nfo))
    print("target forms", nfo, nfg, "identical" if same_forms else "DIFFER"); bad += not same_forms
    t = time.time(); no = o.step2(); to = time.time() - t
    t = time.time(); ng = g.step2(); tgp = time.time() - t
    print("step2 psms oracle %d (%.3fs) gpu %d (%.3fs)" % (no, to, ng, tgp))
    po, pg = o.psms(), g.psms()
    bad += cmp_psms(po, pg, ["score", "count_b", "count_y", "exp_mass", "pep_mass"])
    o.build_reference(*prot); g.build_reference(*prot)
    ro, nro = o.reference_forms(); rg, nrg = g.reference_forms()
    same_ref = nro == nrg and bytes(ro)[:nro * C_sizeof_form()] == bytes(rg)[:nrg * C_sizeof_form()]
    print("reference forms", nro, nrg, "identical" if same_ref else "DIFFER"); bad += not same_ref
    t = time.time(); o.step3(); to = time.time() - t
    t = time.time(); g.step3(); tgp = time.time() - t
    print("step3 oracle %.3fs gpu %.3fs" % (to, tgp))
    po, pg = o.psms(), g.psms()
    bad += cmp_psms(po, pg, ["n_db", "total_db", "valid"])
    co, cg = o.competitors(3), g.competitors(3)
    co = np.sort(co, order=["spectrum", "ref_form"]); cg = np.sort(cg, order=["spectrum", "ref_form"])
    okc = len(co) == len(cg) and np.array_equal(co["spectrum"], cg["spectrum"]) and np.array_equal(co["ref_form"], cg["ref_form"]) \
        and np.allclose(co["score"], cg["score"], rtol=1e-9, atol=0)
    print("detail rows", len(co), len(cg), "ok" if okc else "DIFFER"); bad += not okc
    t = time.time(); o.step4(); to = time.time() - t
    t = time.time(); g.step4(); tgp = time.time() - t
    print("step4 oracle %.3fs gpu %.3fs" % (to, tgp))
    o.rank(); g.rank()
    po, pg = o.psms(), g.psms()
    bad += cmp_psms(po, pg, ["n_random", "total_random", "p_value", "rank", "confident"])
    print("pairs oracle", o.pairs(), "gpu", [g.counters().pairs_step2, g.counters().pairs_step3, g.counters().pairs_step4])
    c = g.counters()
    print("gpu ms: prepare %.3f window %.3f score %.3f shuffle %.3f" % (c.ms_prepare, c.ms_window, c.ms_score, c.ms_shuffle))
    # shuffle strings of a few targets
    toff, tres = tg
    nsh = 0
    for ts in range(min(n_tgt, 20)):
        seq = bytes(tres[toff[ts]:toff[ts + 1]]).decode()
        sg = g.shuffles(ts, len(seq))
        if not sg:
            continue
        so = pq_oracle.shuffles(seq, n_random)
        nsh += 1
        if so != sg:
            bad += 1
            print("shuffles differ for", seq, len(so), len(sg), so[:2], sg[:2])
    print("shuffle lists compared:", nsh)
    print("TOTAL DISAGREEMENTS:", bad)
    return bad


def C_sizeof_form():
    import ctypes
    return ctypes.sizeof(_abi.pq_form)


def cmp_psms(po, pg, fields):
    bad = 0
    ko = list(zip(po["spectrum"], po["form"])); kg = list(zip(pg["spectrum"], pg["form"]))
    if ko != kg:
        so, sg = set(ko), set(kg)
        print("  PSM key sets differ: only oracle %d, only gpu %d" % (len(so - sg), len(sg - so)), sorted(so - sg)[:5], sorted(sg - so)[:5])
        return 1
    for f in fields:
        a, b = po[f], pg[f]
        if a.dtype.kind == "f":
            if f == "score":
                exact = np.sum(a == b)
                ok = np.allclose(a, b, rtol=1e-6, atol=0)
                print("  %-12s rel<=1e-6: %s  bit-identical: %d/%d  max rel %.3g" % (f, ok, exact, len(a), np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)) if len(a) else 0))
            else:
                ok = np.array_equal(a, b)
                print("  %-12s identical: %s" % (f, ok))
        else:
            ok = np.array_equal(a, b)
            print("  %-12s identical: %s" % (f, ok) + ("" if ok else "  mismatches %d e.g. %s" % (np.sum(a != b), [(int(x), int(y)) for x, y in zip(a[a != b][:5], b[a != b][:5])])))
        bad += not ok
    return bad


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--prot", type=int, default=2000); ap.add_argument("--tgt", type=int, default=100)
    ap.add_argument("--spec", type=int, default=5000); ap.add_argument("--nrand", type=int, default=200)
    ap.add_argument("--method", type=int, default=1)
    a = ap.parse_args()
    sys.exit(1 if compare(a.prot, a.tgt, a.spec, a.nrand, a.method) else 0)
