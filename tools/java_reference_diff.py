#!/usr/bin/env python
"""SURVEY.md Appendix A.11 as a runnable recipe: "if a JVM ever becomes available".

  write   DIR            synthetic inputs in the reference's own file formats (syn.mgf, syn.fasta, targets.txt) plus the oracle's
                         result tables (oracle_psm_rank.tsv, oracle_detail.tsv, oracle_ptm.tsv) for the same search
  diff    DIR JAVA_OUT   compare the reference's own output directory (psm_rank.txt, detail.txt, ptm.txt written by
                         `java -jar pepquery-2.0.4.jar -i targets.txt -ms syn.mgf -db syn.fasta -n 1000 -o JAVA_OUT`) with the
                         oracle tables: integer columns exact, scores within 1e-6 relative (BASELINE.json north_star)

Nothing here runs on the product path: the oracle is test infrastructure.  No JVM exists in the build image or on the GPU
box (bench.py probes it), so `diff` has never seen real reference output; the column names it reads are the headers the
reference writes (psm.txt pg/PeptideSearchMT.java:1634, +rank pg/RankPSM.java:80, +n_ptm/confident :1007, detail.txt
pg/DatabaseInput.java:714, ptm.txt OpenModificationSearch/ModificationDB.java:1433) and the modification string is
name@site[%.4f] joined by ';' (ModificationDB.getModificationString :1695-1708).  A disagreement pattern names the
compomics assumption (A1-A13) to flip in oracle/pq_oracle.cpp.
"""
import argparse
import csv
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MOD_NAMES = {1: "Carbamidomethylation of C", 2: "Oxidation of M"}      # -fixMod 1 / -varMod 2 of the reference's printPTM table


def form_string(f, params):
    """pq_form -> (sequence, modification string in the reference's text form)."""
    parts = []
    for k in range(f.n_mods):
        idx = int(f.mod_index[k])
        m = params.var[idx] if idx >= 0 else params.fixed[-idx - 1]
        parts.append("%s@%d[%.4f]" % (MOD_NAMES.get(int(m.id), "mod%d" % m.id), int(f.mod_site[k]), m.delta))
    return f.seq.decode(), (";".join(parts) if parts else "-")


def write(args):
    from pepquery_b200 import _abi, synth, host
    from pepquery_b200.unimod import load_ptms
    from oracle import pq_oracle
    os.makedirs(args.dir, exist_ok=True)
    prot = synth.proteins(args.proteins)
    tgt = synth.targets(args.targets)
    sp = synth.spectra(args.spectra, tgt, prot)
    params = _abi.default_params(n_random=args.n_random, score_method=args.method, fast=args.fast)
    # ---- inputs, reference formats ----
    poff, pres = prot
    with open(os.path.join(args.dir, "syn.fasta"), "w") as f:
        for i in range(len(poff) - 1):
            s = bytes(pres[int(poff[i]):int(poff[i + 1])]).decode()
            f.write(">SYN%05d\n" % i)
            for a in range(0, len(s), 60):
                f.write(s[a:a + 60] + "\n")
    toff, tres = tgt
    with open(os.path.join(args.dir, "targets.txt"), "w") as f:
        for i in range(len(toff) - 1):
            f.write(bytes(tres[int(toff[i]):int(toff[i + 1])]).decode() + "\n")
    titles = [str(i + 1) for i in range(len(sp["prec_mz"]))]        # SpectraInput.java:161-166: title = 1-based ordinal
    with open(os.path.join(args.dir, "syn.mgf"), "w") as f:
        f.write(host.write_mgf(sp, titles))
    # ---- the oracle's tables ----
    o = pq_oracle.Oracle(params, threads=args.threads)
    o.load_spectra(sp); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
    if args.step5:
        o.step5(load_ptms())
        o.rank()
    psm = o.psms()
    tf, ntf = o.target_forms()
    rf, nrf = o.reference_forms()
    ref_delta, mod_delta = o.delta_scores()
    with open(os.path.join(args.dir, "oracle_psm_rank.tsv"), "w") as f:
        f.write("peptide\tmodification\tspectrum_title\tcharge\texp_mass\tisotope_error\tpep_mass\tscore\tn_db\ttotal_db\tn_random\ttotal_random\tpvalue\trank\tn_ptm\tconfident\tref_delta_score\tmod_delta_score\n")
        for k, r in enumerate(psm):
            seq, mod = form_string(tf[int(r["form"])], params)
            f.write("%s\t%s\t%s\t%d\t%r\t%d\t%r\t%r\t%d\t%d\t%d\t%d\t%r\t%d\t%d\t%s\t%r\t%r\n" % (
                seq, mod, titles[int(r["spectrum"])], r["charge"], float(r["exp_mass"]), r["isotope_error"], float(r["pep_mass"]), float(r["score"]),
                r["n_db"], r["total_db"], r["n_random"], r["total_random"], float(r["p_value"]), r["rank"], r["n_ptm"],
                "Yes" if r["confident"] else "No", float(ref_delta[k]), float(mod_delta[k])))
    for step, name in ((3, "oracle_detail.tsv"), (5, "oracle_ptm.tsv")):
        if step == 5 and not args.step5:
            continue
        rows = o.competitors(step)
        with open(os.path.join(args.dir, name), "w") as f:
            f.write("spectrum_title\tpeptide\tmodification\tpep_mass\tscore\tptm\tptm_site\n")
            for r in rows:
                seq, mod = form_string(rf[int(r["ref_form"])], params)
                f.write("%s\t%s\t%s\t%r\t%r\t%d\t%d\n" % (titles[int(r["spectrum"])], seq, mod, float(r["pep_mass"]), float(r["score"]), r["ptm"], r["ptm_site"]))
    print("wrote %s: %d spectra, %d targets, %d proteins; oracle: %d PSMs" % (args.dir, len(titles), len(toff) - 1, len(poff) - 1, len(psm)))
    print("now run:  java -jar pepquery-2.0.4.jar -i %s/targets.txt -ms %s/syn.mgf -db %s/syn.fasta -fixMod 1 -varMod 2 -n %d -m %d%s -o JAVA_OUT"
          % (args.dir, args.dir, args.dir, args.n_random, args.method, " -fast" if args.fast else ""))
    print("then:     python tools/java_reference_diff.py diff %s JAVA_OUT" % args.dir)


def read_table(path):
    with open(path) as f:
        return list(csv.DictReader(f, delimiter="\t"))


def mod_key(s):
    """Order-free key of a modification string (the reference lists variable before fixed entries; site and mass decide)."""
    if s in ("-", "", None):
        return ()
    out = []
    for p in s.split(";"):
        site = p[p.index("@") + 1:p.index("[")]
        mass = p[p.index("[") + 1:p.index("]")]
        out.append((int(site), mass))
    return tuple(sorted(out))


def diff(args):
    bad = 0
    want = {(r["peptide"], mod_key(r["modification"]), r["spectrum_title"]): r for r in read_table(os.path.join(args.dir, "oracle_psm_rank.tsv"))}
    got_rows = read_table(os.path.join(args.java_out, "psm_rank.txt"))
    got = {(r["peptide"], mod_key(r["modification"]), r["spectrum_title"]): r for r in got_rows}
    for k in sorted(set(want) - set(got)):
        print("PSM only in the oracle:", k); bad += 1
    for k in sorted(set(got) - set(want)):
        print("PSM only in the reference:", k); bad += 1
    ints = ["charge", "isotope_error", "n_db", "total_db", "n_random", "total_random", "rank", "n_ptm"]
    reals = ["score", "pvalue", "exp_mass", "pep_mass", "ref_delta_score", "mod_delta_score"]
    for k in sorted(set(want) & set(got)):
        w, g = want[k], got[k]
        for c in ints:
            if c in g and g[c] not in ("", "NA") and int(float(g[c])) != int(w[c]):
                print("%s %s: reference %s, oracle %s" % (k, c, g[c], w[c])); bad += 1
        for c in reals:
            if c in g and g[c] not in ("", "NA", "-"):
                a, b = float(g[c]), float(w[c])
                if not np.isclose(a, b, rtol=1e-6, atol=0.0 if c == "score" else 1e-9):
                    print("%s %s: reference %r, oracle %r" % (k, c, a, b)); bad += 1
        if "confident" in g and g["confident"] != w["confident"]:
            print("%s confident: reference %s, oracle %s" % (k, g["confident"], w["confident"])); bad += 1
    for java_name, ours in (("detail.txt", "oracle_detail.tsv"), ("ptm.txt", "oracle_ptm.tsv")):
        jp, op = os.path.join(args.java_out, java_name), os.path.join(args.dir, ours)
        if not (os.path.exists(jp) and os.path.exists(op)):
            print("skipped %s (missing on one side)" % java_name)
            continue
        w = {(r["spectrum_title"], r["peptide"], mod_key(r["modification"])): float(r["score"]) for r in read_table(op)}
        g = {(r["spectrum_title"], r["peptide"], mod_key(r["modification"])): float(r["score"]) for r in read_table(jp)}
        for k in sorted(set(w) ^ set(g)):
            print("%s row only in the %s: %s" % (java_name, "oracle" if k in w else "reference", k)); bad += 1
        for k in sorted(set(w) & set(g)):
            if not np.isclose(w[k], g[k], rtol=1e-6, atol=0.0):
                print("%s %s score: reference %r, oracle %r" % (java_name, k, g[k], w[k])); bad += 1
    print("%d PSMs compared, %d differences" % (len(set(want) & set(got)), bad))
    return 1 if bad else 0


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    w = sub.add_parser("write")
    w.add_argument("dir")
    w.add_argument("--spectra", type=int, default=10000)
    w.add_argument("--targets", type=int, default=100)
    w.add_argument("--proteins", type=int, default=20000)
    w.add_argument("--n-random", type=int, default=1000)
    w.add_argument("--method", type=int, default=1)
    w.add_argument("--fast", type=int, default=0)
    w.add_argument("--step5", action="store_true")
    w.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    d = sub.add_parser("diff")
    d.add_argument("dir")
    d.add_argument("java_out")
    a = ap.parse_args()
    sys.exit(write(a) if a.cmd == "write" else diff(a))


if __name__ == "__main__":
    main()
