"""Generates the committed fixtures under tests/golden/.  Run in the build container only (it reads the read-only
reference tree at /root/reference, which does not exist on the GPU box):

    python tools/make_golden.py

  mod_masses.json        name -> monoisotopic mass, from the reference's resources/top_modifications.tsv (the only numeric
                         ground truth the reference tree ships; it pins the compomics atom masses)
  ../pepquery_b200/data/unimod_ptms.tsv
                         the Step-5 PTM specificities derived from resources/unimod.xml with the filters of
                         ModificationDB.importPTMsFromUnimod (OpenModificationSearch/ModificationDB.java:334-533)
  oracle_vectors.json    outputs of the CPU oracle on a small seeded search (regression anchor for the oracle itself and the
                         fixture the GPU tests replay without needing the oracle build)
"""
import json
import os
import re
import sys
import xml.etree.ElementTree as ET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/src/main/resources"
OUT = os.path.join(ROOT, "tests", "golden")


def mod_masses():
    rows = open(os.path.join(REF, "top_modifications.tsv")).read().strip().split("\n")
    head = rows[0].split("\t")
    out = {}
    for r in rows[1:]:
        f = dict(zip(head, r.split("\t")))
        out[f["mod_name"]] = {"id": int(f["mod_id"]), "mass": float(f["mod_mass"]), "type": f["mod_type"], "unimod": f["unimod_accession"]}
    json.dump(out, open(os.path.join(OUT, "mod_masses.json"), "w"), indent=1, sort_keys=True)
    return out


def unimod_ptms():
    """One row per (modification, specificity): id, name, residue-or-terminus, position, classification, mono mass.
    Filters as in importPTMsFromUnimod: |mass| window is applied at search time (leftMass/rightMass = -250/250,
    ModificationDB.java:61-62); 'AA substitution' specificities are skipped (:390-392); hidden flags are ignored."""
    ns = {"u": "http://www.unimod.org/xmlns/schema/unimod_2"}
    tree = ET.parse(os.path.join(REF, "unimod.xml"))
    rows = []
    for mod in tree.getroot().find("u:modifications", ns).findall("u:mod", ns):
        title = mod.get("title")
        rid = int(mod.get("record_id"))
        mono = float(mod.find("u:delta", ns).get("mono_mass"))
        for sp in mod.findall("u:specificity", ns):
            cls = sp.get("classification")
            rows.append((rid, title, sp.get("site"), sp.get("position"), cls, mono))
    with open(os.path.join(ROOT, "pepquery_b200", "data", "unimod_ptms.tsv"), "w") as f:
        f.write("record_id\ttitle\tsite\tposition\tclassification\tmono_mass\n")
        for r in rows:
            f.write("%d\t%s\t%s\t%s\t%s\t%r\n" % r)
    return rows


def oracle_vectors():
    import numpy as np
    from oracle import pq_oracle
    from pepquery_b200 import _abi, synth
    pq_oracle.build(); synth.build()
    prot = synth.proteins(300); tgt = synth.targets(25); sp = synth.spectra(700, tgt, prot)
    out = {"inputs": {"proteins": 300, "targets": 25, "spectra": 700, "seeds": "defaults of pepquery_b200.synth"}}
    for method in (1, 2):
        p = _abi.default_params(n_random=100, score_method=method)
        o = pq_oracle.Oracle(p, threads=4)
        o.load_spectra(sp); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
        ps = o.psms()
        out["method%d" % method] = {f: ps[f].tolist() for f in ps.dtype.names}
        out["method%d" % method]["pairs"] = [int(x) for x in o.pairs()]
        _, nref = o.reference_forms()
        out["method%d" % method]["reference_forms"] = nref
        out["method%d" % method]["ref_delta_score"] = o.delta_scores()[0].tolist()
        o.close()
    out["java_random_12457_next31"] = pq_oracle.java_random(12457, 0, 3).tolist()
    out["java_random_12457_nextInt10"] = pq_oracle.java_random(12457, 10, 8).tolist()
    out["java_random_2022_nextInt3"] = pq_oracle.java_random(2022, 3, 8).tolist()
    out["shuffles_PEPTLDEK"] = pq_oracle.shuffles("PEPTLDEK", 5)
    out["shuffles_ELVLSQLGCK"] = pq_oracle.shuffles("ELVLSQLGCK", 5)
    toff, tres, want = synth.planted_targets(prot, np.random.default_rng(7))
    out["novelty_seed7"] = pq_oracle.find_in_proteome(toff, tres, prot[0], prot[1]).tolist()
    assert out["novelty_seed7"] == want.tolist()
    json.dump(out, open(os.path.join(OUT, "oracle_vectors.json"), "w"))
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    print(len(mod_masses()), "modification masses")
    print(len(unimod_ptms()), "unimod specificities")
    v = oracle_vectors()
    print("oracle vectors:", len(v["method1"]["score"]), "hyperscore PSMs,", len(v["method2"]["score"]), "MVH PSMs")
