"""PTM specificities for Step 5, in the shape the C ABI takes (pq_ptm[]).

In PepQuery the Java host imports Unimod itself (ModificationDB.importPTMsFromUnimod,
OpenModificationSearch/ModificationDB.java:334-533) and would pass the resolved table through the shim.  For tests and
the bench the same table is derived from the reference's resources/unimod.xml by tools/make_golden.py and committed as
pepquery_b200/data/unimod_ptms.tsv (record_id, title, site, position, classification, mono_mass)."""
import os

from . import _abi

_DEFAULT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "unimod_ptms.tsv")

_POS = {
    ("Anywhere", True): _abi.PQ_MOD_ANYWHERE,
    ("Any N-term", False): _abi.PQ_MOD_PEP_NTERM, ("Any N-term", True): _abi.PQ_MOD_PEP_NTERM_AA,
    ("Any C-term", False): _abi.PQ_MOD_PEP_CTERM, ("Any C-term", True): _abi.PQ_MOD_PEP_CTERM_AA,
    ("Protein N-term", False): _abi.PQ_MOD_PROT_NTERM, ("Protein N-term", True): _abi.PQ_MOD_PROT_NTERM_AA,
    ("Protein C-term", False): _abi.PQ_MOD_PROT_CTERM, ("Protein C-term", True): _abi.PQ_MOD_PROT_CTERM_AA,
}


def load_ptms(path=_DEFAULT, left=-250.0, right=250.0, skip_classes=("AA substitution",)):
    """-> list of pq_ptm.  Filters of importPTMsFromUnimod: 'AA substitution' specificities are skipped (:390-392);
    the -250..250 mass window (:61-62) and the protein-terminal exclusion (:1037-1040) are also applied by the kernel."""
    out = []
    with open(path) as f:
        head = f.readline().rstrip("\n").split("\t")
        for line in f:
            d = dict(zip(head, line.rstrip("\n").split("\t")))
            if d["classification"] in skip_classes:
                continue
            mass = float(d["mono_mass"])
            if not (left <= mass <= right):
                continue
            site = d["site"]
            is_res = len(site) == 1 and site.isalpha()
            if not is_res and site not in ("N-term", "C-term"):
                continue
            pos = _POS.get((d["position"], is_res))
            if pos is None:
                # "Anywhere" on a terminus token: treat as the peptide terminus
                pos = _abi.PQ_MOD_PEP_NTERM if site == "N-term" else _abi.PQ_MOD_PEP_CTERM
            p = _abi.pq_ptm()
            p.id = int(d["record_id"]); p.position = pos; p.residue = site.encode() if is_res else b"\0"; p.delta = mass
            out.append(p)
    return out
