"""Multi-GPU plumbing of the path: spectra shard by precursor-mass range, nothing else is exchanged.

Every (peptide, spectrum) pair involves exactly one spectrum and all candidates of a spectrum (targets, reference forms,
shuffles, PTM forms) are decided by its precursor mass, so the units are independent: partition the spectra into
`world` contiguous mass ranges of equal estimated work, replicate the (small) peptide tables on every rank, run the
unchanged single-GPU pipeline per rank, and gather the fixed-size PSM records to rank 0.  That gather is the only
collective (NCCL on GPUs, gloo in the CPU tests).  The reference has no counterpart: it is one JVM with thread pools
(SURVEY.md §2 rows 19-20).
"""
import numpy as np

PROTON = 1.007276466812


def precursor_mass(prec_mz, charge):
    return prec_mz * charge - charge * PROTON


def shard_bounds(masses, world, weights=None):
    """world+1 mass boundaries; shard r takes spectra with bounds[r] <= mass < bounds[r+1].
    Equal estimated work, not equal mass width: weights (e.g. candidate counts) are prefix-summed over the mass order."""
    m = np.asarray(masses, dtype=np.float64)
    order = np.argsort(m, kind="stable")
    w = np.ones(len(m)) if weights is None else np.asarray(weights, dtype=np.float64)[order]
    cum = np.cumsum(w)
    total = cum[-1] if len(cum) else 0.0
    bounds = [-np.inf]
    for r in range(1, world):
        k = int(np.searchsorted(cum, total * r / world, side="left"))
        k = min(max(k, 0), len(m) - 1)
        bounds.append(float(m[order[k]]))
    bounds.append(np.inf)
    for r in range(1, world + 1):          # strictly usable as half-open intervals
        bounds[r] = max(bounds[r], bounds[r - 1])
    return np.array(bounds)


def take_shard(sp, bounds, rank):
    """The sub-CSR of shard `rank` plus the map local index -> global index."""
    mass = precursor_mass(sp["prec_mz"], sp["charge"])
    sel = np.nonzero((mass >= bounds[rank]) & (mass < bounds[rank + 1]))[0]
    off = sp["peak_off"]
    counts = (off[sel + 1] - off[sel]).astype(np.uint64)
    new_off = np.zeros(len(sel) + 1, dtype=np.uint64)
    np.cumsum(counts, out=new_off[1:])
    # peak indices of the selected spectra: for every output peak, its spectrum's first input peak + its position inside it
    total = int(new_off[-1])
    idx = (np.repeat(off[sel].astype(np.int64) - new_off[:-1].astype(np.int64), counts.astype(np.int64)) + np.arange(total, dtype=np.int64)) if total else np.zeros(0, dtype=np.int64)
    sub = dict(prec_mz=np.ascontiguousarray(sp["prec_mz"][sel]), charge=np.ascontiguousarray(sp["charge"][sel]), peak_off=new_off,
               mz=np.ascontiguousarray(sp["mz"][idx]), inten=np.ascontiguousarray(sp["inten"][idx]))
    return sub, sel.astype(np.int64)


def gather_records(records, dist, rank, world, device="cpu"):
    """Gather a structured numpy array of fixed-size records to rank 0 (all_gather of the byte counts, then one padded
    all_gather of the bytes: gloo has no variable-size gather).  Returns the concatenation on rank 0, None elsewhere."""
    import torch
    if world == 1:
        return records
    raw = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).reshape(-1).copy()).to(device)
    n = torch.tensor([raw.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    buf = torch.zeros(max(max(sizes), 1), dtype=torch.uint8, device=device)
    buf[:raw.numel()] = raw
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    if rank != 0:
        return None
    parts = [out[r][:sizes[r]].cpu().numpy().view(records.dtype) for r in range(world)]
    return np.concatenate(parts) if parts else records[:0]
