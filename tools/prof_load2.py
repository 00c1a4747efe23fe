"""Times the end-to-end phases on the GPU box under the loader's modes: candidate-first gather (with / without K0 under it, chunk
sizes), full chunked upload from pinned and from pageable buffers.  PQ_DEBUG=1 prints the phase marks of every call."""
import os, sys, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, bench
from pepquery_b200 import Search, _abi
a = argparse.Namespace(spectra=int(os.environ.get("N_SPECTRA", 1000000)), peptides=10000, proteins=20000, shuffles=1000, method=1, scaling="weak")
keep = []
def pinned(n):
    t = torch.empty(int(n), dtype=torch.float64, pin_memory=True); keep.append(t); return t.numpy()
prot, tgt, sp = bench.make_inputs(a, 0, 1, pinned)
spg = {k: np.array(v) for k, v in sp.items()}
modes = [("cf", {}, 1, sp), ("g8", {"PQ_GATHER_CTAS": "8"}, 1, sp), ("g16", {"PQ_GATHER_CTAS": "16"}, 1, sp), ("g64", {"PQ_GATHER_CTAS": "64"}, 1, sp), ("g148", {"PQ_GATHER_CTAS": "148"}, 1, sp),
         ("g16 noK0", {"PQ_GATHER_CTAS": "16"}, 0, sp), ("g64 c131072", {"PQ_GATHER_CTAS": "64", "PQ_LOAD_CHUNK_SPECTRA": "131072"}, 1, sp), ("cf no K0 in load", {}, 0, sp), ("cf chunk 8192", {"PQ_LOAD_CHUNK_SPECTRA": "8192"}, 1, sp), ("cf chunk 131072", {"PQ_LOAD_CHUNK_SPECTRA": "131072"}, 1, sp),
         ("cf one chunk", {"PQ_LOAD_CHUNK_SPECTRA": "100000000"}, 1, sp), ("full pinned", {"PQ_LOAD_FULL": "1"}, 1, sp), ("full pageable", {}, 1, spg)]
only = os.environ.get("PROF_MODES")
for name, env, at_load, spx in modes:
    if only and name not in only.split(","):
        continue
    for k in ("PQ_LOAD_CHUNK_SPECTRA", "PQ_LOAD_FULL", "PQ_GATHER_CTAS"): os.environ.pop(k, None)
    os.environ.update(env)
    g = Search(_abi.default_params(n_random=1000, prepare_at_load=at_load))
    rows = []
    for it in range(4):
        t = [time.perf_counter()]
        g.set_targets(*tgt); g.stage_reference(*prot); t.append(time.perf_counter())
        g.load_spectra(spx); t.append(time.perf_counter())
        g.step2(); t.append(time.perf_counter())
        g.build_reference(); t.append(time.perf_counter())
        g.step3(); t.append(time.perf_counter())
        g.step4(); t.append(time.perf_counter())
        g.rank(); ps = g.psms(); g.competitors(3); t.append(time.perf_counter())
        rows.append([1e3 * (t[i + 1] - t[i]) for i in range(7)] + [1e3 * (t[-1] - t[0])])
    r = np.array(rows[1:]).mean(axis=0)
    c = g.counters()
    print("%-18s targets+stage %.1f load %.1f step2 %.1f build_ref %.1f step3 %.1f step4 %.1f results %.1f | total %.1f ms  (h2d %.2f GB/step, prepared %d)" %
          ((name,) + tuple(r) + (c.h2d_bytes / 4 / 1e9, c.spectra_prepared)), flush=True)
    g.close()
