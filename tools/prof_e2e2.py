"""End-to-end phase times (host wall clock per call) with and without search_at_load; PQ_DEBUG=1 adds the library's own marks."""
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
out = torch.empty(200 << 20, dtype=torch.uint8, pin_memory=True).numpy()
for name, over in (("search_at_load", {}), ("step by step", {"search_at_load": 0})):
    g = Search(_abi.default_params(n_random=1000, **over))
    rows = []
    for it in range(5):
        if it == 4: os.environ["PQ_DEBUG"] = "1"
        t = [time.perf_counter()]
        g.set_targets(*tgt); g.stage_reference(*prot); t.append(time.perf_counter())
        g.load_spectra(sp); t.append(time.perf_counter())
        g.step2(); t.append(time.perf_counter())
        g.build_reference(); t.append(time.perf_counter())
        g.step3(); t.append(time.perf_counter())
        g.step4(); t.append(time.perf_counter())
        g.rank(); t.append(time.perf_counter())
        ps = g.psms(out); t.append(time.perf_counter())
        g.competitors(3, out[100 << 20:]); t.append(time.perf_counter())
        os.environ.pop("PQ_DEBUG", None)
        rows.append([1e3 * (t[i + 1] - t[i]) for i in range(9)] + [1e3 * (t[-1] - t[0])])
    r = np.array(rows[1:4]).mean(axis=0)
    print("%-16s targets+stage %.1f load %.1f step2 %.1f build_ref %.1f step3 %.1f step4 %.1f rank %.1f psms %.1f competitors %.1f | total %.1f ms" % ((name,) + tuple(r)), flush=True)
    g.close()
