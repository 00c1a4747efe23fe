"""Times pq_load_spectra alone under a few settings (GPU box): chunk size, with / without prepare_at_load."""
import os, sys, time, argparse
sys.path.insert(0, '/root/repo')
import torch, bench
from pepquery_b200 import Search, _abi
a = argparse.Namespace(spectra=1000000, peptides=10000, proteins=20000, shuffles=1000, method=1)
keep = []
def pinned(n):
    t = torch.empty(int(n), dtype=torch.float64, pin_memory=True); keep.append(t); return t.numpy()
prot, tgt, sp = bench.make_inputs(a, 0, 1, pinned)
for at_load, chunk in [(1, None)] if os.environ.get('PQ_LOAD_DEBUG') else [(0, None), (1, None), (1, 16 << 20), (1, 1 << 20), (1, 1 << 30), (0, 1 << 30)]:
    if chunk: os.environ["PQ_LOAD_CHUNK_PEAKS"] = str(chunk)
    else: os.environ.pop("PQ_LOAD_CHUNK_PEAKS", None)
    g = Search(_abi.default_params(n_random=1000, prepare_at_load=at_load))
    ts = []
    for it in range(4):
        t0 = time.perf_counter(); g.load_spectra(sp); ts.append(1e3 * (time.perf_counter() - t0))
    print("prepare_at_load=%d chunk_peaks=%s load ms: %s" % (at_load, chunk, " ".join("%.1f" % t for t in ts)), flush=True)
    g.close()
