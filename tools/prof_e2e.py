import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch, bench, argparse
from pepquery_b200 import Search, _abi
a = argparse.Namespace(spectra=1000000, peptides=10000, proteins=20000, shuffles=1000, method=1)
keep=[]
def pinned(n):
    t = torch.empty(int(n), dtype=torch.float64, pin_memory=True); keep.append(t); return t.numpy()
prot, tgt, sp = bench.make_inputs(a, 0, 1, pinned)
p = _abi.default_params(n_random=1000)
g = Search(p)
for it in range(3):
    t=[time.perf_counter()]
    g.set_targets(*tgt); g.stage_reference(*prot); t.append(time.perf_counter())
    g.load_spectra(sp); t.append(time.perf_counter())
    g.step2(); t.append(time.perf_counter())
    g.build_reference(); t.append(time.perf_counter())
    g.step3(); t.append(time.perf_counter())
    g.step4(); t.append(time.perf_counter())
    g.rank(); ps=g.psms(); t.append(time.perf_counter())
    print("set_targets+stage %.1f load %.1f step2 %.1f build_ref %.1f step3 %.1f step4 %.1f rank+psms %.1f total %.1f ms"%tuple([1e3*(t[i+1]-t[i]) for i in range(7)]+[1e3*(t[-1]-t[0])]))
