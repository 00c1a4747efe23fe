import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench, argparse
from pepquery_b200 import Search, _abi
a = argparse.Namespace(spectra=1000000, peptides=10000, proteins=20000, shuffles=1000, method=1)
prot, tgt, sp = bench.make_inputs(a, 0, 1, None)
p = _abi.default_params(n_random=1000)
g = Search(p)
g.load_spectra(sp); g.set_targets(*tgt); g.step2(); g.build_reference(*prot)
for it in range(4):
    t=[time.perf_counter()]
    g.step2(); t.append(time.perf_counter())
    g.step3(); t.append(time.perf_counter())
    g.step4(); t.append(time.perf_counter())
    g.rank(); t.append(time.perf_counter())
    ps=g.psms(); t.append(time.perf_counter())
    print("step2 %.1f step3 %.1f step4 %.1f rank %.1f psms %.1f total %.1f ms"%tuple([1e3*(t[i+1]-t[i]) for i in range(5)]+[1e3*(t[-1]-t[0])]))
