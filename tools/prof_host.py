import sys,time
sys.path.insert(0,'/root/repo')
from pepquery_b200 import Search,_abi,synth
prot=synth.proteins(20000); tg=synth.targets(2000); sp=synth.spectra(200000,tg,prot)
g=Search(_abi.default_params())
g.load_spectra(sp); g.set_targets(*tg); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
print("=== second pass", file=sys.stderr)
t=time.time(); g.step2(); g.step3(); g.step4(); g.rank(); x=g.psms(); print("step wall", time.time()-t, file=sys.stderr)
