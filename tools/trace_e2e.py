"""GPU timeline of one end-to-end step (torch.profiler = CUPTI activity records of every kernel / copy in the process, whatever
library launched it): per-stream busy time, union busy time, idle gaps, and a coarse time-ordered listing.  Diagnostic only."""
import os, sys, json, argparse, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, bench
from torch.profiler import profile, ProfilerActivity
from pepquery_b200 import Search, _abi
a = argparse.Namespace(spectra=int(os.environ.get("N_SPECTRA", 1000000)), peptides=10000, proteins=20000, shuffles=1000, method=1, scaling="weak")
keep = []
def pinned(n):
    t = torch.empty(int(n), dtype=torch.float64, pin_memory=True); keep.append(t); return t.numpy()
prot, tgt, sp = bench.make_inputs(a, 0, 1, pinned)
out = torch.empty(200 << 20, dtype=torch.uint8, pin_memory=True).numpy()
over = {"search_at_load": int(os.environ.get("SAL", "1"))}
g = Search(_abi.default_params(n_random=1000, **over))
if os.environ.get("RESIDENT"):      # the resident step of bench.py: Steps 2 (incl. K0) + 3 + 4 + results, inputs and reference table in HBM
    g.close(); g = Search(_abi.default_params(n_random=1000, prepare_at_load=0))
    g.load_spectra(sp); g.set_targets(*tgt); g.step2(); g.build_reference(*prot)
    def step():
        g.step2(); g.step3(); g.step4(); g.rank(); g.psms(out); g.competitors(3, out[100 << 20:])
else:
    def step():
        g.set_targets(*tgt); g.stage_reference(*prot); g.load_spectra(sp); g.step2(); g.build_reference(); g.step3(); g.step4(); g.rank()
        g.psms(out); g.competitors(3, out[100 << 20:])
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
path = os.environ.get("TRACE_OUT", "gpurun_out/trace_e2e.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
print("GPU activity span %.2f ms, %d records" % ((t1 - t0) / 1e3, len(ev)))
streams = collections.defaultdict(float)
for e in ev: streams[e["args"].get("stream")] += e["dur"]
print("busy per stream (ms):", {k: round(v / 1e3, 2) for k, v in streams.items()})
# union busy + gaps
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
busy = 0.0; cur_s, cur_e = iv[0]; gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        if s - cur_e > 100: gaps.append((cur_e - t0, s - cur_e))
        cur_s, cur_e = s, e
    else: cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("union busy %.2f ms, idle %.2f ms; gaps > 0.1 ms: %s" % (busy / 1e3, (t1 - t0 - busy) / 1e3, [(round(a / 1e3, 2), round(b / 1e3, 2)) for a, b in gaps]))
# kernels-only union (copies excluded): how long an SM-side kernel was running
kv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev if e["cat"] == "kernel")
kb = 0.0; cs, ce = kv[0]
for s, e in kv[1:]:
    if s > ce: kb += ce - cs; cs, ce = s, e
    else: ce = max(ce, e)
kb += ce - cs
print("kernel union busy %.2f ms" % (kb / 1e3))
# coarse listing: events >= 0.15 ms
for e in ev:
    if e["dur"] >= 150:
        print("%8.2f ms  +%7.2f ms  s%-3s %s" % ((e["ts"] - t0) / 1e3, e["dur"] / 1e3, e["args"].get("stream"), e["name"][:100]))
