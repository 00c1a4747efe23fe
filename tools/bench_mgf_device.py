"""Throughput of the device MGF parser on the GPU box: pq_parse_mgf_device (sizes call = upload + parse, text in pinned memory) and
pq_load_mgf (parse + load, no targets), beside the host-threaded pq_parse_mgf on the same text.  Prints one JSON line."""
import os, sys, time, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pepquery_b200 import Search, _abi, synth, host

n = int(os.environ.get("N_SPECTRA", 300000))
prot = synth.proteins(2000); tgt = synth.targets(500)
sp = synth.spectra(n, tgt, prot)
keep = []
def pinned(nbytes):
    t = torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True); keep.append(t); return t.numpy()
text = synth.mgf_text(sp, out=pinned)
gb = text.nbytes / 1e9
g = Search(_abi.default_params(n_random=100))
L = g.L
ns, npk = C.c_int64(0), C.c_uint64(0)
p = lambda a: a.ctypes.data_as(C.c_void_p)
times = []
for it in range(4):
    other = text if it % 2 == 0 else text[:text.nbytes - 1]           # defeat the "same text" cache between repetitions
    t0 = time.perf_counter()
    rc = L.pq_parse_mgf_device(g.h, p(other), C.c_uint64(other.nbytes), C.byref(ns), C.byref(npk), None, None, None, None, None)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
    assert rc == 0, (rc, L.pq_last_error(g.h))
t_dev = min(times[1:])
out = dict(prec_mz=np.zeros(ns.value), charge=np.zeros(ns.value, dtype=np.int32), peak_off=np.zeros(ns.value + 1, dtype=np.uint64), mz=np.zeros(npk.value), inten=np.zeros(npk.value))
rc = L.pq_parse_mgf_device(g.h, p(text[:text.nbytes - 1]), C.c_uint64(text.nbytes - 1), C.byref(ns), C.byref(npk), p(out["prec_mz"]), p(out["charge"]), p(out["peak_off"]), p(out["mz"]), p(out["inten"]))
same = bool(np.array_equal(out["mz"], sp["mz"]) and np.array_equal(out["inten"], sp["inten"]) and np.array_equal(out["prec_mz"], sp["prec_mz"]))
tl = []
for it in range(3):
    t0 = time.perf_counter(); g.load_mgf(text); torch.cuda.synchronize(); tl.append(time.perf_counter() - t0)
t_load = min(tl[1:])
th = []
for it in range(2):
    t0 = time.perf_counter(); h = host.parse_mgf(bytes(text)); th.append(time.perf_counter() - t0)
print(json.dumps({"text_GB": round(gb, 3), "spectra": int(ns.value), "peaks": int(npk.value),
                  "device_parse_GBps": round(gb / t_dev, 2), "device_parse_ms": round(1e3 * t_dev, 2),
                  "load_mgf_GBps": round(gb / t_load, 2), "load_mgf_ms": round(1e3 * t_load, 2),
                  "host_parse_GBps": round(gb / min(th), 2), "host_threads": int(os.environ.get("PQ_MGF_THREADS", "0")) or "default",
                  "arrays_equal_to_source": same}))
