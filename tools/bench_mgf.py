"""Host-side throughput of pq_parse_mgf (no GPU needed): synthetic MGF text of ~N MB, count + fill, per thread count."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pepquery_b200 import _abi

def make_text(n_spectra, seed=1):
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(n_spectra):
        k = int(rng.integers(60, 160))
        mz = np.sort(rng.uniform(100, 2000, k)); it = rng.lognormal(7, 1, k)
        body = "\n".join("%.4f %.2f" % (a, b) for a, b in zip(mz, it))
        parts.append("BEGIN IONS\nTITLE=%d\nPEPMASS=%.5f\nCHARGE=%d+\n%s\nEND IONS\n" % (i + 1, rng.uniform(400, 1200), int(rng.integers(2, 4)), body))
    return "".join(parts).encode()

if __name__ == "__main__":
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    one = make_text(2000)
    text = one * max(1, mb * 1000000 // len(one))
    L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pepquery_b200", "libpepquery_b200.so"))
    for threads in (1, 4, 8, 16):
        os.environ["PQ_MGF_THREADS"] = str(threads)
        ns, npk = C.c_int64(0), C.c_uint64(0)
        t0 = time.perf_counter()
        L.pq_parse_mgf(text, C.c_uint64(len(text)), C.byref(ns), C.byref(npk), None, None, None, None, None)
        t1 = time.perf_counter()
        prec = np.zeros(ns.value); chg = np.zeros(ns.value, dtype=np.int32); off = np.zeros(ns.value + 1, dtype=np.uint64)
        mz = np.empty(npk.value); it = np.empty(npk.value)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        t2 = time.perf_counter()
        rc = L.pq_parse_mgf(text, C.c_uint64(len(text)), C.byref(ns), C.byref(npk), p(prec), p(chg), p(off), p(mz), p(it))
        t3 = time.perf_counter()
        print("threads %2d: %d MB, %d spectra, %d peaks: count %.3f s, count+fill %.3f s -> %.2f GB/s (rc %d)" %
              (threads, len(text) // 1000000, ns.value, npk.value, t1 - t0, t3 - t2, len(text) / 1e9 / (t3 - t2), rc))
