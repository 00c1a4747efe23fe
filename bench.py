#!/usr/bin/env python
"""bench.py — peptide-spectrum pairs scored per second (incl. shuffles) on the C2-shaped workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong] [--spectra S] [--peptides P]

One "step" = one pass of the hot path (Step 2 target matching + Step 3 competitor search + Step 4 shuffles,
SURVEY.md §3.2-3.4) over one resident batch of spectra.  N > 1: launched by torchrun, one rank per GPU, spectra sharded
by precursor-mass range (targets and the protein database are replicated); the per-rank PSM records are gathered to rank 0
over NCCL inside the timed region.
  --scaling weak   (default) each rank holds --spectra spectra of its own interleaved mass strips;
  --scaling strong ONE dataset of --spectra spectra, cut into N contiguous mass ranges of equal estimated pair count
                   (pepquery_b200.shard.shard_bounds / take_shard — the partition a deployment would use).

Keys of the JSON line (rank 0): see the task contract.
  value             pairs/s with inputs resident in HBM (Step 4 decides shuffles by their score bound where it can)
  value_all_scored  the same with every shuffle scored in full (pq_params.score_all_shuffles)
  e2e               the same metric through the C ABI from HOST (pinned) buffers, H2D/D2H and the reference-table build inside the
                    timed region; e2e.pageable: the same from pageable buffers (needed spectra packed into a page-locked ring by host threads)
  roofline          the Step-4 K1 launch: physical DRAM traffic against the measured HBM peak, and warp-instruction issue rate
                    against the issue peak (the resource that binds); the algorithmic-bytes figure of SURVEY.md §8d beside them
  configs           driver-visible numbers of the other BASELINE.json configs (C1, C3, C4, one mass shard of C5)
  cpu_baseline      the CPU oracle ("port") on a bounded sample of the same workload: as written (pre-processing per pair) and hoisted
"""
import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STRIPS_PER_GPU = 16
METRIC = "peptide_spectrum_pairs_scored_per_sec"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--spectra", type=int, default=1000000, help="spectra per GPU (weak) / in total (strong)")
    ap.add_argument("--peptides", type=int, default=10000)
    ap.add_argument("--proteins", type=int, default=20000)
    ap.add_argument("--shuffles", type=int, default=1000)
    ap.add_argument("--method", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-warmup", type=int, default=2)
    ap.add_argument("--e2e-two-in-flight", type=int, default=0, help="also time the end-to-end step with two searches in flight (informational)")
    ap.add_argument("--cpu-sample", type=int, default=12000, help="spectra in the CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 / C5 lines")
    ap.add_argument("--score-all-shuffles", action="store_true", help="compute every shuffle score in full (no upper-bound decision in Step 4)")
    ap.add_argument("--step5", action="store_true", help="config C4 only: time the unrestricted-modification search (Step 5) after Steps 2-4")
    return ap.parse_args()


def workload_name(a, world=1):
    return ("C2: %d peptides x %d spectra%s x %d-protein DB, %s, %d shuffles, fixMod C+57 varMod M+16" %
            (a.peptides, a.spectra, " in total" if a.scaling == "strong" and world > 1 else "/GPU", a.proteins,
             "hyperscore" if a.method == 1 else "MVH", a.shuffles))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe: start before, kill after).
    The process is started ahead of the timed region (its start-up — driver queries of several 10 ms — was seen to cost the first
    timed steps ~1 ms each when it began inside); only the samples whose own time stamps fall inside the region are used."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def region_begin(self):
        self.t0 = time.time()

    def region_end(self):
        self.t1 = time.time()

    @staticmethod
    def _stamp(txt):
        import datetime
        try:
            return datetime.datetime.strptime(txt.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                rows.append((self._stamp(f[0]), float(f[2]), float(f[3]), [nm for k, nm in enumerate(names) if f[6 + k].lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if r[0] is not None and self.t0 is not None and self.t0 - 0.03 <= r[0] <= (self.t1 or 1e18) + 0.03]
        use = inside if inside else rows[-3:]          # a region shorter than the sampling period: the samples next to it
        sm = [r[1] for r in use]; mx = [r[2] for r in use]; reasons = set(x for r in use for x in r[3])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_region": len(inside)}


def java_probe():
    """Is the reference itself runnable on this box?  (the intended CPU baseline is PepQuery's own Java path.)"""
    out = {"java": None, "pepquery_jar": None}
    try:
        r = subprocess.run(["java", "-version"], capture_output=True, text=True, timeout=20)
        txt = (r.stderr or r.stdout or "").strip().splitlines()
        out["java"] = txt[0] if txt else "java present, no version text"
    except FileNotFoundError:
        out["java"] = "not found (java: command not found)"
    except Exception as e:                                                    # noqa: BLE001
        out["java"] = "probe failed: %r" % (e,)
    skip = {".venv", "site-packages", "node_modules", ".git", "proc", "sys", "__pycache__"}
    for top in ("/opt", "/usr/local", "/usr/share", "/root", "/home", ROOT):      # a bounded look for a PepQuery jar (depth <= 3)
        if out["pepquery_jar"] or not os.path.isdir(top):
            continue
        base = top.rstrip("/").count("/")
        for d, dirs, files in os.walk(top):
            dirs[:] = [x for x in dirs if x not in skip]
            if d.count("/") - base >= 3:
                dirs[:] = []
            hit = [f for f in files if f.lower().startswith("pepquery") and f.endswith(".jar")]
            if hit:
                out["pepquery_jar"] = os.path.join(d, hit[0])
                break
    out["reference_runnable"] = bool(out["pepquery_jar"]) and out["java"] is not None and not str(out["java"]).startswith(("not found", "probe failed"))
    return out


def pin_copy(arr, pinned):
    buf = pinned((arr.nbytes + 7) // 8)                                       # float64 words
    out = buf.view(np.uint8)[:arr.nbytes].view(arr.dtype)
    out[:] = arr
    return out


def make_inputs(a, rank, world, pinned, peptides=None, spectra=None, strips=True):
    """Synthetic proteins / targets / spectra of one rank.  weak: the rank's own interleaved mass strips;
    strong: the full dataset, cut by shard_bounds / take_shard (see shard_inputs)."""
    from pepquery_b200 import synth
    prot = synth.proteins(a.proteins)
    tgt = synth.targets(peptides or a.peptides)
    n = spectra or a.spectra
    if world > 1 and strips:
        # mass-range sharding with equal expected work: the precursor-mass axis is cut into STRIPS_PER_GPU * world quantile
        # strips and rank r owns strips r, r + world, ... (contiguous mass ranges, interleaved so that every rank sees the
        # same mix of peptide lengths); each rank holds n spectra of its own strips (weak scaling)
        b = synth.mass_bounds(STRIPS_PER_GPU * world, tgt, prot)
        synth.set_strips(b[:-1], world, rank)
    sp = synth.spectra(n, tgt, prot, first_index=rank * n if strips else 0, pinned=pinned)
    synth.set_strips(None, 1, 0)
    if pinned is not None:
        for k in ("prec_mz", "charge", "peak_off"):
            sp[k] = pin_copy(sp[k], pinned)
        prot = (pin_copy(prot[0], pinned), pin_copy(prot[1], pinned))
    return prot, tgt, sp


def shard_inputs(sp, tgt_masses, tol_ppm, rank, world, pinned):
    """Strong scaling: contiguous mass ranges of equal estimated pair count.  Weight of a spectrum = target forms in its
    precursor window (each becomes up to 1000 shuffle pairs) x its precursor mass (peptide length ~ ladder length)."""
    from pepquery_b200 import shard
    mass = shard.precursor_mass(sp["prec_mz"], sp["charge"])
    w = tol_ppm * 1e-6 * mass
    cand = np.searchsorted(tgt_masses, mass + w) - np.searchsorted(tgt_masses, mass - w)
    bounds = shard.shard_bounds(mass, world, weights=cand * mass + 1e-3)
    sub, sel = shard.take_shard(sp, bounds, rank)
    if pinned is not None:
        sub = {k: pin_copy(v, pinned) for k, v in sub.items()}
    return sub, bounds, len(sel)


def oracle_sample_rate(params, sp, tgt, prot, ns, threads, hoisted, timed_reference=False):
    """Steps 2-4 of the CPU oracle on the first ns spectra; -> (pairs/s, pairs, seconds)."""
    from oracle import pq_oracle
    sub = {"prec_mz": sp["prec_mz"][:ns].copy(), "charge": sp["charge"][:ns].copy(), "peak_off": sp["peak_off"][:ns + 1].copy(),
           "mz": np.array(sp["mz"][:int(sp["peak_off"][ns])]), "inten": np.array(sp["inten"][:int(sp["peak_off"][ns])])}
    o = pq_oracle.Oracle(params, threads=threads)
    o.set_cost_model(hoisted)
    o.load_spectra(sub); o.set_targets(*tgt)
    t0 = time.perf_counter(); o.step2(); t1 = time.perf_counter()
    o.build_reference(*prot)                  # not timed: fixed cost, independent of the sample size
    t2 = time.perf_counter(); o.step3(); o.step4(); o.rank(); t3 = time.perf_counter()
    pairs = int(o.pairs()[:3].sum())
    o.close()
    sec = (t1 - t0) + (t3 - t2)
    return pairs / sec, pairs, sec


def run_reference(a, rank, world):
    """The reference arm: the CPU oracle ("port"; the Java reference cannot run here — no JVM, un-vendored jars — see
    cpu_baseline.java), all host threads, each step a bounded sample of the workload.  The as-written cost model (spectrum
    pre-processing redone per pair, like PeptideSearchMT.scorePeptide2Spectrum) is the line's value; the hoisted model is beside it."""
    if rank != 0:
        return
    from pepquery_b200 import _abi
    threads = os.cpu_count() or 1
    prot, tgt, sp = make_inputs(argparse.Namespace(**{**vars(a), "spectra": a.cpu_sample}), 0, 1, None)
    p = _abi.default_params(n_random=a.shuffles, score_method=a.method)
    times, pairs = [], 0
    warm = a.warmup
    for it in range(a.warmup + a.steps):
        rate, pairs, sec = oracle_sample_rate(p, sp, tgt, prot, a.cpu_sample, threads, hoisted=False)
        if it >= warm:
            times.append(sec)
        if it == 0 and sec > 60 and warm > 1:
            warm = 1
        if len(times) >= a.steps:
            break
    tot = sum(times)
    val = pairs * len(times) / tot
    h_rate, h_pairs, h_sec = oracle_sample_rate(p, sp, tgt, prot, a.cpu_sample, threads, hoisted=True)
    sample = "%d spectra of the workload x all %d peptides x full DB, Steps 2-4, %d pairs per step" % (a.cpu_sample, a.peptides, pairs)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": len(times), "warmup": warm,
            "ms_per_step": 1000.0 * tot / len(times), "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "cost_model": "as written (pre-processing redone per pair)",
                             "hoisted": {"value": h_rate, "unit": UNIT, "cores": threads, "seconds": h_sec,
                                         "cost_model": "spectrum pre-processing once per spectrum (the best a CPU port can do)"},
                             "java": java_probe()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_near_gpu(index):
    """Run this rank on the CPUs NVML reports as local to its GPU, so the pinned host buffers it allocates (first touch) sit on
    the NUMA node the GPU's PCIe root hangs off.  Best effort: silently keeps the inherited affinity when NVML says nothing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def step4_profile(a):
    """Counters of the Step-4 K1 launch from the committed ncu capture of this workload (profiles/k1_step4_profile.json): DRAM bytes
    and warp instructions per launch are properties of the kernel + workload, the duration they are divided by is measured live."""
    path = os.path.join(ROOT, "profiles", "k1_step4_profile.json")
    if not os.path.exists(path):
        return None
    tj = json.load(open(path)); w = tj.get("workload", {})
    if (w.get("spectra"), w.get("peptides"), w.get("proteins"), w.get("shuffles"), w.get("method")) != (a.spectra, a.peptides, a.proteins, a.shuffles, a.method):
        return None
    return tj


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    import torch
    import torch.distributed as dist
    from pepquery_b200 import Search, _abi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    full_affinity = os.sched_getaffinity(0)
    bind_near_gpu(local)                       # pinned buffers and the calling thread on the GPU's NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    keep = []

    def pinned(n):
        t = torch.empty(int(n), dtype=torch.float64, pin_memory=True)
        keep.append(t)
        return t.numpy()

    def mk_params(at_load, method=None, score_all=None, **over):
        return _abi.default_params(n_random=a.shuffles, score_method=method or a.method, device=local, prepare_at_load=at_load,
                                   score_all_shuffles=(1 if a.score_all_shuffles else 0) if score_all is None else score_all, **over)

    t_gen = time.time()
    strong = a.scaling == "strong" and world > 1
    shard_info = None
    if strong:
        prot, tgt, sp_all = make_inputs(a, 0, 1, None, strips=False)          # the ONE dataset, identical on every rank
        g0 = Search(mk_params(0)); g0.set_targets(*tgt)
        forms, nf = g0.target_forms(); g0.close()
        tmass = np.sort(np.array([forms[i].mass for i in range(nf)]))
        sp, bounds, n_mine = shard_inputs(sp_all, tmass, 10.0, rank, world, pinned)
        prot = (pin_copy(prot[0], pinned), pin_copy(prot[1], pinned))
        shard_info = {"bounds": [float(x) for x in bounds[1:-1]], "spectra_this_rank": int(n_mine)}
        del sp_all
    else:
        prot, tgt, sp = make_inputs(a, rank, world, pinned)
    t_gen = time.time() - t_gen
    n_local = len(sp["prec_mz"])

    class _DevView:
        """CUDA array interface over a device address (the context's export buffer)."""
        def __init__(self, addr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (addr, False), "version": 2}

    REC = 96                                    # sizeof(pq_psm)
    HDR = 64                                    # header of the gather buffer: the rank's record count
    gbuf = {}

    def size_gather_buffers(g):
        """Outside the timed region: one fixed capacity for every rank's record buffer (largest count over the ranks + 25 %), so
        that a step needs no size exchange and no host sync before the gather."""
        n = g.psm_count()
        cap = torch.tensor([n], device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        cap = int(cap.item()); cap = cap + cap // 16 + 1024
        gbuf["cap"] = cap
        gbuf["pin_psm"] = torch.empty(HDR + cap * REC, dtype=torch.uint8, pin_memory=True).numpy()
        gbuf["pin_comp"] = torch.empty(max(1, cap) * 32 + 4096, dtype=torch.uint8, pin_memory=True).numpy()
        if world > 1:
            gbuf["send"] = torch.zeros(HDR + cap * REC, dtype=torch.uint8, device="cuda")
            gbuf["recv"] = [torch.zeros(HDR + cap * REC, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
            gbuf["host"] = torch.empty((HDR + cap * REC) * world, dtype=torch.uint8, pin_memory=True) if rank == 0 else None
            gbuf["hdr"] = torch.zeros(HDR // 8, dtype=torch.int64, device="cuda")

    def gather_psms(g):
        """Results to the host.  One GPU: device -> pinned host.  N GPUs: NCCL gather of fixed-capacity record buffers (count in the
        header) from every rank's device buffer to rank 0 — the only collective of the path —, the detail rows of each rank going to
        its host while the gather is in flight, then device -> pinned host on rank 0."""
        g.rank()
        if world == 1:
            ps = g.psms(gbuf["pin_psm"][HDR:]) if g.psm_count() <= gbuf["cap"] else g.psms()
            g.competitors(3, gbuf["pin_comp"])
            return len(ps)
        addr, n = g.psms_device()
        if n > gbuf["cap"]:
            raise SystemExit("gather buffer too small: size_gather_buffers() must see the final record counts")
        gbuf["hdr"][0] = n
        gbuf["send"][:HDR].copy_(gbuf["hdr"].view(torch.uint8))
        if n:
            gbuf["send"][HDR:HDR + n * REC].copy_(torch.as_tensor(_DevView(addr, n * REC), device="cuda"))
        work = dist.gather(gbuf["send"], gbuf["recv"], dst=0, async_op=True)
        g.competitors(3, gbuf["pin_comp"])      # detail rows stay with the rank that owns the spectra; overlaps the gather
        work.wait()
        if rank == 0:
            stride = HDR + gbuf["cap"] * REC
            for r in range(world):                 # the headers first (one short wait), then exactly the records each rank sent
                gbuf["host"][r * stride:r * stride + HDR].copy_(gbuf["recv"][r][:HDR], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            hv = gbuf["host"].numpy()
            counts = [int(hv[r * stride:r * stride + 8].view(np.int64)[0]) for r in range(world)]
            for r in range(world):
                used = HDR + counts[r] * REC
                gbuf["host"][r * stride + HDR:r * stride + used].copy_(gbuf["recv"][r][HDR:used], non_blocking=True)
            torch.cuda.synchronize()
            return int(sum(counts))
        torch.cuda.synchronize()
        return 0

    def timed(fn, steps, g):
        """steps x fn() between two barriers; -> (seconds of the slowest rank, this rank's seconds)."""
        barrier()
        t0 = time.perf_counter()
        g.event_record(0)                      # CUDA events on the library's own stream
        out = None
        for _ in range(steps):
            out = fn()
        g.event_record(1)
        dev_ms = g.event_elapsed_ms()
        mine = max(time.perf_counter() - t0, dev_ms / 1e3)      # device-event time; wall between syncs is the same interval seen from the host
        barrier()
        tm = torch.tensor([mine], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item()), mine, out

    def total_over_ranks(x):
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def resident_run(params, steps, warm, before_warmup=None):
        """Spectra, targets and the reference table resident; one step = Steps 2 (incl. K0) + 3 + 4 + rank + results to the host."""
        g = Search(params)
        g.load_spectra(sp); g.set_targets(*tgt); g.step2(); g.build_reference(*prot)
        ms_ref = g.counters().ms_host_reference

        def step():
            g.step2(); g.step3(); g.step4()
            return gather_psms(g)
        g.step3(); g.step4(); g.rank()
        g.ms_ladder_once = g.counters().ms_ladder        # fragment-ladder tables of the reference rows: built with the table, not per step
        size_gather_buffers(g)
        if before_warmup:
            before_warmup()                     # the clock sampler starts here: its start-up falls into the warm-up steps
            time.sleep(0.25)
        for _ in range(max(warm, 1)):
            step()
        g.reset_counters()
        return g, step, ms_ref

    if a.step5:
        run_c4(a, mk_params, sp, tgt, prot, rank)
        return

    # ---- resident-input measurement ----------------------------------------------------------------------------
    # spectrum preparation (K0) stays inside the timed Step 2, so prepare_at_load is off here; the end-to-end arm below uses the
    # library defaults (candidate-first upload, K0 chunk by chunk under it)
    warm = max(a.warmup, 3)
    clocks = ClockSampler(local)
    g, step, ms_host_reference = resident_run(mk_params(0), a.steps, warm, before_warmup=clocks.start)
    barrier()
    clocks.region_begin()
    t_max, t_mine, n_psm_total = timed(step, a.steps, g)
    clocks.region_end()
    clk = clocks.stop()
    c = g.counters()
    pairs_step = (c.pairs_step2 + c.pairs_step3 + c.pairs_step4) / a.steps
    launches = (c.score_kernel_launches + c.other_kernel_launches)
    sec_per_step = t_max / a.steps
    tot_pairs = total_over_ranks(pairs_step)
    value = tot_pairs / sec_per_step
    per_rank_ms = None
    if world > 1:
        tl = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(tl, torch.tensor([1e3 * t_mine / a.steps], device="cuda", dtype=torch.float64))
        per_rank_ms = [round(float(x.item()), 3) for x in tl]

    # roofline of the dominant kernel: the Step-4 K1 launch (CUDA events on the library's stream, around the launch)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    k1_ms = c.ms_score_step[2] / a.steps
    k1_bytes = c.bytes_score_step[2] / a.steps
    k1_pairs = c.pairs_step4 / a.steps
    prof = step4_profile(a) if not a.score_all_shuffles else None
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    sm_mhz = clk.get("sm_mhz") or 1965.0
    traffic = prof["dram_bytes_per_launch"] if prof else None
    achieved = (traffic / 1e9) / (k1_ms / 1e3) if (traffic and k1_ms > 0) else None
    issue = None
    if prof and k1_ms > 0:
        ips = prof["warp_inst_per_launch"] / (k1_ms / 1e3)
        ipeak = sm_count * 4 * sm_mhz * 1e6
        issue = {"warp_inst_per_launch": prof["warp_inst_per_launch"], "warp_inst_per_pair": prof["warp_inst_per_launch"] / max(k1_pairs, 1),
                 "inst_per_s": ips, "peak": ipeak, "peak_is": "%d SMs x 4 schedulers x %.0f MHz (one warp instruction per scheduler and cycle)" % (sm_count, sm_mhz),
                 "frac": ips / ipeak, "ncu_issue_active_pct": prof.get("issue_active_pct")}
        if prof.get("lsu_wavefronts_per_launch"):
            # the L1 / shared-memory data pipe: one wavefront per SM and cycle; the count pass is a byte look-up per fragment cell in shared memory
            wps = prof["lsu_wavefronts_per_launch"] / (k1_ms / 1e3); wpeak = sm_count * sm_mhz * 1e6
            issue["lsu_data_pipe"] = {"wavefronts_per_launch": prof["lsu_wavefronts_per_launch"], "shared_memory_share": prof["lsu_wavefronts_shared_per_launch"] / prof["lsu_wavefronts_per_launch"],
                                      "wavefronts_per_s": wps, "peak": wpeak, "peak_is": "%d SMs x %.0f MHz (one L1/shared-memory wavefront per SM and cycle)" % (sm_count, sm_mhz),
                                      "frac": wps / wpeak, "ncu_pct": prof.get("lsu_wavefronts_pct")}
    alg = (k1_bytes / 1e9) / (k1_ms / 1e3) if k1_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": "k1_score_kernel<128,%d> (Step 4, one launch per step)" % a.method, "peak_source": peak_src,
                "ms_per_launch": k1_ms, "pairs_per_launch": k1_pairs,
                "issue": issue, "binding": ("the L1/shared-memory data pipe (issue.lsu_data_pipe.frac) and warp-instruction issue slots (issue.frac), not HBM" if issue and issue.get("lsu_data_pipe")
                                            else "issue slots (warp-instruction issue), not HBM") if issue else None,
                "algorithmic": {"bytes_per_launch": k1_bytes, "GBps": alg, "x_hbm_peak": alg / peak,
                                "note": "SURVEY.md 8d accounting: sum over pairs of 16*P_s+96 bytes, as if every pair re-read its spectrum.  The kernel stages a "
                                        "spectrum once per ~1000 shuffles, so this figure is NOT a roofline fraction (it exceeds 1); achieved/frac above are "
                                        "physical DRAM bytes (ncu dram__bytes_read+write of this launch) over the live launch time"},
                "profile": prof.get("source") if prof else "no committed ncu capture for this workload: traffic / issue not reported"}
    kernel_ms = {"k1_step2": c.ms_score_step[0] / a.steps, "k1_step3": c.ms_score_step[1] / a.steps, "k1_step4": k1_ms,
                 "k0_prepare+pack": c.ms_prepare / a.steps, "k3_window": c.ms_window / a.steps, "k2_shuffle": c.ms_shuffle / a.steps,
                 "k1_ladder_tables_once": g.ms_ladder_once}
    bounded_step = float(c.pairs_step4_bounded) / a.steps
    g.close()

    # ---- the same with every shuffle scored in full ------------------------------------------------------------------
    all_scored = None
    if not a.score_all_shuffles:
        g, step, _ = resident_run(mk_params(0, score_all=1), 3, 2)
        tm, _, _ = timed(step, 3, g)
        ca = g.counters()
        pa = total_over_ranks((ca.pairs_step2 + ca.pairs_step3 + ca.pairs_step4) / 3)
        all_scored = {"value": pa / (tm / 3), "unit": UNIT, "ms_per_step": 1e3 * tm / 3, "steps": 3, "k1_step4_ms": ca.ms_score_step[2] / 3,
                      "pairs_decided_by_score_bound": float(ca.pairs_step4_bounded)}
        g.close()

    # ---- end-to-end from host buffers ---------------------------------------------------------------------------
    e2e = None
    if a.e2e_steps > 0:
        def e2e_arm(spx, steps, warm_steps):
            g = Search(mk_params(1))                   # a fresh context with the library defaults

            def e2e_step():
                g.set_targets(*tgt)                    # the reference's order: targets (Step 1) before the spectra are read
                g.stage_reference(*prot)               # H2D of the proteins (the database is known from the command line)
                g.load_spectra(spx)                    # precursor arrays up, windows enumerated, the peaks of spectra with a target in window gathered; K0 and the digest under it
                g.step2()
                g.build_reference()                    # isoform expansion inside the matched-spectra mass window + mass sort (K6)
                g.step3(); g.step4()
                return gather_psms(g)                  # rank/accept + D2H of the result records
            g.set_targets(*tgt); g.load_spectra(spx); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
            size_gather_buffers(g)
            for _ in range(max(warm_steps, 1)):
                e2e_step()                             # untimed: first-use allocations of the new context
            g.reset_counters()
            tm, _, _ = timed(e2e_step, steps, g)
            ce = g.counters()
            pe = total_over_ranks((ce.pairs_step2 + ce.pairs_step3 + ce.pairs_step4) / steps)
            out = {"value": pe / (tm / steps), "unit": UNIT,
                   "h2d_bytes_per_step": int(ce.h2d_bytes / steps), "d2h_bytes_per_step": int(ce.d2h_bytes / steps),
                   "ms_per_step": 1000.0 * tm / steps, "steps": steps, "spectra_prepared_per_step": int(ce.spectra_prepared),
                   "ms_host_reference_build": ce.ms_host_reference / steps}
            return g, out
        g, e2e = e2e_arm(sp, a.e2e_steps, a.e2e_warmup)
        e2e["buffers"] = "pinned host arrays: candidate-first upload (pq_params.candidate_first)"
        g.close()
        # the same call sequence from pageable buffers (what a JVM hands over unless it registers its off-heap memory): candidate-first through a host-packed page-locked ring
        if world == 1:
            spg = {k: np.array(v) for k, v in sp.items()}
            g, pg = e2e_arm(spg, 2, 1)
            e2e["pageable"] = {k: pg[k] for k in ("value", "unit", "ms_per_step", "steps", "h2d_bytes_per_step")}
            e2e["pageable"]["buffers"] = "pageable host arrays: the needed spectra packed by host threads into a page-locked ring, uploaded chunk by chunk"
            g.close(); del spg
        # Informational (not `value`): the same end-to-end step with TWO searches in flight, each in its own context driven by its
        # own host thread — what a sweep over many spectrum files does: one search uploads while the other computes.
        if world == 1 and a.e2e_two_in_flight:
            try:
                ctxs = [Search(mk_params(1)), Search(mk_params(1))]
                bufs = [None, None]
                upload, compute = threading.Lock(), threading.Lock()

                def full_step(i):
                    cx = ctxs[i]
                    with upload:
                        cx.set_targets(*tgt); cx.stage_reference(*prot); cx.load_spectra(sp)
                    with compute:
                        cx.step2(); cx.build_reference(); cx.step3(); cx.step4(); cx.rank()
                        if bufs[i] is None:
                            bufs[i] = torch.empty(max(1, 2 * cx.psm_count()) * 96 + 4096, dtype=torch.uint8, pin_memory=True).numpy()
                        nn = len(cx.psms(bufs[i]) if cx.psm_count() * 96 <= bufs[i].nbytes else cx.psms())
                        cx.competitors(3)
                    return nn
                for i in range(2):
                    full_step(i); full_step(i)
                per = max(2, a.e2e_steps)
                ctxs[0].reset_counters(); ctxs[1].reset_counters()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                th = [threading.Thread(target=lambda i=i: [full_step(i) for _ in range(per)]) for i in range(2)]
                for t in th: t.start()
                for t in th: t.join()
                torch.cuda.synchronize()
                w2 = time.perf_counter() - t0
                c0, c1 = ctxs[0].counters(), ctxs[1].counters()
                p2 = (c0.pairs_step2 + c0.pairs_step3 + c0.pairs_step4 + c1.pairs_step2 + c1.pairs_step3 + c1.pairs_step4)
                e2e["two_in_flight"] = {"value": p2 / w2, "unit": UNIT, "ms_per_step": 1000.0 * w2 / (2 * per), "steps": 2 * per,
                                        "note": "two contexts, two host threads; informational, e2e.value is one search at a time"}
                ctxs[0].close(); ctxs[1].close()
            except Exception as e:                 # informational only: never lose the bench line over it
                e2e["two_in_flight"] = {"error": repr(e)[:200]}

    # ---- the other BASELINE.json configs ----------------------------------------------------------------------------
    configs = None
    if not a.no_configs:
        configs = {}
        try:
            if world == 1:
                configs["C1"] = run_small_config(a, mk_params, pinned, gather_psms, size_gather_buffers, timed)
                # C3: the C2 workload with MVH scoring (same spectra, targets, database)
                g, step, _ = resident_run(mk_params(0, method=2), 3, 2)
                tm, _, npsm = timed(step, 3, g)
                c3 = g.counters()
                configs["C3"] = {"workload": "C2 with MVH scoring (-m 2), fixMod C+57 varMod M+16", "value": (c3.pairs_step2 + c3.pairs_step3 + c3.pairs_step4) / 3 / (tm / 3), "unit": UNIT,
                                 "ms_per_step": 1e3 * tm / 3, "pairs_per_step": (c3.pairs_step2 + c3.pairs_step3 + c3.pairs_step4) / 3, "psms": int(npsm),
                                 "pairs_decided_by_score_bound": float(c3.pairs_step4_bounded) / 3,
                                 "kernel_ms": {"k1_step3": c3.ms_score_step[1] / 3, "k1_step4": c3.ms_score_step[2] / 3, "k0": c3.ms_prepare / 3, "k2+k4": c3.ms_shuffle / 3}}
                g.close()
                configs["C4"] = run_c4(a, mk_params, sp, tgt, prot, rank, quiet=True)
                configs["F1_mgf"] = run_mgf(a, mk_params, sp, pinned)
            configs["C5_shard"] = run_c5_shard(a, mk_params, pinned, rank, world, timed, total_over_ranks, gather_psms, size_gather_buffers)
        except Exception as e:                     # the headline line must survive a failing side config
            configs["error"] = repr(e)[:300]

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        os.sched_setaffinity(0, full_affinity)      # the CPU arm gets every host core again
        threads = os.cpu_count() or 1
        ns = min(a.cpu_sample, n_local)
        rate, cp, sec = oracle_sample_rate(mk_params(0), sp, tgt, prot, ns, threads, hoisted=False)
        h_rate, _, h_sec = oracle_sample_rate(mk_params(0), sp, tgt, prot, ns, threads, hoisted=True)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d spectra of the workload x all %d peptides x full DB, Steps 2-4 (%d pairs, %.1f s; reference-table build "
                         "not timed)" % (ns, a.peptides, cp, sec),
               "cost_model": "as written: spectrum pre-processing redone per pair, like PeptideSearchMT.scorePeptide2Spectrum",
               "hoisted": {"value": h_rate, "unit": UNIT, "cores": threads, "seconds": h_sec,
                           "cost_model": "spectrum pre-processing once per spectrum (the best a CPU port can do); same results"},
               "java": java_probe()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warm,
                "ms_per_step": 1000.0 * sec_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a, world),
                           "sharding": (("ONE dataset cut into %d contiguous precursor-mass ranges of equal estimated pair count (shard_bounds/take_shard)" % world) if strong else
                                        ("spectra by precursor-mass strips (%d per GPU, interleaved), tables replicated" % STRIPS_PER_GPU)) if world > 1 else "single GPU",
                           "shard": shard_info,
                           "l2": "inputs larger than L2 (prepared spectra + shuffle rows >> 126 MB per step)",
                           "pairs_per_step": tot_pairs, "psms_per_step": int(n_psm_total),
                           "step4_pairs_decided_by_score_bound": bounded_step,
                           "step4_rule": ("every shuffle scored in full" if a.score_all_shuffles else
                                          "Step 4 needs only shuffleScore >= targetScore (StatisticScoreWorker.java:51): a shuffle whose score "
                                          "upper bound (from its count of possible matches) is below the target score is decided without resolving "
                                          "its matches; counts and p-values identical (tests); value_all_scored is the rate with every shuffle scored in full"),
                           "timed_region": "Steps 2 (incl. spectrum preparation K0)+3+4 + rank/accept + PSM and detail rows to the host, spectra/targets/reference table resident in HBM; "
                                           "e2e: targets + database staging + load (candidate-first gather, K0 and digest under it) + Steps 2-4 + reference build + results to the host",
                           "gen_seconds": round(t_gen, 1), "ms_host_reference_build_once": ms_host_reference},
                "clocks": clk, "gpu_launches": int(launches), "kernel_ms_per_step": kernel_ms, "per_rank_ms_per_step": per_rank_ms,
                "value_all_scored": all_scored, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "configs": configs}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_mgf(a, mk_params, sp, pinned):
    """MGF text -> CSR arrays on the device (pq_parse_mgf_device; text in pinned host memory, the upload is inside the time)
    and text -> loaded spectra (pq_load_mgf), on the text of the first 200 000 spectra of the workload; the host-threaded parser beside it."""
    import torch
    from pepquery_b200 import Search, host, synth
    n = min(200000, len(sp["prec_mz"]))
    end = int(sp["peak_off"][n])
    sub = {"prec_mz": sp["prec_mz"][:n], "charge": sp["charge"][:n], "peak_off": sp["peak_off"][:n + 1], "mz": sp["mz"][:end], "inten": sp["inten"][:end]}
    text = synth.mgf_text(sub, out=lambda nb: pinned((nb + 7) // 8).view(np.uint8)[:nb])
    gb = text.nbytes / 1e9
    g = Search(mk_params(1))
    ptr = lambda t: t.ctypes.data_as(C.c_void_p)
    ns, npk = C.c_int64(0), C.c_uint64(0)
    td = []
    for it in range(4):
        t = text if it % 2 == 0 else text[:text.nbytes - 1]          # (a repeated call on the same text would return the cached parse)
        t0 = time.perf_counter()
        rc = g.L.pq_parse_mgf_device(g.h, ptr(t), C.c_uint64(t.nbytes), C.byref(ns), C.byref(npk), None, None, None, None, None)
        torch.cuda.synchronize()
        td.append(time.perf_counter() - t0)
        if rc != 0:
            raise RuntimeError("pq_parse_mgf_device rc %d" % rc)
    tl = []
    for it in range(3):
        t0 = time.perf_counter(); g.load_mgf(text); torch.cuda.synchronize(); tl.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); host.parse_mgf(bytes(text)); th = time.perf_counter() - t0
    g.close()
    return {"workload": "MGF text of %d spectra (%.2f GB, TITLE/PEPMASS/CHARGE + '%%.4f %%.2f' peak lines), pinned host memory" % (n, gb),
            "device_parse_GBps": gb / min(td[1:]), "device_parse_ms": 1e3 * min(td[1:]), "load_mgf_GBps": gb / min(tl[1:]), "load_mgf_ms": 1e3 * min(tl[1:]),
            "host_parse_GBps": gb / th, "spectra": int(ns.value), "peaks": int(npk.value)}


def run_small_config(a, mk_params, pinned, gather_psms, size_gather_buffers, timed):
    """C1: 100 novel peptides vs 10k spectra + 20k-protein DB, hyperscore, 1000 shuffles (the reference's own CPU-runnable case)."""
    from pepquery_b200 import Search, synth
    prot = synth.proteins(a.proteins); tgt = synth.targets(100)
    sp = synth.spectra(10000, tgt, prot, pinned=pinned)
    for k in ("prec_mz", "charge", "peak_off"):
        sp[k] = pin_copy(sp[k], pinned)
    g = Search(mk_params(1, method=1))

    def step():
        g.set_targets(*tgt); g.stage_reference(*prot); g.load_spectra(sp); g.step2(); g.build_reference(); g.step3(); g.step4()
        return gather_psms(g)
    g.set_targets(*tgt); g.load_spectra(sp); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
    size_gather_buffers(g)
    step(); step()
    g.reset_counters()
    tm, _, npsm = timed(step, 5, g)
    c = g.counters()
    pairs = (c.pairs_step2 + c.pairs_step3 + c.pairs_step4) / 5
    g.close()
    return {"workload": "C1: 100 peptides x 10000 spectra x %d-protein DB, hyperscore, %d shuffles, end to end from host buffers (incl. digest of the database)" % (a.proteins, a.shuffles),
            "value": pairs / (tm / 5), "unit": UNIT, "ms_per_step": 1e3 * tm / 5, "pairs_per_step": pairs, "psms": int(npsm),
            "pairs_by_step": [c.pairs_step2 / 5, c.pairs_step3 / 5, c.pairs_step4 / 5]}


def run_c4(a, mk_params, sp, tgt, prot, rank, quiet=False):
    """C4: Step 5 (ModificationDB.doPTMValidation) over the PSMs of 100k spectra that passed Steps 2-4, full Unimod-sized PTM table."""
    from pepquery_b200 import Search, unimod
    ptms = unimod.load_ptms()
    ns = min(100000, len(sp["prec_mz"]))
    sub = {"prec_mz": sp["prec_mz"][:ns], "charge": sp["charge"][:ns], "peak_off": sp["peak_off"][:ns + 1],
           "mz": sp["mz"][:int(sp["peak_off"][ns])], "inten": sp["inten"][:int(sp["peak_off"][ns])]}
    g = Search(mk_params(1, method=1))
    g.load_spectra(sub); g.set_targets(*tgt); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
    times = []
    for it in range(3):
        g.reset_counters()
        t0 = time.perf_counter()
        g.step5(ptms)
        ps = g.psms()
        dt = time.perf_counter() - t0
        c = g.counters()
        if it >= 1:
            times.append((dt, c.pairs_step5, c.ms_score_step[3], c.bytes_score_step[3], c.pairs_step5_bounded))
    dt = sum(t[0] for t in times) / len(times); pairs = times[-1][1]; kms = sum(t[2] for t in times) / len(times)
    out = {"workload": "C4: unrestricted-modification search, %d PTM specificities, %d spectra x %d peptides (Step 5 only; Steps 2-4 done before)" % (len(ptms), ns, a.peptides),
           "value": pairs / dt, "unit": UNIT, "pairs_step5": int(pairs), "pairs_decided_by_score_bound": int(times[-1][4]), "validated_psms": int((ps["n_ptm"] >= 0).sum()),
           "ms_call": 1e3 * dt, "ms_kernel": kms, "value_kernel": pairs / (kms / 1e3) if kms else None}
    g.close()
    if not quiet and rank == 0:
        print(json.dumps(dict(out, metric=METRIC)), flush=True)
    return out


def run_c5_shard(a, mk_params, pinned, rank, world, timed, total_over_ranks, gather_psms, size_gather_buffers):
    """C5 (100k peptides vs 10M spectra sharded by precursor mass over 8 GPUs): every rank searches 1.25M spectra of its own mass
    strips against all 100k peptides — at 8 GPUs that is the 10M-spectrum sweep, at fewer GPUs the same per-GPU shard."""
    from pepquery_b200 import Search
    per = 1250000
    b = argparse.Namespace(**{**vars(a), "peptides": 100000, "spectra": per})
    prot, tgt, sp = make_inputs(b, rank, max(world, 8) if world == 1 else world, pinned)     # one GPU alone: strips of an 8-way cut
    g = Search(mk_params(1, method=1))

    def step():
        g.set_targets(*tgt); g.stage_reference(*prot); g.load_spectra(sp); g.step2(); g.build_reference(); g.step3(); g.step4()
        return gather_psms(g)
    g.set_targets(*tgt); g.load_spectra(sp); g.step2(); g.build_reference(*prot); g.step3(); g.step4(); g.rank()
    size_gather_buffers(g)
    step()
    g.reset_counters()
    tm, _, npsm = timed(step, 2, g)
    c = g.counters()
    pairs = total_over_ranks((c.pairs_step2 + c.pairs_step3 + c.pairs_step4) / 2)
    out = {"workload": "C5 shard: 100000 peptides x %d spectra per GPU (%d GPUs: %d spectra in total; 10M at 8 GPUs), hyperscore, %d shuffles, end to end from host buffers" %
                       (per, world, per * world, a.shuffles),
           "value": pairs / (tm / 2), "unit": UNIT, "ms_per_step": 1e3 * tm / 2, "pairs_per_step": pairs, "psms_rank0": int(c.psms),
           "h2d_bytes_per_step_rank0": int(c.h2d_bytes / 2), "kernel_ms_rank0": {"k1_step2": c.ms_score_step[0] / 2, "k1_step3": c.ms_score_step[1] / 2, "k1_step4": c.ms_score_step[2] / 2}}
    g.close()
    return out


if __name__ == "__main__":
    main()
