"""profiles/k1_step4_profile.json from an `ncu --set full` capture of bench.py's Step-4 K1 launch (k1_score_kernel<128, METHOD>; <256, METHOD> with PQ_STEP4_BLOCK=256):
DRAM bytes, warp instructions and issue utilisation of ONE launch.  bench.py divides the first two by the launch time it measures live.

    python tools/make_step4_profile.py gpurun_out/prof_TAG.ncu-rep profiles/TAG_k1_ncu_full.txt"""
import csv, io, json, os, subprocess, sys

rep, src_note = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
pick = [r for r in rows[2:] if "k1_score_kernel<128" in r[col["Kernel Name"]].replace("(int)", "") or "k1_score_kernel<256" in r[col["Kernel Name"]].replace("(int)", "")]
assert pick, "no k1_score_kernel<128 / 256, ...> launch in the capture"
r = pick[-1]
unit = {h: rows[1][i] for h, i in col.items()}


def val(name):
    v = float(r[col[name]]); u = unit[name]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


sm_count = int(float(r[col["device__attribute_multiprocessor_count"]])) if "device__attribute_multiprocessor_count" in col else 148
d = {"kernel": r[col["Kernel Name"]].split("(")[0] + " (Step-4 launch)",
     "workload": {"spectra": 1000000, "peptides": 10000, "proteins": 20000, "shuffles": 1000, "method": 1},
     "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
     "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
     "warp_inst_per_launch": float(r[col["smsp__inst_executed.sum"]]),
     "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
     # the L1 / shared-memory data pipe (LSU wavefronts; peak = one per SM and cycle): the unit this kernel loads most
     "lsu_wavefronts_per_launch": float(r[col["SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg"]]) * sm_count,
     "lsu_wavefronts_shared_per_launch": float(r[col["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]]),
     "lsu_wavefronts_pct": float(r[col["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]]),
     "sm_count": sm_count,
     "ncu_duration_ms": float(r[col["gpu__time_duration.sum"]]),
     "registers_per_thread": float(r[col["launch__registers_per_thread"]]),
     "source": "ncu --set full --clock-control none, %s (one launch of bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-configs)" % src_note}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "k1_step4_profile.json")
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d, indent=1))
