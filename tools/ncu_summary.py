"""Summaries of ncu exports for profiles/: launch list shares, key raw metrics, per-source-line instruction shares."""
import collections, csv, io, subprocess, sys

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']


def launches(path):
    txt = open(path).read().splitlines()
    i = [k for k, l in enumerate(txt) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(l for l in txt[i:] if l.startswith('"')))))
    agg = collections.OrderedDict()
    for r in rows:
        k = r['Kernel Name'].split('(')[0][-90:]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r['Metric Value'])
    tot = sum(v[1] for v in agg.values())
    s = "%d launches, %.3f ms total device time (cold-cache, serialised: compare shares)\n" % (len(rows), tot / 1e6)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        s += "%-92s n=%3d %10.3f ms %5.1f%%\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot)
    return s, rows


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    s = ""
    for r in rows[2:]:
        s += "-----\n"
        for wn in KEEP:
            for i, h in enumerate(hdr):
                if h == wn:
                    s += "%-86s %s %s\n" % (h, r[i], units[i])
    return s


def source(rep, pattern, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    tot = collections.OrderedDict(); smp = {}
    fp = cur = None
    for r in rows:
        if not r: continue
        if r[0] == 'File Path': fp = r[1].split('/')[-1]; continue
        if r[0] == 'Function Name': cur = r[1]; continue
        if r[0] == 'Line No' or pattern not in (cur or ''): continue
        if r[0] != '':
            try:
                key = (fp, int(r[0]), r[1].strip()[:120]); tot[key] = tot.get(key, 0) + int(r[7]); smp[key] = smp.get(key, 0) + int(r[6])
            except Exception:
                pass
    T = sum(tot.values()) or 1; S = sum(smp.values()) or 1
    s = "function pattern %r: %d warp-instructions executed (first matching launches summed)\n inst%%  stall-samples%%  file:line  source\n" % (pattern, T)
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        s += "%5.1f%% %5.1f%% %s:%d  %s\n" % (100 * v / T, 100 * smp[k] / S, k[0], k[1], k[2])
    return s


def step_shares(path):
    """Per-kernel shares of ONE resident step of `bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-configs`: the last step whose
    Step-4 launch took < 15 ms (the all-scored arm takes > 30 ms), from its k_enumerate_targets to the kernel before the next load
    (k_virtual_meta) or the next step."""
    _, rows = launches(path)
    names = [r['Kernel Name'].split('(')[0] for r in rows]
    s4 = [i for i, n in enumerate(names) if ('k1_score_kernel<128' in n or 'k1_score_kernel<256' in n) and float(rows[i]['Metric Value']) < 15e6]
    e = s4[-1]
    a = max(i for i, n in enumerate(names) if 'k_enumerate_targets' in n and i < e)
    b = min([i for i, n in enumerate(names) if ('k_enumerate_targets' in n or 'k_virtual_meta' in n) and i > e] + [len(rows)])
    agg = collections.OrderedDict()
    for r in rows[a:b]:
        k = r['Kernel Name'].split('(')[0][-70:]
        v = agg.setdefault(k, [0, 0.0]); v[0] += 1; v[1] += float(r['Metric Value'])
    tot = sum(v[1] for v in agg.values())
    out = "one resident step (bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-configs under ncu, launches %d..%d): %d launches, %.3f ms device time (cold-cache, serialised: compare shares)\n" % (a, b - 1, b - a, tot / 1e6)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out += "%-72s n=%3d %9.3f ms %5.1f%%\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot)
    return out


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches": print(launches(sys.argv[2])[0])
    elif cmd == "step-shares": print(step_shares(sys.argv[2]), end="")
    elif cmd == "raw": print(raw(sys.argv[2]))
    elif cmd == "source": print(source(sys.argv[2], sys.argv[3]))
