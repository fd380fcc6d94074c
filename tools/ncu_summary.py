#!/usr/bin/env python
"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep} into the text summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
           'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
           'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
           'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ,
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] == 'ns' else v * 1000 if r[ui] == 'ms' else v
        agg[r[ki].split('(')[0]].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, 'w') as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n# source: {path}\n")
        f.write(f"{'kernel':52s} {'launches':>8s} {'avg_us':>10s} {'total_us':>11s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:52]:52s} {len(v):8d} {sum(v) / len(v):10.2f} {sum(v):11.1f} {sum(v) / tot * 100:6.1f}%\n")


def kernels(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name')
    seen = collections.OrderedDict()
    for r in data:
        seen.setdefault(r[ki].split('(')[0], r)
    with open(out, 'w') as f:
        f.write(f"# ncu --set full --clock-control none; first captured launch of each kernel\n# source: {rep}\n")
        for name, r in seen.items():
            f.write(f"\n== {name}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"   {m:86s} {r[i]:>16s} {units[i]}\n")


def traffic(rep, out):
    """dram bytes per launch of every profiled kernel (first captured launch) -> profiles/traffic.json for bench.py"""
    import json
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, ri, wi = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    res = {}
    for r in data:
        name = r[ki].split('(')[0].replace('void ', '').split('<')[0]
        if name not in res:
            res[name] = float(r[ri]) * scale.get(units[ri], 1.0) + float(r[wi]) * scale.get(units[wi], 1.0)
    json.dump({'source': rep, 'note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (cold caches)',
               'bytes_per_launch': res}, open(out, 'w'), indent=1)


if __name__ == '__main__':
    tag = sys.argv[1]
    traffic(f'gpurun_out/prof_{tag}.ncu-rep', 'profiles/traffic.json')
    launches(f'gpurun_out/launches_{tag}.csv', f'profiles/{tag}_launches.txt')
    kernels(f'gpurun_out/prof_{tag}.ncu-rep', f'profiles/{tag}_kernels.txt')
