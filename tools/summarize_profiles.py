#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run here, no GPU needed).

    python tools/summarize_profiles.py r01a      # tag = round + capture id
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max']


def launches(tag):
    path = os.path.join(OUT, 'launches.csv')
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    n = 0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000.0 if u in ('ns', 'nsecond') else (v * 1000.0 if u in ('ms', 'msecond') else v)
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('<unnamed>::', '').replace('void ', '')[:64]
        agg[name][0] += 1
        agg[name][1] += v
        total += v
        n += 1
    with open(os.path.join(PROF, '%s_launches.md' % tag), 'w') as f:
        f.write('# ncu launch list of ONE bench step (eager launches, `--metrics gpu__time_duration.sum '
                '--clock-control none`)\n\n')
        f.write('Command: `SB200_CUDA_GRAPH=0 ncu ... python bench.py --steps 1 --warmup 3 --lite` '
                '(see tools/profile_gpu.sh).  Per-launch times are cold-cache and serialised: compare SHARES.\n\n')
        f.write('%d launches, %.1f us total\n\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n' % (n, total))
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.2f | %.1f%% |\n' % (k, c, t, t / c, 100 * t / total))
    print('wrote launches summary (%d launches)' % n)


def report(name, tag):
    rep = os.path.join(OUT, name + '.ncu-rep')
    if not os.path.exists(rep):
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, '%s_%s.md' % (tag, name)), 'w') as f:
        f.write('# ncu --set full: %s\n\n' % name)
        for vals in rows[2:]:
            kn = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
            f.write('## %s\n\n| metric | value | unit |\n|---|---:|---|\n' % kn[:120])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write('| %s | %s | %s |\n' % (k, vals[i], units[i]))
            f.write('\n')
    print('wrote', name)


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    for n in ('prof_critic', 'prof_gae', 'prof_skinny'):
        report(n, tag)
