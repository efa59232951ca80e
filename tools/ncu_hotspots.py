#!/usr/bin/env python
"""Where does a kernel's time go?  Reads an `ncu --set full --import-source on` report (run here, no GPU needed) and
prints the warp-stall samples bucketed along the SASS plus the hottest instructions.

    python tools/ncu_hotspots.py gpurun_out/prof_rollout.ncu-rep [bucket=100]
"""
import csv
import io
import subprocess
import sys

MARKS = ('LDGSTS', 'FFMA', 'HMMA', 'ERRBAR', 'UCGABAR_ARV', 'BAR.SYNC', 'STG', 'LDG', 'MUFU', 'SHFL', 'LDS', 'STS',
         'ST.E', 'CALL', 'MEMBAR', 'ATOM', 'RED')


def main():
    rep = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    if len(r) > 2:
        want = ('gpu__time_duration.sum', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'sm__cycles_active.avg',
                'sm__cycles_elapsed.max', 'sm__warps_active.avg.pct', 'launch__grid_size', 'launch__registers_per_thread',
                'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'smsp__average_warps_issue_stalled')
        for k, x in zip(r[0], r[2]):
            if any(k.startswith(w) for w in want) and not k.endswith(('.per_second', 'pct_of_peak_sustained_elapsed')):
                if 'issue_stalled' in k and float(x or 0) < 0.3:
                    continue
                print('%-90s %s' % (k, x))
    sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True,
                          text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    hdr, data = rows[1], rows[2:]
    iS, iI, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
    tot = sum(int(x[iS]) for x in data if x[iS].isdigit())
    print('total samples', tot)
    for b in range(0, len(data), B):
        chunk = data[b:b + B]
        s = sum(int(x[iS]) for x in chunk if x[iS].isdigit())
        n = sum(int(x[iI]) for x in chunk if x[iI].isdigit())
        marks = set()
        for x in chunk:
            op = [o for o in x[iSrc].strip().split() if not o.startswith('@')]
            op = op[0] if op else ''
            for m in MARKS:
                if op.startswith(m):
                    marks.add(m)
        if s > tot * 0.01:
            print('sass %5d..  samples %6d (%4.1f%%)  instr %10d  %s' % (b, s, 100.0 * s / tot, n, sorted(marks)))
    top = sorted([(int(x[iS]), int(x[iI]), i, x[iSrc].strip()) for i, x in enumerate(data) if x[iS].isdigit()], reverse=True)
    for s, n, i, src in top[:20]:
        print('%6d samples  %9d exec  @%5d  %s' % (s, n, i, src[:80]))


if __name__ == '__main__':
    main()
