"""Prints the key counters of every kernel in an .ncu-rep (raw page) and, with --src, the hottest source lines.
Usage: python profiles/ncu_summary.py report.ncu-rep [--src KERNEL_REGEX]"""
import csv
import io
import subprocess
import sys
from collections import Counter, defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('##', r[hdr.index('Kernel Name')][:70])
        for k in KEYS:
            if k in hdr:
                print(f'   {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
        stalls = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
        for h, v in sorted(stalls, key=lambda t: -t[1])[:6]:
            print(f'   stall {h.split("issue_stalled_")[1].split("_per_issue")[0]:30s} {v:8.2f}')


def src(rep, regex, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + regex, '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # find the header row
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    ex, st = col['Instructions Executed'], col['Warp Stall Sampling (All Samples)']
    key = col.get('Source', 1)
    agg = defaultdict(lambda: [0, 0])
    for r in rows[hi + 1:]:
        if len(r) <= max(ex, st):
            continue
        try:
            agg[r[key][:110]][0] += int(r[ex] or 0)
            agg[r[key][:110]][1] += int(r[st] or 0)
        except ValueError:
            pass
    ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    print(f'total inst {ti} samples {ts}')
    for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{100 * s / max(ts, 1):5.1f}% stall {100 * i / max(ti, 1):5.1f}% inst | {k}')


if __name__ == '__main__':
    if '--src' in sys.argv:
        src(sys.argv[1], sys.argv[sys.argv.index('--src') + 1])
    else:
        raw(sys.argv[1])
