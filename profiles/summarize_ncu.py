#!/usr/bin/env python
"""Markdown summaries of the two ncu passes of profiles/run_r02_evidence.sh.

  python profiles/summarize_ncu.py launches <launches.csv[.gz]> <bench.json>   > profiles/ncu_r02_launch_list_summary.md
  python profiles/summarize_ncu.py kernels  <hot_k100_raw.csv>                 > profiles/ncu_r02_hot_kernels_k100.md

`launches`: per-kernel totals of `ncu --metrics gpu__time_duration.sum --clock-control none` over the bench command
(serialised, cold-cache launches: only the SHARES are comparable with the CUDA-event shares of bench.py's roofline
table, which are printed next to them).  `kernels`: the `--set full` capture of profiles/run_hot_kernels.py --k 100.
"""
import collections
import csv
import gzip
import io
import json
import sys


def read(path):
    f = io.TextIOWrapper(gzip.open(path)) if path.endswith('.gz') else open(path)
    rows = list(csv.reader(f))
    i = [k for k, r in enumerate(rows) if r and r[0] == 'ID'][0]
    return rows[i], rows[i + 1:]


def short(name):
    return name.split('(')[0].replace('void ', '').replace('<unnamed>::', '').strip()


CLASS = [('multidot_kernel', 'multidot'), ('fused_', 'fused_update_dot'), ('update_kernel<0', 'update'),
         ('update_kernel<2', 'update'), ('normalize_kernel', 'normalize'), ('axhelm', 'axhelm'), ('gs_kernel', 'gather_scatter'),
         ('halo', 'gather_scatter'), ('blas1_kernel', 'blas1'), ('wdot_kernel', 'blas1')]


def launches(path, bench):
    h, rows = read(path)
    kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}
    t, c = collections.defaultdict(float), collections.Counter()
    for r in rows:
        if len(r) <= mv or not r[mv]:
            continue
        n = short(r[kn])
        t[n] += float(r[mv].replace(',', '')) * scale.get(r[mu], 1.0)
        c[n] += 1
    tot = sum(t.values())
    print(f'# ncu launch list of `bench.py --steps 1 --warmup 1` (round 2, final code)\n')
    print(f'{sum(c.values())} launches, {tot:.1f} ms of kernel time under ncu (serialised, cold cache).\n')
    print('| kernel | launches | total ms | ms each | share |')
    print('|---|---:|---:|---:|---:|')
    for n, v in sorted(t.items(), key=lambda x: -x[1])[:18]:
        print(f'| `{n}` | {c[n]} | {v:.1f} | {v / c[n]:.3f} | {100 * v / tot:.1f} % |')
    cls = collections.defaultdict(float)
    for n, v in t.items():
        for pat, k in CLASS:
            if pat in n:
                cls[k] += v
                break
    b = json.load(open(bench))['roofline']['kernels']
    ctot = sum(cls[k] for k in b if k in cls) or 1.0
    btot = sum(v['ms'] for k, v in b.items() if k in cls) or 1.0
    print('\nShares by kernel class, ncu vs the CUDA-event profile of the same command in `bench.py` (un-profiled run):\n')
    print('| class | ncu share | CUDA-event share (bench.py) |')
    print('|---|---:|---:|')
    for k, v in b.items():
        if k in cls:
            print(f'| {k} | {100 * cls[k] / ctot:.1f} % | {100 * v["ms"] / btot:.1f} % |')


def kernels(path):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    want = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
            ('launch__registers_per_thread', 'regs'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
            ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared pipe %'),
            ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64 pipe %')]
    want = [(m, t) for m, t in want if m in col]
    print('# ncu --set full, hot kernels of one Arnoldi step at k = 100 (box 32^3, N = 7, 3 components; round 2)\n')
    print('| kernel | grid | ' + ' | '.join(f'{t} [{u[col[m]]}]' for m, t in want) + ' | DRAM GB/s |')
    print('|---|---|' + '---:|' * (len(want) + 1))
    unit = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    tsc = {'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 'msecond': 1e-3, 'usecond': 1e-6, 'second': 1.0}
    for r in rows[2:]:
        vals = [r[col[m]] for m, _ in want]
        rd = float(r[col['dram__bytes_read.sum']]) * unit[u[col['dram__bytes_read.sum']]]
        wr = float(r[col['dram__bytes_write.sum']]) * unit[u[col['dram__bytes_write.sum']]]
        tt = float(r[col['gpu__time_duration.sum']]) * tsc[u[col['gpu__time_duration.sum']]]
        print(f"| `{short(r[col['Kernel Name']])}` | {r[col['Grid Size']]} | " + ' | '.join(f'{float(v):.4g}' for v in vals) +
              f' | {(rd + wr) / tt / 1e9:.0f} |')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        kernels(sys.argv[2])
