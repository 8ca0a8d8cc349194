#!/usr/bin/env python
"""profiles/ncu_traffic.json from one `ncu --set full` capture of profiles/run_hot_kernels.py --k K.

  python profiles/make_ncu_traffic.py gpurun_out/r2ev/hot_k100_raw.csv gpurun_out/r2ev/bench.json --k 100

The CSV is `ncu -i <rep> --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,...` of the capture
(profiles/run_r02_evidence.sh).  Per kernel class of bench.py's roofline table the file keeps the measured DRAM
bytes of the captured launch next to the algorithmic bytes of that launch, so that bench.py can report
`roofline.traffic` from a counter instead of a constant.  Algorithmic bytes: DESIGN.md section 4 (orthogonalisation
kernels, from n and k); for the two SEM kernels the per-launch figure bench.py itself accounts (they do not depend
on k), read from the bench line given as second argument.
"""
import argparse
import csv
import json
from pathlib import Path

ap = argparse.ArgumentParser()
ap.add_argument('csv')
ap.add_argument('bench')
ap.add_argument('--k', type=int, default=100)
ap.add_argument('--out', default=str(Path(__file__).resolve().parent / 'ncu_traffic.json'))
a = ap.parse_args()

bench = json.loads(Path(a.bench).read_text())
n = bench['config']['ndof']
k = a.k
unit = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
classes = [('axhelm3d', 'axhelm', None), ('gs_kernel', 'gather_scatter', None),
           ('multidot_kernel', 'multidot', 8.0 * n * (k + 2)),
           ('fused_', 'fused_update_dot', 8.0 * (n * (k + 2) + n)),
           ('update_kernel<2', 'update', 8.0 * n * (k + 2)),          # third sweep with the folded normalisation: no W
           ('update_kernel', 'update', 8.0 * (n * (k + 2) + n)),
           ('normalize_kernel', 'normalize', 16.0 * n)]
rows = list(csv.reader(open(a.csv)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
out = {}
for r in rows[2:]:
    name = r[col['Kernel Name']]
    for pat, cls, alg in classes:
        if pat in name and cls not in out:
            rd = float(r[col['dram__bytes_read.sum']]) * unit[units[col['dram__bytes_read.sum']]]
            wr = float(r[col['dram__bytes_write.sum']]) * unit[units[col['dram__bytes_write.sum']]]
            if alg is None:
                kv = bench['roofline']['kernels'][cls]
                alg = kv['achieved_gbs'] * 1e9 * kv['ms'] * 1e-3 / kv['launches']
            out[cls] = dict(kernel=name.split('(')[0].replace('void <unnamed>::', '').strip(), k=k, n=n,
                            dram_read_bytes=rd, dram_write_bytes=wr, dram_bytes=rd + wr, algorithmic_bytes=alg,
                            ratio=round((rd + wr) / alg, 4), ncu_ms=float(r[col['gpu__time_duration.sum']]),
                            source=f'profiles/{Path(a.csv).name} (ncu --set full --clock-control none, '
                                   f'run_hot_kernels.py --k {k})')
Path(a.out).write_text(json.dumps(out, indent=1) + '\n')
for c, v in out.items():
    print(f"{c:18s} dram {v['dram_bytes'] / 1e9:8.3f} GB  algorithmic {v['algorithmic_bytes'] / 1e9:8.3f} GB  ratio {v['ratio']:.3f}")
