"""Per-launch time of the N = 7 axhelm kernel (three components per launch) and of the gather-scatter
at the benchmark's mesh size; run once per variant with NSB_AX_KB / NSB_AX_DMMA / NSB_AX_STAGES in the
environment (the context reads them at nsb_init)."""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nekstab_next_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nelx', type=int, default=32)
ap.add_argument('--reps', type=int, default=20)
ap.add_argument('--ncomp', type=int, default=3)
ap.add_argument('--check', action='store_true', help='print a checksum of the result (variants must agree bit for bit)')
a = ap.parse_args()

ctx = nb.Context(0)
m = nb.mesh.box_mesh(a.nelx, a.nelx, a.nelx, 7, deform=0.05)
sem = nb.Sem(ctx, 7, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
npts = sem.npts
lay = nb.Layout(ctx, [npts] * a.ncomp, [True] * a.ncomp)
lay.set_weight([sem.get('bm1')] * a.ncomp)
Q = nb.Basis(lay, 2)
rng = np.random.default_rng(0)
op = nb.sem_operator(sem, a.ncomp, 1.0, -1e-4, 1.0, 0.1)
Q[0].upload([rng.standard_normal(npts) for _ in range(a.ncomp)])
for _ in range(3):
    op.matvec(Q[0], Q[1])
ctx.prof_enable(True)
for _ in range(a.reps):
    op.matvec(Q[0], Q[1])
rep = ctx.prof_report()
ctx.prof_enable(False)
tag = ' '.join(f'{k}={os.environ[k]}' for k in ('NSB_AX_KB', 'NSB_AX_DMMA', 'NSB_AX_STAGES') if k in os.environ)
row = []
for name in ('axhelm', 'gather_scatter'):
    v = rep[name]
    row.append(f"{name}: {v['ms'] / v['launches']:.4f} ms {v['bytes'] / v['ms'] / 1e6:7.0f} GB/s")
if a.check:
    out, _ = Q[1].download()
    row.append('checksum %.17e' % float(sum(np.dot(f, f) for f in out)))
print(f'[{tag}]', ' | '.join(row), flush=True)
ctx.close()
