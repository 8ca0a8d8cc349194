"""Small driver for ncu: the hot kernels of one Arnoldi step at the benchmark's full mesh size.

  python profiles/run_hot_kernels.py [--k 50] [--nelx 32] [--reps 3]

Launches, per repetition: axhelm (three components in one launch) + gather-scatter, then the CGS2
orthonormalisation against k columns: multidot, fused update+multidot, update(+norm), normalize.
Use `ncu --profile-from-start off`: only the last repetition sits between cudaProfilerStart/Stop.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nekstab_next_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--k', type=int, default=50)
ap.add_argument('--nelx', type=int, default=32)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--ncomp', type=int, default=3)
a = ap.parse_args()

ctx = nb.Context(0)
m = nb.mesh.box_mesh(a.nelx, a.nelx, a.nelx, 7, deform=0.05)
sem = nb.Sem(ctx, 7, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
npts = sem.npts
bm1 = sem.get('bm1')
lay = nb.Layout(ctx, [npts] * a.ncomp, [True] * a.ncomp)
lay.set_weight([bm1] * a.ncomp)
Q = nb.Basis(lay, a.k + 2)
rng = np.random.default_rng(0)
op = nb.sem_operator(sem, a.ncomp, 1.0, -1e-4, 1.0, 0.1)
Q[0].upload([rng.standard_normal(npts) for _ in range(a.ncomp)])
nb.k_normalize(Q[0])
H = np.zeros((a.k + 2, a.k + 1), order='F')
nb.arnoldi_factorization(Q, H, 1, a.k, a.k + 1, op)      # fills k+1 orthonormal columns
ctx.sync()
lib = nb.load()
for r in range(a.reps):
    if r == a.reps - 1:
        lib.nsb_profiler_start()       # ncu --profile-from-start off captures the last repetition only
    op.matvec(Q[a.k - 1], Q[a.k + 1])
    h, _ = nb.orthonormalize(Q, a.k, a.k + 1, nb.ORTH_CGS2)
ctx.sync()
lib.nsb_profiler_stop()
print('ok', float(h[-1]))
