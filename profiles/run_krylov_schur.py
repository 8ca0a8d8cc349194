"""Krylov-Schur (nsb_krylov_schur) on the benchmark mesh: exercises the restart path at full size
(dense Schur / reorder on the host, basis rotation Q <- Q Z on the device).

  python profiles/run_krylov_schur.py [--nelx 32] [--kdim 100] [--restarts 2] [--conv]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nekstab_next_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nelx', type=int, default=32)
ap.add_argument('--kdim', type=int, default=100)
ap.add_argument('--restarts', type=int, default=2)
ap.add_argument('--ncomp', type=int, default=3)
ap.add_argument('--conv', action='store_true')
ap.add_argument('--tol', type=float, default=1e-6)
a = ap.parse_args()

ctx = nb.Context(0)
m = nb.mesh.box_mesh(a.nelx, a.nelx, a.nelx, 7, deform=0.05)
sem = nb.Sem(ctx, 7, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
npts = sem.npts
bm1 = sem.get('bm1')
lay = nb.Layout(ctx, [npts] * a.ncomp, [True] * a.ncomp)
lay.set_weight([bm1] * a.ncomp)
Q = nb.Basis(lay, a.kdim + 1)
conv = nb.mesh.taylor_green(m['x'], m['y'], m['z']) if a.conv else None
rng = np.random.default_rng(7)
Q[0].upload([rng.standard_normal(npts) for _ in range(a.ncomp)])
for f in range(a.ncomp):
    sem.dssum(Q[0], f)
    sem.col2(Q[0], f, 'vmult')
    sem.col2(Q[0], f, 'mask')
nb.k_normalize(Q[0])
Lop = nb.sem_operator(sem, a.ncomp, 0.0, 1.0, 1.0, 0.1, conv=conv)
nb.k_copy(Q[1], Q[0])
rho = 1.0
for _ in range(15):
    Lop.matvec(Q[1], Q[2])
    rho = nb.k_normalize(Q[2])
    nb.k_copy(Q[1], Q[2])
op = nb.sem_operator(sem, a.ncomp, 1.0, -1.0 / (1.05 * rho), 1.0, 0.1, conv=conv)
ctx.prof_enable(True)
t0 = time.perf_counter()
res = nb.krylov_schur(Q, op, k_dim=a.kdim, schur_tgt=2, eigen_tol=a.tol, schur_del=0.10, max_restarts=a.restarts)
ctx.sync()
wall = time.perf_counter() - t0
rep = ctx.prof_report()
ctx.prof_enable(False)
G = Q.gram(min(a.kdim, 24))
out = dict(ndof=a.ncomp * npts, k_dim=a.kdim, restarts=res.schur_cnt, converged=res.cnt, wall_s=wall,
           matvecs=op.count(), leading_ritz=[[float(v.real), float(v.imag)] for v in res.vals[:6]],
           residuals=[float(r) for r in res.residual[:6]],
           orth_first24=float(np.max(np.abs(G - np.eye(G.shape[0])))),
           kernels={k: dict(ms=round(v['ms'], 2), launches=v['launches'],
                            gbs=round(v['bytes'] / v['ms'] / 1e6, 1)) for k, v in rep.items()})
print(json.dumps(out))
