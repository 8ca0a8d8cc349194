"""The device time-stepper operator (nsb_op_create_stepper) at the benchmark's mesh size: three scalars
advected by a Taylor-Green flow, BDF3/EXT3, one operator application = NSTEPS time steps.

  python profiles/run_stepper.py [--nelx 32] [--nsteps 10]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nekstab_next_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nelx', type=int, default=32)
ap.add_argument('--nsteps', type=int, default=10)
ap.add_argument('--kappa', type=float, default=1e-2)
ap.add_argument('--dt', type=float, default=1e-3)
a = ap.parse_args()

ctx = nb.Context(0)
m = nb.mesh.box_mesh(a.nelx, a.nelx, a.nelx, 7, deform=0.05)
sem = nb.Sem(ctx, 7, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
npts = sem.npts
lay = nb.Layout(ctx, [npts] * 3, [True] * 3)
lay.set_weight([sem.get('bm1')] * 3)
Q = nb.Basis(lay, 3)
x, y, z = (m[k].ravel() for k in 'xyz')
tp = 2 * np.pi
Q[0].upload([np.sin(tp * x) * np.cos(tp * y) * np.cos(tp * z), -np.cos(tp * x) * np.sin(tp * y) * np.cos(tp * z),
             0 * x])
sem.dealias_setup()
sem.set_convect(0, Q[0])
rng = np.random.default_rng(0)
Q[1].upload([rng.standard_normal(npts) for _ in range(3)])
for f in range(3):
    sem.dssum(Q[1], f)
    sem.col2(Q[1], f, 'vmult')
    sem.col2(Q[1], f, 'mask')
op = nb.stepper_operator(sem, lay, 3, 0, a.kappa, a.dt, a.nsteps, tol=1e-8, maxit=500)
op.matvec(Q[1], Q[2])                     # warm-up (allocations, preconditioner diagonal)
ctx.sync()
l0 = ctx.launch_count() if hasattr(ctx, 'launch_count') else 0
t0 = time.perf_counter()
op.matvec(Q[1], Q[2])
ctx.sync()
dt = time.perf_counter() - t0
print(f'stepper: {a.nsteps} steps of 3 scalars on {npts} points in {dt * 1e3:.1f} ms = {dt / a.nsteps * 1e3:.2f} ms/step '
      f'({3 * npts * a.nsteps / dt / 1e9:.2f} GDOF-steps/s); |out| = {nb.k_norm(Q[2]):.6e}')
ctx.close()
