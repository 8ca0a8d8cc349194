"""The pressure-coupled perturbation step (nsb_op_create_ns_stepper) on a 3-D box mesh: linearised Navier-Stokes
about a Taylor-Green flow, BDF3/EXT3, P_N - P_N-2; one operator application = NSTEPS time steps.  Prints the
time per step, the Helmholtz / pressure iteration counts and the cost of one application of the consistent
Poisson operator E = D B^-1 D^T (opgradt -> gather-scatter -> opdiv).

  python profiles/run_ns_stepper.py [--nelx 16] [--nsteps 5] [--tol 1e-6]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nekstab_next_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nelx', type=int, default=16)
ap.add_argument('--nsteps', type=int, default=5)
ap.add_argument('--nu', type=float, default=1e-2)
ap.add_argument('--dt', type=float, default=1e-3)
ap.add_argument('--tol', type=float, default=1e-6)
ap.add_argument('--deform', type=float, default=0.0)
a = ap.parse_args()

ctx = nb.Context(0)
m = nb.mesh.box_mesh(a.nelx, a.nelx, a.nelx, 7, deform=a.deform)
sem = nb.Sem(ctx, 7, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
npts = sem.npts
n2 = sem.pressure_setup()
lay = nb.Layout(ctx, [npts] * 3 + [n2], [True] * 3 + [False])
lay.set_weight([sem.get('bm1')] * 3)
Q = nb.Basis(lay, 4)
x, y, z = (m[k].ravel() for k in 'xyz')
tp = 2 * np.pi
Q[0].upload([np.sin(tp * x) * np.cos(tp * y) * np.cos(tp * z), -np.cos(tp * x) * np.sin(tp * y) * np.cos(tp * z),
             0 * x, np.zeros(n2)])
sem.dealias_setup()
rng = np.random.default_rng(0)
Q[1].upload([rng.standard_normal(npts) for _ in range(3)] + [np.zeros(n2)])
for f in range(3):
    sem.dssum(Q[1], f)
    sem.col2(Q[1], f, 'vmult')
    sem.col2(Q[1], f, 'mask')
# E apply
Q[3].upload([np.zeros(npts)] * 3 + [rng.standard_normal(n2)])
sem.cdabdtp(Q[3], Q[2])
ctx.sync()
t0 = time.perf_counter()
for _ in range(20):
    sem.cdabdtp(Q[3], Q[2])
ctx.sync()
te = (time.perf_counter() - t0) / 20
print(f'E = D B^-1 D^T on {n2} pressure / 3 x {npts} velocity points: {te * 1e3:.3f} ms per application')
op = nb.ns_stepper_operator(sem, lay, Q[0], a.nu, a.dt, a.nsteps, tol_v=a.tol, tol_p=a.tol, maxit=2000,
                            mean_free=(a.deform == 0.0))
op.matvec(Q[1], Q[2])                     # warm-up (allocations, preconditioner diagonal)
ctx.sync()
h0, p0 = nb.ns_iterations(op)
t0 = time.perf_counter()
op.matvec(Q[1], Q[2])
ctx.sync()
dt = time.perf_counter() - t0
h1, p1 = nb.ns_iterations(op)
sem.opdiv(Q[2], Q[3])
div = np.max(np.abs(Q[3].download()[0][3]))
print(f'ns stepper: {a.nsteps} steps, 3 x {npts} velocity + {n2} pressure points, tol {a.tol:g}: {dt * 1e3:.1f} ms = '
      f'{dt / a.nsteps * 1e3:.2f} ms/step; per step {(h1 - h0) / a.nsteps / 3:.1f} Helmholtz iterations per component, '
      f'{(p1 - p0) / a.nsteps:.1f} pressure iterations ({(dt / a.nsteps) / max((p1 - p0) / a.nsteps, 1) * 1e3:.3f} ms per '
      f'pressure iteration if they were all of it); max |D v| = {div:.2e}; |out| = {nb.k_norm(Q[2]):.6e}')
ctx.close()
