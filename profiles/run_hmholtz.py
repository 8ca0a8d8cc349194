"""Helmholtz PCG solve (nsb_sem_hmholtz) on the benchmark mesh: time per iteration."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nekstab_next_b200 as nb  # noqa: E402

nelx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = nb.Context(0)
m = nb.mesh.box_mesh(nelx, nelx, nelx, 7, deform=0.05)
sem = nb.Sem(ctx, 7, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
npts = sem.npts
lay = nb.Layout(ctx, [npts], [True])
lay.set_weight([sem.get('bm1')])
B = nb.Basis(lay, 3)
rng = np.random.default_rng(0)
B[0].upload([rng.standard_normal(npts)])
sem.dssum(B[0], 0)
sem.col2(B[0], 0, 'mask')
for tol, maxit in ((1e-30, 50), (1e-8, 2000)):
    ctx.sync()
    t0 = time.perf_counter()
    it, res = sem.hmholtz(B[0], B[1], 0, 1.0 / 100.0, 1.0 / 1e-2, tol=tol, maxit=maxit)   # h1 = 1/Re, h2 = 1/dt
    ctx.sync()
    dt = time.perf_counter() - t0
    print(f'npts={npts} tol={tol:g}: {it} iterations, residual drop {res:.2e}, {dt * 1e3:.1f} ms, '
          f'{dt / it * 1e3:.3f} ms/iteration, {npts * it / dt / 1e9:.2f} GDOF/s per iteration')

# three systems side by side (Nek's ophinv): one axhelm + one gather-scatter launch per iteration for all
lay3 = nb.Layout(ctx, [npts] * 3, [True] * 3)
B3 = nb.Basis(lay3, 2)
B3[0].upload([rng.standard_normal(npts) for _ in range(3)])
for f in range(3):
    sem.dssum(B3[0], f)
    sem.col2(B3[0], f, 'mask')
for _ in range(2):
    ctx.sync()
    t0 = time.perf_counter()
    its, ress = sem.hmholtz_vec(B3[0], B3[1], 0, 3, 1.0 / 100.0, 1.0 / 1e-2, tol=1e-30, maxit=50)
    ctx.sync()
    dt = time.perf_counter() - t0
print(f'3 systems side by side: {max(its)} iterations, {dt * 1e3:.1f} ms, {dt / max(its) * 1e3:.3f} ms/iteration '
      f'({dt / max(its) / 3 * 1e3:.3f} ms per system and iteration)')
