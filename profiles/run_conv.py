"""Timing (CUDA events) of the time-stepper pieces around ax at the benchmark's mesh size:
dealiased convection (set_convect, convect on 3 components) and the fused EXT/BDF pass.

  python profiles/run_conv.py [--nelx 32] [--reps 10]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nekstab_next_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nelx', type=int, default=32)
ap.add_argument('--reps', type=int, default=10)
a = ap.parse_args()

ctx = nb.Context(0)
m = nb.mesh.box_mesh(a.nelx, a.nelx, a.nelx, 7, deform=0.05)
sem = nb.Sem(ctx, 7, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
npts = sem.npts
nel = a.nelx ** 3
lay = nb.Layout(ctx, [npts] * 3, [True] * 3)
lay.set_weight([sem.get('bm1')] * 3)
Q = nb.Basis(lay, 7)
rng = np.random.default_rng(0)
for c in range(6):
    Q[c].upload([rng.standard_normal(npts) for _ in range(3)])
ctx.timer_start()
sem.dealias_setup()
t_setup = ctx.timer_stop()


def timed(fn):
    for _ in range(2):
        fn()
    ctx.timer_start()
    for _ in range(a.reps):
        fn()
    return ctx.timer_stop() / a.reps


nfine = nel * 12 ** 3
t_sc = timed(lambda: sem.set_convect(0, Q[0]))
t_cv = timed(lambda: sem.convect(0, Q[1], Q[6], 0, 3))
t_cva = timed(lambda: sem.convect(0, Q[1], Q[6], 0, 3, scale=-1.0, accumulate=True))
ab, bd = [23 / 12, -16 / 12, 5 / 12], [11 / 6, -3.0, 1.5, -1 / 3]
t_be = timed(lambda: sem.bdf_ext(Q[0], Q[1], Q[2], [Q[3], Q[4], Q[5]], ab, bd, 500.0, 0, 3))
gf = 2.0 * 3 * nel * (8 * (768 + 1152 + 1728) * 2 + 3 * 12 * 1728) * 1e-9   # flop of one 3-component convect
print(f'dealias_setup {t_setup:.1f} ms (once per mesh; fine metrics {9 * nfine * 8 / 1e9:.2f} GB)')
print(f'set_convect   {t_sc:.3f} ms   reads {(3 * npts + 9 * nfine) * 8 / 1e9:.2f} GB, writes {3 * nfine * 8 / 1e9:.2f} GB')
print(f'convect x3    {t_cv:.3f} ms   {gf / t_cv:.2f} TFLOP/s fp64; '
      f'algorithmic bytes {(6 * npts + 3 * nfine) * 8 / 1e9:.2f} GB -> {(6 * npts + 3 * nfine) * 8 / t_cv / 1e6:.0f} GB/s')
print(f'convect x3 accumulate {t_cva:.3f} ms')
print(f'bdf_ext x3    {t_be:.3f} ms   {8.0 * npts * (3 * 9 + 1) / t_be / 1e6:.0f} GB/s')
ctx.close()
