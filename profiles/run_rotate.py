"""Basis rotation Q <- Q Z at the benchmark size (n = 50.3 M rows, k = 100): fp64 tensor-core kernel vs the
register-tiled FMA kernel (NSB_ROTATE_DMMA=0).  python profiles/run_rotate.py"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def run(dmma, n=50331648, k=100, reps=3):
    os.environ['NSB_ROTATE_DMMA'] = '1' if dmma else '0'
    import nekstab_next_b200 as nb
    ctx = nb.Context(0)
    lay = nb.Layout(ctx, [n], [True])
    lay.set_weight([np.full(n, 1.0 / n)])
    Q = nb.Basis(lay, k)
    base = np.random.default_rng(0).standard_normal(1 << 20)
    for c in range(k):
        Q[c].upload([np.resize(np.roll(base, 31 * c), n)])
    Z, _ = np.linalg.qr(np.random.default_rng(1).standard_normal((k, k)))
    Q.rotate(k, Z)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        Q.rotate(k, Z)
    ms = ctx.timer_stop() / reps
    print(f'rotate {"DMMA" if dmma else "FMA "} n={n} k={k}: {ms:.2f} ms  {2.0 * n * k * k / ms / 1e9:.1f} TFLOP/s fp64  '
          f'{16.0 * n * k / ms / 1e6:.0f} GB/s algorithmic', flush=True)
    Q.close(); lay.close(); ctx.close()


if __name__ == '__main__':
    run(True)
    run(False)
