"""Per-kernel bandwidth of the orthogonalisation kernels at the benchmark size for several k.
Variants are selected with NSB_FUSED_LOADER / NSB_FUSED_RC / NSB_NO_FUSED in the environment."""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nekstab_next_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--ks', default='10,25,50,75,100')
ap.add_argument('--n', type=int, default=50331648)
ap.add_argument('--reps', type=int, default=5)
a = ap.parse_args()
ks = [int(x) for x in a.ks.split(',')]
ctx = nb.Context(0)
lay = nb.Layout(ctx, [a.n], [True])
lay.set_weight([np.full(a.n, 1.0 / a.n)])
Q = nb.Basis(lay, max(ks) + 1)
rng = np.random.default_rng(0)
base = rng.standard_normal(1 << 20)
for c in range(max(ks) + 1):
    Q[c].upload([np.resize(np.roll(base, 17 * c), a.n)])
tag = ' '.join(f'{k}={os.environ[k]}' for k in ('NSB_FUSED_LOADER', 'NSB_FUSED_RC', 'NSB_NO_FUSED', 'NSB_FUSED_ALLWARPS', 'NSB_TAIL') if k in os.environ)
for k in ks:
    for _ in range(2):
        nb.orthonormalize(Q, k, k, nb.ORTH_CGS2)
    ctx.prof_enable(True)
    for _ in range(a.reps):
        nb.orthonormalize(Q, k, k, nb.ORTH_CGS2)
    rep = ctx.prof_report()
    ctx.prof_enable(False)
    row = [f'k={k:4d}']
    tot = 0.0
    for name in ('multidot', 'fused_update_dot', 'update', 'normalize'):
        if name in rep:
            v = rep[name]
            tot += v['ms'] / a.reps
            row.append(f"{name}: {v['ms'] / v['launches']:.3f} ms {v['bytes'] / v['ms'] / 1e6:7.0f} GB/s")
    print(f'[{tag}]', ' | '.join(row), f'| orth total {tot:.3f} ms', flush=True)
ctx.close()
