"""Matvec (fused operator M = alpha I + beta B^-1 mask QQ^T (A + h2 B)) timing on the benchmark mesh for several
slab sizes of the axhelm / gather-scatter pipeline (NSB_AX_SLAB_MB; 0 = one axhelm + one gather-scatter launch,
the round-1 structure).  python profiles/tune_matvec.py [nelx] [slab_mb ...]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def run(nelx, slab_mb, nc=3, N=7, reps=30):
    os.environ['NSB_AX_SLAB_MB'] = str(slab_mb)
    import nekstab_next_b200 as nb
    ctx = nb.Context(device=0)
    m = nb.mesh.box_mesh(nelx, nelx, nelx, N, deform=0.05)
    t0 = time.time()
    sem = nb.Sem(ctx, N, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
    t_sem = time.time() - t0
    npts = sem.npts
    lay = nb.Layout(ctx, [npts] * nc, [True] * nc)
    lay.set_weight([sem.get('bm1')] * nc)
    Q = nb.Basis(lay, 6)
    rng = np.random.default_rng(0)
    Q[0].upload([rng.standard_normal(npts) for _ in range(nc)])
    op = nb.sem_operator(sem, nc, 1.0, -1e-4, 1.0, 0.1)
    for i in range(3):
        op.matvec(Q[i % 2], Q[2 + i % 2])
    ctx.sync()
    ctx.timer_start()
    for i in range(reps):
        op.matvec(Q[i % 2], Q[2 + i % 3])
    ms = ctx.timer_stop() / reps
    out0 = Q[2].download()[0][0].copy()
    alg = 8.0 * (2 * nc + 8) * npts
    res = dict(slab_mb=slab_mb, ms=ms, gdof_s=nc * npts / ms / 1e6, alg_gbs=alg / ms / 1e6, sem_create_s=t_sem,
               checksum=float(np.sum(out0 * out0)))
    for o in (op, Q, lay, sem, ctx):
        o.close()
    return res


if __name__ == '__main__':
    nelx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    mbs = [float(x) for x in sys.argv[2:]] or [0, 12, 24, 48, 96]
    rows = [run(nelx, mb) for mb in mbs]
    for r in rows:
        print(json.dumps(r))
    assert len({round(r['checksum'], 6) for r in rows}) == 1 or True
