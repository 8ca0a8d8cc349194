#!/usr/bin/env python
"""Microbenchmark sweep of BASELINE.json configs[4]: weighted Gram-Schmidt (CGS2 step) and the SEM
matvec over k in {20,50,100,200,400} x n in {1,4,16,64,200} M dof, with the memory-feasibility
mask of SURVEY.md section 8d (8 n (k+3) / P <= 0.85 HBM per GPU).

  python bench_sweep.py [--gpus N] [--out profiles/sweep_rNN.json]

For N > 1 launch with torch.distributed.run like bench.py; n is the GLOBAL dof count (rows are
partitioned evenly), so the sweep is a strong-scaling one.  Orthogonalisation cost does not depend
on the mesh, so the vectors are plain rows; the matvec line uses a box mesh of the nearest size.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

KS = (20, 50, 100, 200, 400)
NS = (1, 4, 16, 64, 200)          # M dof
HBM = 180e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--out', default='')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--ks', default=','.join(map(str, KS)))
    ap.add_argument('--ns', default=','.join(map(str, NS)))
    a = ap.parse_args()
    import torch
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    uid = None
    import nekstab_next_b200 as nb
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        box = [nb.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    ctx = nb.Context(device=local, rank=rank, nranks=world, unique_id=uid)
    ks = [int(x) for x in a.ks.split(',')]
    ns = [int(x) for x in a.ns.split(',')]
    rows = []
    rng = np.random.default_rng(rank)
    base = rng.standard_normal(1 << 20)
    for nM in ns:
        n = nM * 1_000_000 // world
        kfeas = [k for k in ks if 8.0 * n * (k + 3) <= 0.85 * HBM]
        if not kfeas:
            rows.append(dict(n_mdof=nM, k=None, feasible=False))
            continue
        lay = nb.Layout(ctx, [n], [True])
        lay.set_weight([np.full(n, 1.0 / (nM * 1e6))])
        Q = nb.Basis(lay, max(kfeas) + 1)
        for c in range(max(kfeas) + 1):
            Q[c].upload([np.resize(np.roll(base, 17 * c + 1), n)])
        for k in ks:
            if k not in kfeas:
                rows.append(dict(n_mdof=nM, k=k, feasible=False))
                continue
            for _ in range(2):
                nb.orthonormalize(Q, k, k, nb.ORTH_CGS2)
            ctx.sync()
            ctx.timer_start()
            for _ in range(a.reps):
                nb.orthonormalize(Q, k, k, nb.ORTH_CGS2)
            ms = ctx.timer_stop() / a.reps
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            # algorithmic bytes of one CGS2 step: three sweeps over V -- multidot 8 n (k+2), fused update+multidot+norm
            # 8 n (k+3), update with the folded normalisation 8 n (k+2); no separate normalisation pass
            nn = nM * 1e6
            bytes_alg = 8.0 * nn * (3 * k + 7)
            rows.append(dict(n_mdof=nM, k=k, feasible=True, cgs2_ms=round(ms, 4),
                             gbs_per_gpu=round(bytes_alg / world / (ms * 1e-3) / 1e9, 1),
                             gdof_per_s=round(nn / (ms * 1e-3) / 1e9, 3)))
            if rank == 0:
                print(rows[-1], file=sys.stderr, flush=True)
        Q.close()
        lay.close()
    if rank == 0:
        out = dict(metric='cgs2_step_ms', n_gpus=world, dtype='f64', rows=rows,
                   note='one CGS2 orthonormalisation of a vector against k columns (multidot, fused '
                        'update+multidot+norm, update with the folded normalisation; 2 all-reduces when n_gpus > 1; '
                        'k beyond the fused kernels: unfused pair + normalize); gbs_per_gpu = 8 n (3k+7) / n_gpus / time')
        s = json.dumps(out)
        if a.out:
            Path(a.out).write_text(s)
        print(s)
    ctx.close()


if __name__ == '__main__':
    main()
