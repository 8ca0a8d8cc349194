"""Multi-rank parity check, launched by torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multirank_check.py

Every rank owns a z-slab of the box mesh (element partition like Nek's MPI ranks); inner products
go through the NCCL all-reduce and dssum through the interface exchange.  Results are compared with
the single-rank numpy oracle evaluated on the whole mesh.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import nekstab_next_b200 as nb
    from helpers import BoxProblem
    from oracle import krylov as okr, sem as osem
    box = [nb.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = nb.Context(device=local, rank=rank, nranks=world, unique_id=box[0])
    if os.environ.get('NSB_TEST_P2P', '0') == '1':
        def allgather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        assert ctx.connect_peers(allgather, halo_bytes=8 << 20) and ctx.p2p_enabled()

    nel, N, nc, K = (3, 2, 2 * world), 5, 2, 12
    P = BoxProblem(nel=nel, N=N, deform=0.04, nfields=nc, conv=True, time_in_dot=True, seed=21)  # whole mesh
    c = P.octx()
    per = nel[0] * nel[1]
    e0, e1 = nb.mesh.partition_range(P.shape[0], rank, world, granule=per)
    sl = slice(e0, e1)
    x, y, z = (a[sl] for a in P.coords)
    sem = nb.Sem(ctx, N, x, y, z, mask=P.mask[sl], glo_num=P.glo[sl])
    sem.setup_exchange()
    npts = sem.npts
    errs = {}

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    errs['binvm1'] = rel(sem.get('binvm1'), P.binv[sl])
    errs['vmult'] = rel(sem.get('vmult'), P.vmult[sl])
    lay = nb.Layout(ctx, [npts] * nc, [True] * nc, time_in_dot=True)
    lay.set_weight([P.bm1[sl]] * nc)
    Q = nb.Basis(lay, K + 1)
    conv = tuple(a[sl] for a in P.conv)
    op = nb.sem_operator(sem, nc, P.alpha, P.beta, P.h1, P.h2, conv=conv)
    # dssum / ax on a discontinuous field
    u = P.rng.standard_normal(P.shape)
    Q[0].upload([u[sl], u[sl]])
    sem.dssum(Q[0], 0)
    errs['dssum'] = rel(Q[0].download()[0][0], osem.dssum(u, P.glo)[sl].ravel())
    Q[0].upload([u[sl], u[sl]])
    sem.ax(Q[0], Q[1], 1, 1.0, 0.1)
    errs['ax'] = rel(Q[1].download()[0][1], osem.ax(u, P.geo['g'], P.d, P.glo, P.mask, 1.0, 0.1, P.bm1)[sl].ravel())
    # weighted dot with %time (counted once globally)
    a, b = P.random_kvec(), P.random_kvec()
    a.time, b.time = 1.5, -2.0
    Q[0].upload([f[sl] for f in a.f], a.time)
    Q[1].upload([f[sl] for f in b.f], b.time)
    ref = okr.k_dot(c, a, b)
    errs['dot'] = abs(Q[0].dot(Q[1]) - ref) / abs(ref)
    # Arnoldi: H identical on every rank and equal to the oracle's
    q0 = P.random_kvec()
    okr.k_normalize(c, q0)
    Qo = [okr.k_zero_like(q0) for _ in range(K + 1)]
    okr.k_copy(Qo[0], q0)
    Ho = np.zeros((K + 1, K))
    okr.arnoldi_factorization(c, P.omatvec, Qo, Ho, 1, K, K)
    Q[0].upload([f[sl] for f in q0.f], q0.time)
    H = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(Q, H, 1, K, K, op)
    errs['arnoldi_H'] = rel(H, Ho)
    errs['arnoldi_Q'] = rel(Q[K].download()[0][0], Qo[K].f[0][sl].ravel())
    G = Q.gram(K + 1)
    errs['orth'] = float(np.max(np.abs(G - np.eye(K + 1))))
    Hall = [None] * world
    dist.all_gather_object(Hall, H)
    errs['H_replicated'] = max(float(np.max(np.abs(h - Hall[0]))) for h in Hall)
    # replay of the captured steps (peer-memory transport: the whole step is a CUDA graph) gives the same bits
    Q[0].upload([f[sl] for f in q0.f], q0.time)
    Hr = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(Q, Hr, 1, K, K, op)
    errs['H_replay'] = float(np.max(np.abs(Hr - H)))
    # DGKS: the re-orthogonalisation decision is taken on every rank's device from the all-reduced norms --
    # all ranks must take the same branch (otherwise the next collective hangs or H diverges)
    Q[0].upload([f[sl] for f in q0.f], q0.time)
    Hd = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(Q, Hd, 1, K, K, op, nb.ORTH_DGKS)
    errs['dgks_H'] = rel(Hd, Ho)
    Gd = Q.gram(K + 1)
    errs['dgks_orth'] = float(np.max(np.abs(Gd - np.eye(K + 1))))
    pall = [None] * world
    dist.all_gather_object(pall, (Hd, nb.arnoldi_passes(Q, 1, K, nb.ORTH_DGKS)))
    errs['dgks_replicated'] = max(float(np.max(np.abs(h - pall[0][0]))) + float(np.max(np.abs(p - pall[0][1])))
                                  for h, p in pall)
    # a second mesh on the same context (velocity + pressure mesh in Nek) with a DIFFERENT neighbour set:
    # slabs handed out in a permuted order, lower order; both meshes exchange alternately
    perm = [0, 2, 1, 3][:world] if world == 4 else list(range(world))[::-1]
    P2 = BoxProblem(nel=nel, N=3, deform=0.02, nfields=1, beta=-1e-3, seed=5)
    s0, s1 = nb.mesh.partition_range(P2.shape[0], perm[rank], world, granule=per)
    sl2 = slice(s0, s1)
    sem2 = nb.Sem(ctx, 3, *(a[sl2] for a in P2.coords), mask=P2.mask[sl2], glo_num=P2.glo[sl2])
    sem2.setup_exchange()
    lay2 = nb.Layout(ctx, [sem2.npts], [True])
    B2 = nb.Basis(lay2, 1)
    errs['two_meshes'] = 0.0
    for it in range(3):
        ua, ub = P.rng.standard_normal(P.shape), P2.rng.standard_normal(P2.shape)
        Q[0].upload([ua[sl], ua[sl]])
        B2[0].upload([ub[sl2]])
        sem.dssum(Q[0], 0)
        sem2.dssum(B2[0], 0)
        if it == 1:
            sem2.dssum(B2[0], 0)          # uneven call counts: the sequence numbers are per mesh
            ub = osem.dssum(ub, P2.glo)
        errs['two_meshes'] = max(errs['two_meshes'], rel(Q[0].download()[0][0], osem.dssum(ua, P.glo)[sl].ravel()),
                                 rel(B2[0].download()[0][0], osem.dssum(ub, P2.glo)[sl2].ravel()))
    B2.close()
    sem2.close()
    # Helmholtz solves side by side: interface exchange inside every iteration, CG scalars all-reduced
    rhs = [osem.dssum(P.bm1 * P.random_field(), P.glo) * P.mask for _ in range(nc)]
    Q[0].upload([r[sl] for r in rhs])
    its, ress = sem.hmholtz_vec(Q[0], Q[1], 0, nc, 0.3, 5.0, tol=1e-11, maxit=400)
    xs = Q[1].download()[0]
    errs['hmholtz'] = 0.0
    for f in range(nc):
        xo, ito, _ = osem.cggo(rhs[f], P.geo['g'], P.d, P.glo, P.mask, P.bm1, 0.3, 5.0, tol=1e-11, maxit=400)
        errs['hmholtz'] = max(errs['hmholtz'], rel(xs[f], xo[sl].ravel()), 1.0 if abs(its[f] - ito) > 1 else 0.0)
    # the device time-stepper operator (dealiased convection is element-local; dssum + Helmholtz exchange)
    from test_gpu_conv import _oracle_scalar_steps
    vel = [np.sin(np.pi * P.coords[0]), -np.cos(np.pi * P.coords[1]), 0.3 + 0 * P.coords[2]]
    lay3 = nb.Layout(ctx, [npts] * 3, [True] * 3)
    Bv = nb.Basis(lay3, 1)
    Bv[0].upload([v[sl] for v in vel])
    sem.dealias_setup()
    sem.set_convect(0, Bv[0])
    step = nb.stepper_operator(sem, lay, 1, 0, 0.05, 4e-3, 3, tol=1e-13)
    T0 = P.random_field()
    Q[0].upload([T0[sl], T0[sl]], 0.0)
    step.matvec(Q[0], Q[1])
    Tref = _oracle_scalar_steps(*P.coords, P.glo, P.mask, P.geo, vel, T0, 0.05, 4e-3, 3, N=N)
    errs['stepper'] = rel(Q[1].download()[0][0], Tref[sl].ravel())
    step.close()
    # the pressure-coupled step: D^T p -> interface exchange -> D inside every pressure iteration, CG scalars all-reduced
    from oracle import ns as ons
    ps = ons.pressure_setup(N, P.geo)
    dl = osem.dealias_setup(N, 3 * (N + 1) // 2, P.geo['rst'])
    n2loc = sem.pressure_setup()
    n2e = ps['bm2'][0].size
    assert n2loc == (e1 - e0) * n2e
    layn = nb.Layout(ctx, [npts] * 3 + [n2loc], [True] * 3 + [False])
    Bn = nb.Basis(layn, 3)
    v0 = [P.random_field() for _ in range(3)]
    p0 = P.rng.standard_normal(ps['bm2'].shape)
    vo, po = ons.ns_steps(P.glo, P.mask, P.geo, N, ps, dl, vel, v0, p0, 0.05, 4e-3, 3, mean_free=False)
    Bn[2].upload([v[sl] for v in vel] + [0 * p0[sl]])
    nsop = nb.ns_stepper_operator(sem, layn, Bn[2], 0.05, 4e-3, 3, tol_v=1e-13, tol_p=1e-13, mean_free=False)
    Bn[0].upload([v[sl] for v in v0] + [p0[sl]])
    nsop.matvec(Bn[0], Bn[1])
    fo = Bn[1].download()[0]
    scale = max(np.max(np.abs(a)) for a in vo)
    errs['ns_stepper'] = max(float(np.max(np.abs(fo[b] - vo[b][sl].ravel()))) / scale for b in range(3))
    pm = po - po.mean()
    pall_ = [None] * world
    dist.all_gather_object(pall_, fo[3])
    pg = np.concatenate(pall_)
    errs['ns_pressure'] = float(np.max(np.abs((pg - pg.mean()) - pm.ravel())) / np.max(np.abs(pm)))
    nsop.close()
    Bn.close()
    Bv.close()
    tol = dict(ns_stepper=1e-9, ns_pressure=1e-6, binvm1=1e-12, vmult=0, dssum=1e-13, ax=1e-12, dot=1e-12, arnoldi_H=1e-10, arnoldi_Q=1e-9,
               orth=1e-10, H_replicated=0, hmholtz=1e-8, stepper=1e-9, H_replay=0, dgks_H=1e-10, dgks_orth=1e-10,
               dgks_replicated=0, two_meshes=1e-13)
    bad = {k: v for k, v in errs.items() if not (v <= tol[k])}
    print(f'[rank {rank}/{world}] ' + ' '.join(f'{k}={v:.2e}' for k, v in errs.items()), flush=True)
    op.close()
    ctx.close()
    dist.destroy_process_group()
    if bad:
        print(f'[rank {rank}] FAILED: {bad}', flush=True)
        sys.exit(1)


if __name__ == '__main__':
    main()
