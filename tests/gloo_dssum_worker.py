"""Worker for tests/test_multirank_cpu.py (gloo, CPU only).

Each rank owns a z-slab of a box mesh.  The gather-scatter plan and the exchange plan come from the
library's host-only entry points (the code the GPU path runs in nsb_sem_create /
nsb_sem_setup_exchange); the data movement the GPU does with kernels + NCCL is emulated with numpy +
gloo send/recv.  The result must equal the single-rank oracle dssum of the whole mesh, and a
partitioned BM1-weighted dot (local partial + all-reduce) must equal the global one.
"""
import ctypes as C
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
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo')
    import nekstab_next_b200 as nb
    from nekstab_next_b200 import _capi
    from oracle import sem as osem
    lib = _capi.load()
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    N, nel3 = 4, (3, 2, 2 * world + 1)        # odd layer count -> uneven slabs
    full = nb.mesh.box_mesh(*nel3, N, deform=0.03)
    mine = nb.mesh.box_mesh(*nel3, N, deform=0.03, rank=rank, nranks=world)
    e0, nel = mine['e0'], mine['nel']
    assert np.array_equal(mine['glo'], full['glo'][e0:e0 + nel]) and np.array_equal(mine['x'], full['x'][e0:e0 + nel])
    glo = np.ascontiguousarray(mine['glo'], dtype=np.int64)
    # --- gather-scatter plan (host part of nsb_sem_create)
    nn, nnz = C.c_int64(), C.c_int64()
    _capi.check(lib.nsb_host_gs_plan(3, N, nel, glo.ctypes.data_as(i64p), C.byref(nn), C.byref(nnz), None, None, None))
    off, idx, gid = np.zeros(nn.value + 1, np.int64), np.zeros(nnz.value, np.int32), np.zeros(nn.value, np.int64)
    _capi.check(lib.nsb_host_gs_plan(3, N, nel, glo.ctypes.data_as(i64p), C.byref(nn), C.byref(nnz),
                                     off.ctypes.data_as(i64p), idx.ctypes.data_as(i32p), gid.ctypes.data_as(i64p)))
    assert len(np.unique(gid)) == nn.value and np.array_equal(glo.ravel()[idx[off[:-1]]], gid)
    # --- exchange plan (host part of nsb_sem_setup_exchange): all-gather sorted ids with gloo
    counts = [None] * world
    dist.all_gather_object(counts, int(nn.value))
    mx = max(counts)
    srt = np.full(mx, -1, np.int64)
    srt[:nn.value] = np.sort(gid)
    allg = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allg, torch.from_numpy(srt))
    all_sorted = np.ascontiguousarray(np.stack([t.numpy() for t in allg]))
    cnt = np.array(counts, np.int64)
    newpos, nloc = np.zeros(nn.value, np.int64), C.c_int64()
    pcount, pnodes = np.zeros(world, np.int64), np.zeros((world, mx), np.int32)
    _capi.check(lib.nsb_host_exchange_plan(rank, world, nn.value, gid.ctypes.data_as(i64p), cnt.ctypes.data_as(i64p),
                                           all_sorted.ctypes.data_as(i64p), mx, newpos.ctypes.data_as(i64p),
                                           C.byref(nloc), pcount.ctypes.data_as(i64p), pnodes.ctypes.data_as(i32p)))
    nloc = nloc.value
    assert sorted(newpos.tolist()) == list(range(nn.value)) and pcount[rank] == 0
    inv = np.argsort(newpos)                          # new position -> old node
    # slab partition: only the neighbouring slabs share nodes, one plane of (3N+1)(2N+1) nodes each
    plane = (3 * N + 1) * (2 * N + 1)
    for r in range(world):
        assert pcount[r] == (plane if abs(r - rank) == 1 else 0), (rank, r, pcount[r])
    assert nn.value - nloc == plane * ((rank > 0) + (rank < world - 1))
    # --- emulate the device data path
    rng = np.random.default_rng(5)
    u_full = rng.standard_normal(full['x'].shape)
    u = u_full[e0:e0 + nel].copy().ravel()
    node_sum = np.array([u[idx[off[n]:off[n + 1]]].sum() for n in range(nn.value)])[inv]   # reordered
    ifc = node_sum[nloc:].copy()
    reqs, recv = [], {}
    for r in range(world):
        if pcount[r] == 0:
            continue
        sel = pnodes[r, :pcount[r]]
        recv[r] = torch.zeros(int(pcount[r]), dtype=torch.float64)
        reqs.append(dist.isend(torch.from_numpy(ifc[sel].copy()), dst=r))
        reqs.append(dist.irecv(recv[r], src=r))
    for q in reqs:
        q.wait()
    for r, buf in recv.items():
        node_sum[nloc + pnodes[r, :pcount[r]]] += buf.numpy()
    out = u.copy()
    for m in range(nn.value):
        n = inv[m]
        out[idx[off[n]:off[n + 1]]] = node_sum[m]
    ref = osem.dssum(u_full, full['glo'])[e0:e0 + nel].ravel()
    err = float(np.max(np.abs(out - ref)))
    # --- partitioned weighted dot = local partial + all-reduce
    geo = osem.geometry(N, full['x'], full['y'], full['z'])
    a, b = rng.standard_normal(u_full.shape), rng.standard_normal(u_full.shape)
    part = torch.tensor([osem.glsc3(a[e0:e0 + nel], b[e0:e0 + nel], geo['bm1'][e0:e0 + nel])], dtype=torch.float64)
    dist.all_reduce(part)
    dref = osem.glsc3(a, b, geo['bm1'])
    derr = abs(part.item() - dref) / abs(dref)
    print(f'[rank {rank}/{world}] nodes={nn.value} private={nloc} dssum err={err:.2e} dot err={derr:.2e}', flush=True)
    dist.destroy_process_group()
    if not (err < 1e-13 and derr < 1e-13):
        sys.exit(1)


if __name__ == '__main__':
    main()
