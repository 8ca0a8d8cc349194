"""The host part of the multi-rank gather-scatter (nsb_host_gs_plan + nsb_host_exchange_plan, the code behind
nsb_sem_create / nsb_sem_setup_exchange) on ARBITRARY element partitions, emulated in one process: every
"rank" owns a random subset of the elements, so nodes are shared by up to eight ranks and every rank has many
neighbours (the slab partitions of the gloo and GPU tests only ever share a node between two ranks).
Pairwise exchange of the local node sums + add must reproduce the global dssum."""
import ctypes as C

import numpy as np
import pytest

from oracle import sem as osem

I64P, I32P = C.POINTER(C.c_int64), C.POINTER(C.c_int32)


def plans(lib, check, dim, N, glo_parts):
    P = len(glo_parts)
    gs = []
    for glo in glo_parts:
        g = np.ascontiguousarray(glo, dtype=np.int64)
        nel = g.shape[0]
        nn, nnz = C.c_int64(), C.c_int64()
        check(lib.nsb_host_gs_plan(dim, N, nel, g.ctypes.data_as(I64P), C.byref(nn), C.byref(nnz), None, None, None))
        off, idx = np.zeros(nn.value + 1, np.int64), np.zeros(max(nnz.value, 1), np.int32)
        gid = np.zeros(max(nn.value, 1), np.int64)
        check(lib.nsb_host_gs_plan(dim, N, nel, g.ctypes.data_as(I64P), C.byref(nn), C.byref(nnz),
                                   off.ctypes.data_as(I64P), idx.ctypes.data_as(I32P), gid.ctypes.data_as(I64P)))
        gs.append((nn.value, off, idx[:nnz.value], gid[:nn.value]))
    cnt = np.array([g[0] for g in gs], np.int64)
    mx = int(cnt.max())
    all_sorted = np.full((P, mx), -1, np.int64)
    for r, g in enumerate(gs):
        all_sorted[r, :g[0]] = np.sort(g[3])
    ex = []
    for r, (nn, off, idx, gid) in enumerate(gs):
        newpos, nloc = np.zeros(nn, np.int64), C.c_int64()
        pcount, pnodes = np.zeros(P, np.int64), np.zeros((P, mx), np.int32)
        check(lib.nsb_host_exchange_plan(r, P, nn, gid.ctypes.data_as(I64P), cnt.ctypes.data_as(I64P),
                                         all_sorted.ctypes.data_as(I64P), mx, newpos.ctypes.data_as(I64P),
                                         C.byref(nloc), pcount.ctypes.data_as(I64P), pnodes.ctypes.data_as(I32P)))
        assert sorted(newpos.tolist()) == list(range(nn))
        ex.append((newpos, nloc.value, pcount, pnodes))
    return gs, ex


@pytest.mark.parametrize('dim,nel,N,P,seed', [(3, (3, 3, 2), 3, 4, 0), (3, (2, 2, 3), 4, 5, 1), (2, (5, 4), 5, 3, 2),
                                                (3, (4, 2, 2), 2, 8, 3)])
def test_random_partitions_reproduce_the_global_dssum(lib, dim, nel, N, P, seed):
    from nekstab_next_b200 import _capi
    rng = np.random.default_rng(seed)
    if dim == 3:
        x, y, z, glo = osem.box_mesh(*nel, N, deform=0.03)
    else:
        x, y, glo = osem.box_mesh_2d(*nel, N, deform=0.03)
    E = glo.shape[0]
    owner = rng.integers(0, P, E)
    owner[:P] = np.arange(P)                                   # nobody is empty
    elems = [np.where(owner == r)[0] for r in range(P)]
    gs, ex = plans(_capi.load(), _capi.check, dim, N, [glo[e] for e in elems])
    u = rng.standard_normal(glo.shape)
    ref = osem.dssum(u, glo)
    # local sums, in the exchange plan's order (private nodes first, interface nodes last)
    sums, ifc = [], []
    for r in range(P):
        nn, off, idx, gid = gs[r]
        newpos, nloc, pcount, pnodes = ex[r]
        ul = u[elems[r]].ravel()
        s = np.zeros(nn)
        s[newpos] = [ul[idx[off[n]:off[n + 1]]].sum() for n in range(nn)]
        sums.append(s)
        ifc.append(s[nloc:].copy())
    # both sides of a pair list their shared nodes in ascending global id: the packed buffers line up
    for r in range(P):
        newpos, nloc, pcount, pnodes = ex[r]
        for q in range(P):
            assert pcount[q] == ex[q][2][r]
            if q == r or pcount[q] == 0:
                continue
            sums[r][nloc + pnodes[q, :pcount[q]]] += ifc[q][ex[q][3][r, :pcount[q]]]
    shared_by_many = 0
    for r in range(P):
        nn, off, idx, gid = gs[r]
        newpos, nloc, pcount, pnodes = ex[r]
        out = u[elems[r]].ravel().copy()
        for n in range(nn):
            out[idx[off[n]:off[n + 1]]] = sums[r][newpos[n]]
        assert np.max(np.abs(out - ref[elems[r]].ravel())) < 1e-12
        hits = np.zeros(nn - nloc, int)
        for q in range(P):
            hits[pnodes[q, :pcount[q]]] += 1
        shared_by_many += int(np.count_nonzero(hits >= 2))
        assert np.all(hits >= 1)                               # every interface node has at least one peer
    assert shared_by_many > 0                                  # the partition really has nodes on 3+ ranks
