"""The reference's .re2 mesh files through nekstab_next_b200.mesh.read_re2 / dirichlet_mask, pinned by the reference's
own data: the files parse to the last byte, the element corners are the corner nodes of the base-flow field files,
and the boundary conditions agree with what those base flows do on the faces (walls at rest, the inflow profile,
periodic partners).  Host-only; the reference-gated part runs where /root/reference exists, the rest on the committed
tables tests/golden/*_bc.npz (written by tests/golden/make_golden.py)."""
import collections
from pathlib import Path

import numpy as np
import pytest

from nekstab_next_b200 import mesh
from oracle import nekfld

REF = Path('/root/reference/examples')
GOLD = Path(__file__).parent / 'golden'
needs_ref = pytest.mark.skipif(not REF.exists(), reason='the reference tree is not present')
CASES = {'cyl': ('cylinder/1cyl.re2', 'cylinder/BF_1cyl0.f00001'),
         'bfs': ('back_fstep/baseflow/bfs.re2', 'back_fstep/baseflow/BF_bfs0.f00001')}


def corner_nodes(a):
    """(nel, ly, lx) -> the four corner values in preprocessor order (counter-clockwise from (-1, -1))."""
    return np.stack([a[:, 0, 0], a[:, 0, -1], a[:, -1, -1], a[:, -1, 0]], axis=1)


@needs_ref
@pytest.mark.parametrize('name', ['cyl', 'bfs'])
def test_re2_matches_the_field_file_of_the_same_case(name):
    re2 = mesh.read_re2(REF / CASES[name][0])
    f = nekfld.read_fld(REF / CASES[name][1])
    assert re2['ndim'] == 2 and re2['nel'] == re2['nelv'] == f['nel'] == f['nelg']
    x, y = f['x']
    order = f['elmap'] - 1                                       # file order -> global element ids
    # the .re2 vertices were written from single-precision mesh data and Nek's geometry fix-up (vertex averaging)
    # moves shared nodes by a few 1e-7: the corners agree to that, far below the smallest element (0.05)
    assert np.max(np.abs(corner_nodes(x) - re2['xc'][order])) <= 5e-6
    assert np.max(np.abs(corner_nodes(y) - re2['yc'][order])) <= 5e-6
    assert len(re2['bcs']) == 1                                  # velocity only (no temperature in these cases)
    for e, side, params, typ in re2['bcs'][0] + re2['curves']:
        assert 1 <= e <= re2['nel'] and 1 <= side <= 4
    tags = collections.Counter(t for _, _, _, t in re2['bcs'][0])
    if name == 'cyl':
        assert re2['version'] == '#v002'
        assert tags == {'P  ': 132, 'v  ': 30, 'O  ': 30, 'W  ': 16}
        assert len(re2['curves']) == 80 and {c[3][0] for c in re2['curves']} == {'C'}
        # every curved side is an arc about the origin: the GLL points the field file holds on that side lie on the
        # circle of the recorded radius (the cylinder, R = 0.5, and the rings of its O-grid)
        loc = {int(g): l for l, g in enumerate(f['elmap'])}
        radii = set()
        for e, side, params, _ in re2['curves']:
            idx = (loc[e],) + mesh.face_nodes(f['nx'], 2, side)
            assert np.max(np.abs(np.hypot(x[idx], y[idx]) - abs(params[0]))) <= 5e-6
            radii.add(round(abs(params[0]), 6))
        assert min(radii) == 0.5
    else:
        assert re2['version'] == '#v003' and tags == {'MSH': 236} and not re2['curves']
        assert collections.Counter(int(p[4]) for _, _, p, _ in re2['bcs'][0]) == {3: 206, 2: 20, 4: 10}


@needs_ref
def test_boundary_conditions_agree_with_the_base_flows():
    """Cylinder: the base flow is at rest on the 'W' faces, equal to (1, 0) on the 'v' faces, and takes the same
    values on a 'P' face and on its partner (parameters 1, 2 = partner element and side).  Step: boundary id 3 are
    the walls (at rest), 4 the inflow (parabolic profile, no cross-flow), 2 the outflow."""
    for name in ('cyl', 'bfs'):
        re2 = mesh.read_re2(REF / CASES[name][0])
        f = nekfld.read_fld(REF / CASES[name][1])
        lx = f['nx']
        loc = {int(g): l for l, g in enumerate(f['elmap'])}
        (x, y), (u, v) = f['x'], f['u']
        for e, side, params, typ in re2['bcs'][0]:
            idx = (loc[e],) + mesh.face_nodes(lx, 2, side)
            wall = typ == 'W  ' or (typ == 'MSH' and int(params[4]) == 3)
            if wall:
                assert np.max(np.abs(u[idx])) <= 1e-12 and np.max(np.abs(v[idx])) <= 1e-12
            elif typ == 'v  ':
                assert np.max(np.abs(u[idx] - 1.0)) <= 1e-12 and np.max(np.abs(v[idx])) <= 1e-12
                assert np.max(np.abs(x[idx] + 16.0)) <= 1e-12
            elif typ == 'O  ':
                assert np.max(np.abs(x[idx] - 50.0)) <= 1e-12
            elif typ == 'P  ':
                pidx = (loc[int(params[0])],) + mesh.face_nodes(lx, 2, int(params[1]))
                assert np.max(np.abs(np.abs(y[idx]) - 16.0)) <= 1e-12 and np.max(np.abs(y[idx] + y[pidx])) <= 1e-12
                assert np.max(np.abs(x[idx] - x[pidx])) <= 1e-12
                assert np.max(np.abs(u[idx] - u[pidx])) <= 1e-10 and np.max(np.abs(v[idx] - v[pidx])) <= 1e-10
            elif typ == 'MSH' and int(params[4]) == 4:
                assert np.max(np.abs(x[idx] + 10.0)) <= 1e-12 and np.max(np.abs(v[idx])) <= 1e-12
                assert np.max(np.abs(u[idx] - 4.0 * y[idx] * (1.0 - y[idx]))) <= 1e-6      # single-precision mesh data
            elif typ == 'MSH' and int(params[4]) == 2:
                assert np.max(np.abs(x[idx] - 50.0)) <= 1e-12
            else:
                raise AssertionError((name, typ, params))


@pytest.mark.parametrize('name', ['cyl', 'bfs'])
def test_dirichlet_mask_from_the_committed_boundary_tables(name):
    """dirichlet_mask on the committed tables: zero exactly on the wall / prescribed-velocity faces (where the
    committed base flow is at rest or at its inflow value), one elsewhere; local element lists select and reorder."""
    g = np.load(GOLD / f'{name}_mesh.npz')
    t = np.load(GOLD / f'{name}_bc.npz')
    re2 = dict(ndim=2, nel=int(t['nel']),
               bcs=[[(int(e), int(s), p, str(ty)) for e, s, p, ty in zip(t['elem'], t['side'], t['params'], t['type'])]])
    lx = g['x'].shape[-1]
    kw = dict(types=('W  ', 'v  ')) if name == 'cyl' else dict(types=(), ids=(3, 4))
    order = t['elmap']
    faces_only = mesh.dirichlet_mask(re2, lx, elements=order, **kw)
    mask = mesh.dirichlet_mask(re2, lx, elements=order, glo=g['glo'], **kw)
    assert mask.shape == g['x'].shape and set(np.unique(mask)) == {0.0, 1.0}
    assert np.all(mask <= faces_only)
    u, v, x, y = g['u'], g['v'], g['x'], g['y']
    dirichlet = mask == 0.0
    if name == 'cyl':
        r = np.hypot(x, y)
        want = (np.abs(r - 0.5) <= 1e-5) | (np.abs(x + 16.0) <= 1e-12)      # the cylinder and the inflow plane
        assert np.array_equal(dirichlet, want)
        assert np.max(np.hypot(u, v)[np.abs(r - 0.5) <= 1e-5]) <= 1e-12
        assert np.array_equal(mask, faces_only)                  # an O-grid: no element touches the wall with a corner only
    else:
        assert (mask < faces_only).sum() > 0                     # the step corner is such a node
        inflow = np.abs(x + 10.0) <= 1e-12
        outflow = np.abs(x - 50.0) <= 1e-12
        assert np.array_equal(dirichlet & ~inflow, (np.hypot(u, v) <= 1e-12) & ~inflow & dirichlet)
        assert dirichlet[inflow].all() and not dirichlet[outflow & (np.abs(np.abs(y) - 1.0) > 1e-12)].any()
        interior = (np.abs(np.abs(y) - 1.0) > 1e-12) & (x > 1e-9) & ~outflow
        assert not dirichlet[interior].any()
    sub = order[::7]                                             # a rank's share of the elements, in its own order
    assert np.array_equal(mesh.dirichlet_mask(re2, lx, elements=sub, **kw), faces_only[::7])
    with pytest.raises(ValueError):
        mesh.face_nodes(lx, 2, 5)


def test_face_nodes_numbering():
    lx = 4
    a = np.arange(lx * lx).reshape(lx, lx)                       # (j, i)
    assert list(a[mesh.face_nodes(lx, 2, 1)]) == [0, 1, 2, 3]                # s = -1
    assert list(a[mesh.face_nodes(lx, 2, 2)]) == [3, 7, 11, 15]              # r = +1
    assert list(a[mesh.face_nodes(lx, 2, 3)]) == [12, 13, 14, 15]            # s = +1
    assert list(a[mesh.face_nodes(lx, 2, 4)]) == [0, 4, 8, 12]               # r = -1
    b = np.arange(lx ** 3).reshape(lx, lx, lx)                   # (k, j, i)
    assert b[mesh.face_nodes(lx, 3, 5)].tolist() == b[0].tolist() and b[mesh.face_nodes(lx, 3, 6)].tolist() == b[-1].tolist()
    assert b[mesh.face_nodes(lx, 3, 2)].tolist() == b[:, :, -1].tolist()


BASEFLOW = {'cyl': dict(nu=1.0 / 50.0, kw=dict(types=('W  ', 'v  ')), tol=3e-5, outflow=('O  ',)),
            'bfs': dict(nu=1.0 / 500.0, kw=dict(types=(), ids=(3, 4)), tol=3e-4, outflow=(2,))}


@pytest.mark.parametrize('name', ['cyl', 'bfs'])
def test_base_flow_balances_momentum_up_to_the_outflow_and_periodic_boundaries(name):
    """The reference's base flows with the boundary conditions of their .re2 files: the discrete steady momentum
    balance QQ^T [B C(U) U + nu A U - D^T p] = 0 (oracle restatements of Nek's operators, as in
    tests/test_oracle_fixtures.py) holds at EVERY node the velocity mask leaves free -- on the outflow boundary, where
    'O' is the natural condition of that weak form and nothing is masked, and on the periodic boundary once the
    numbering identifies the partners -- as well as in the interior.  This pins the reader, the mask, the periodic
    numbering and the treatment of outflow by leaving the boundary unmasked against data computed by Nek5000 itself.
    Control: without the periodic identification the balance fails on the periodic faces."""
    import json
    from oracle import ns as ons, sem
    cfg = BASEFLOW[name]
    g = np.load(GOLD / f'{name}_mesh.npz')
    t = np.load(GOLD / f'{name}_bc.npz')
    N = json.loads((GOLD / 'known_answers.json').read_text())[name]['N']
    lx = N + 1
    x, y, u, v, pm1 = g['x'], g['y'], g['u'], g['v'], g['p']
    re2 = dict(ndim=2, nel=int(t['nel']),
               bcs=[[(int(e), int(s), p, str(ty)) for e, s, p, ty in zip(t['elem'], t['side'], t['params'], t['type'])]])
    order = t['elmap']
    loc = {int(gid): l for l, gid in enumerate(order)}
    glo = g['glo'].astype(np.int64)
    glop = mesh.periodic_glo_num(glo, re2, (x, y), lx, elements=order)
    nper = sum(1 for b in re2['bcs'][0] if b[3] == 'P  ')
    if name == 'cyl':
        assert nper == 132 and glop.max() + 1 == glo.max() + 1 - (66 * N + 1)      # one line of 66 elements folded
    else:
        assert nper == 0 and np.array_equal(glop, glo)

    def faces(kinds):
        m = np.zeros(x.shape, bool)
        for e, s, p, ty in re2['bcs'][0]:
            if ty in kinds or (ty == 'MSH' and int(p[4]) in kinds):
                m[(loc[e],) + mesh.face_nodes(lx, 2, s)] = True
        return m

    outflow, periodic = faces(cfg['outflow']), faces(('P  ',))
    assert outflow.sum() > 0
    geo = sem.geometry(N, x, y)
    ps = ons.pressure_setup(N, geo)
    dl = sem.dealias_setup(N, 9, geo['rst'])
    cf = sem.set_convect([u, v], dl)
    gt = ons.opgradt(sem.interp_fine(pm1, ps['I12']), ps)
    d = sem.dgll(N)

    def residual(numbering):
        mask = mesh.dirichlet_mask(re2, lx, elements=order, glo=numbering, **cfg['kw'])
        assert mask[outflow & ~faces(('W  ', 'v  ', 3, 4))].all()                    # the outflow boundary is left free
        worst = dict(all=0.0, outflow=0.0, periodic=0.0)
        scale = 0.0
        for b, a in enumerate((u, v)):
            conv = sem.convect_dealiased(a, cf, dl)
            visc = sem.axhelm(a, geo['g'], d, cfg['nu'], 0.0, geo['bm1'])
            r = np.abs(sem.dssum(conv + visc - gt[b], numbering) * mask)
            scale = max(scale, *(float(np.max(np.abs(sem.dssum(term, numbering) * mask))) for term in (conv, visc, gt[b])))
            worst['all'] = max(worst['all'], float(r.max()))
            worst['outflow'] = max(worst['outflow'], float(r[outflow].max()))
            if periodic.any():
                worst['periodic'] = max(worst['periodic'], float(r[periodic].max()))
        return {k: w / scale for k, w in worst.items()}

    res = residual(glop)
    assert res['all'] <= cfg['tol'] and res['outflow'] <= cfg['tol'] and res['periodic'] <= cfg['tol'], res
    if name == 'cyl':
        assert residual(glo)['periodic'] >= 1000 * max(res['periodic'], 1e-9)
