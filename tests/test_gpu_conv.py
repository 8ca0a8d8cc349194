"""Time-stepper pieces around ax against the oracle restatement of Nek5000's
convect.f / perturb.f routines ([UPSTREAM-RECALL], parity unpinned): dealiased convection on the
lxd Gauss-Legendre mesh and the fused EXT / BDF sums.  fp64, tolerance 1e-12 relative."""
import numpy as np
import pytest

from helpers import relerr
from oracle import sem as osem

pytestmark = pytest.mark.gpu


def _mesh(nel, N, deform):
    x, y, z, glo = osem.box_mesh(*nel, N, deform=deform)
    return x, y, z, glo, osem.geometry(N, x, y, z)


@pytest.mark.parametrize('nel,N,deform', [((2, 2, 2), 7, 0.05), ((3, 2, 1), 7, 0.0), ((2, 2, 2), 5, 0.04),
                                          ((2, 1, 2), 3, 0.05), ((2, 2, 1), 4, 0.03)])
def test_dealiased_convection_matches_oracle(ctx, nel, N, deform):
    import nekstab_next_b200 as nb
    x, y, z, glo, geo = _mesh(nel, N, deform)
    lxd = 8 if N == 4 else 3 * (N + 1) // 2
    dl = osem.dealias_setup(N, lxd, geo['rst'])
    rng = np.random.default_rng(N)
    vel = [np.sin(2 * x) * np.cos(y) + 0.3, 0.5 * np.cos(x + z), rng.standard_normal(x.shape)]
    u = [rng.standard_normal(x.shape) for _ in range(3)]
    cf = osem.set_convect(vel, dl)
    ref = [osem.convect_dealiased(a, cf, dl) for a in u]

    sem = nb.Sem(ctx, N, x, y, z, glo_num=glo)
    lay = nb.Layout(ctx, [x.size] * 3, [True] * 3)
    lay.set_weight([geo['bm1']] * 3)
    B = nb.Basis(lay, 3)
    sem.dealias_setup()
    B[0].upload(vel)
    sem.set_convect(0, B[0])
    B[1].upload(u)
    sem.convect(0, B[1], B[2], field0=0, nf=3)
    out, _ = B[2].download()
    for f in range(3):
        assert relerr(out[f].reshape(x.shape), ref[f]) <= 1e-12
    # second slot, a single field, scaled accumulation:  out_1 += -0.5 * conv(u_1 ; c = u)
    B[0].upload(u)
    sem.set_convect(1, B[0])
    cf2 = osem.set_convect(u, dl)
    sem.convect(1, B[1], B[2], field0=1, nf=1, scale=-0.5, accumulate=True)
    out2, _ = B[2].download()
    assert relerr(out2[1].reshape(x.shape), ref[1] - 0.5 * osem.convect_dealiased(u[1], cf2, dl)) <= 1e-12
    assert np.array_equal(out2[0], out[0]) and np.array_equal(out2[2], out[2])
    # slot 0 is untouched by the second set_convect
    sem.convect(0, B[1], B[2], field0=2, nf=1)
    out3, _ = B[2].download()
    assert relerr(out3[2].reshape(x.shape), ref[2]) <= 1e-12
    B.close()
    sem.close()


def test_dealiased_convection_is_exact_for_polynomials(ctx):
    """On an affine mesh the 3/2 rule integrates (c . grad u) v exactly for c, u, v of degree N:
    sum_p [J^T (c.grad)(J u)]_p = integral of c . grad u."""
    import nekstab_next_b200 as nb
    N = 7
    x, y, z, glo, geo = _mesh((2, 2, 2), N, 0.0)
    sem = nb.Sem(ctx, N, x, y, z, glo_num=glo)
    lay = nb.Layout(ctx, [x.size] * 3, [True] * 3)
    lay.set_weight([geo['bm1']] * 3)
    B = nb.Basis(lay, 3)
    sem.dealias_setup()
    B[0].upload([x ** 3, y * x, 1.0 + z ** 2])              # c
    sem.set_convect(0, B[0])
    B[1].upload([x ** 4 * y, z ** 3 + x, y ** 2 * z ** 2])  # u
    sem.convect(0, B[1], B[2], 0, 3)
    out, _ = B[2].download()
    # integrals over [0,1]^3 of c . grad u, by hand:
    #  u0 = x^4 y : c.grad = x^3 * 4 x^3 y + y x * x^4            -> 4/7 * 1/2 + 1/2 * 1/6
    #  u1 = z^3+x : c.grad = x^3 * 1 + (1 + z^2) * 3 z^2          -> 1/4 + 1 + 3/5
    #  u2 = y^2 z^2: c.grad = y x * 2 y z^2 + (1 + z^2) * 2 y^2 z -> 2 * 1/2 * 1/3 * 1/3 + 2 * 1/3 * (1/2 + 1/4)
    exact = [4 / 7 / 2 + 1 / 12, 1 / 4 + 1 + 3 / 5, 1 / 9 + 0.5]
    for f in range(3):
        assert abs(out[f].sum() - exact[f]) <= 1e-12
    B.close()
    sem.close()


def test_dealias_errors(ctx):
    import nekstab_next_b200 as nb
    x, y, z, glo, geo = _mesh((1, 1, 1), 7, 0.0)
    sem = nb.Sem(ctx, 7, x, y, z, glo_num=glo)
    lay = nb.Layout(ctx, [x.size] * 3, [True] * 3)
    B = nb.Basis(lay, 2)
    with pytest.raises(nb.NsbError):
        sem.set_convect(0, B[0])                 # no dealias_setup yet
    with pytest.raises(nb.NsbError):
        sem.dealias_setup(11)                    # no kernel for (8, 11)
    sem.dealias_setup()
    with pytest.raises(nb.NsbError):
        sem.convect(0, B[0], B[1])               # slot never set
    sem.set_convect(0, B[0])
    with pytest.raises(nb.NsbError):
        sem.convect(0, B[0], B[0])               # in place
    with pytest.raises(nb.NsbError):
        sem.convect(2, B[0], B[1])
    B.close()
    sem.close()


@pytest.mark.parametrize('nbd', [1, 2, 3])
def test_bdf_ext_matches_oracle(ctx, nbd):
    import nekstab_next_b200 as nb
    N = 7
    x, y, z, glo, geo = _mesh((2, 2, 1), N, 0.05)
    sem = nb.Sem(ctx, N, x, y, z, glo_num=glo)
    lay = nb.Layout(ctx, [x.size] * 3, [True] * 3)
    lay.set_weight([geo['bm1']] * 3)
    B = nb.Basis(lay, 6)
    rng = np.random.default_rng(nbd)
    host = [[rng.standard_normal(x.shape) for _ in range(3)] for _ in range(6)]   # bf e1 e2 v0 v1 v2
    for c in range(6):
        B[c].upload(host[c])
    ab = np.array([23.0 / 12, -16.0 / 12, 5.0 / 12])
    bd = np.array([11.0 / 6, -3.0, 1.5, -1.0 / 3])[:nbd + 1] if nbd == 3 else \
        (np.array([1.5, -2.0, 0.5]) if nbd == 2 else np.array([1.0, -1.0]))
    rho_dt = 1.0 / 2e-3
    sem.bdf_ext(B[0], B[1], B[2], [B[3 + i] for i in range(nbd)], ab, bd, rho_dt, field0=0, nf=3)
    for f in range(3):
        bf, e1, e2 = host[0][f].copy(), host[1][f].copy(), host[2][f].copy()
        osem.bdf_ext(bf, e1, e2, [host[3 + i][f] for i in range(nbd)], geo['bm1'], ab, bd, rho_dt)
        for col, ref in ((0, bf), (1, e1), (2, e2)):
            got = B[col].download()[0][f].reshape(x.shape)
            assert relerr(got, ref) <= 1e-13
    B.close()
    sem.close()


# ---- the pieces composed: Nek's scalar step (cdscal: convab -> makeabq -> makebdq -> hmholtz) ------------
BD = {1: [1.0, 1.0], 2: [1.5, 2.0, -0.5], 3: [11.0 / 6, 3.0, -1.5, 1.0 / 3]}
AB = {1: [1.0, 0.0, 0.0], 2: [2.0, -1.0, 0.0], 3: [3.0, -3.0, 1.0]}


def _device_scalar_steps(ctx, x, y, z, glo, mask, bm1, vel, T0, kappa, dt, nsteps, N=7):
    """BDF3/EXT3 advection-diffusion steps of one scalar through the C ABI; the lag rotation is the host's."""
    import nekstab_next_b200 as nb
    sem = nb.Sem(ctx, N, x, y, z, mask=mask, glo_num=glo)
    lay = nb.Layout(ctx, [x.size] * 3, [True] * 3)
    lay.set_weight([bm1] * 3)
    B = nb.Basis(lay, 8)         # 0 vel | 1..3 T lags (field 0) | 4 bq | 5, 6 ext history | 7 new T
    sem.dealias_setup()
    B[0].upload(vel)
    sem.set_convect(0, B[0])
    for c in range(1, 8):
        B[c].zero()
    B[1].upload([T0, 0 * T0, 0 * T0])
    lag, new = [1, 2, 3], 7
    iters = []
    for n in range(1, nsteps + 1):
        o = min(n, 3)
        sem.convect(0, B[lag[0]], B[4], field0=0, nf=1, scale=-1.0)                       # bq = -rho (U.grad) T
        sem.bdf_ext(B[4], B[5], B[6], [B[c] for c in lag[:o]], AB[o], BD[o], 1.0 / dt, field0=0, nf=1)
        sem.dssum(B[4], 0)
        it, _ = sem.hmholtz(B[4], B[new], 0, kappa, BD[o][0] / dt, tol=1e-13, maxit=500)
        iters.append(it)
        lag, new = [new, lag[0], lag[1]], lag[2]
    T = B[lag[0]].download()[0][0].reshape(x.shape)
    B.close()
    sem.close()
    return T, iters


def _oracle_scalar_steps(x, y, z, glo, mask, geo, vel, T0, kappa, dt, nsteps, N=7):
    d = osem.dgll(N)
    dl = osem.dealias_setup(N, 3 * (N + 1) // 2, geo['rst'])
    cf = osem.set_convect(vel, dl)
    lag = [T0.copy(), 0 * T0, 0 * T0]
    e1, e2 = 0 * T0, 0 * T0
    for n in range(1, nsteps + 1):
        o = min(n, 3)
        bq = -osem.convect_dealiased(lag[0], cf, dl)
        osem.bdf_ext(bq, e1, e2, lag[:o], geo['bm1'], AB[o], BD[o], 1.0 / dt)
        rhs = osem.dssum(bq, glo)
        Tn, _, _ = osem.cggo(rhs, geo['g'], d, glo, mask, geo['bm1'], kappa, BD[o][0] / dt, tol=1e-13, maxit=500)
        lag = [Tn, lag[0], lag[1]]
    return lag[0]


def test_scalar_advection_diffusion_steps_match_oracle(ctx):
    N = 7
    x, y, z, glo, geo = _mesh((2, 2, 2), N, 0.04)
    x0, y0, z0, _ = osem.box_mesh(2, 2, 2, N)
    mask = osem.boundary_mask_box(None, x0, y0, z0)
    vel = [np.sin(np.pi * x) * np.cos(np.pi * y), -np.cos(np.pi * x) * np.sin(np.pi * y), 0.2 * np.sin(np.pi * z)]
    T0 = np.sin(np.pi * x0) * np.sin(2 * np.pi * y0) * np.sin(np.pi * z0) * mask
    T_dev, iters = _device_scalar_steps(ctx, x, y, z, glo, mask, geo['bm1'], vel, T0, 0.05, 2e-3, 5)
    T_ref = _oracle_scalar_steps(x, y, z, glo, mask, geo, vel, T0, 0.05, 2e-3, 5)
    assert relerr(T_dev, T_ref) <= 1e-9
    assert max(iters) < 200


def test_scalar_diffusion_decay_rate(ctx):
    """No flow: sin(pi x) sin(pi y) sin(pi z) decays like exp(-3 pi^2 kappa t)."""
    N = 7
    x, y, z, glo, geo = _mesh((2, 2, 2), N, 0.0)
    mask = osem.boundary_mask_box(None, x, y, z)
    kappa, dt, nsteps = 0.1, 1e-3, 20
    T0 = np.sin(np.pi * x) * np.sin(np.pi * y) * np.sin(np.pi * z) * mask
    T, _ = _device_scalar_steps(ctx, x, y, z, glo, mask, geo['bm1'], [0 * x, 0 * x, 0 * x], T0, kappa, dt, nsteps)
    amp = float(np.sum(T * T0 * geo['bm1']) / np.sum(T0 * T0 * geo['bm1']))
    assert abs(amp - np.exp(-3 * np.pi ** 2 * kappa * dt * nsteps)) <= 2e-5


# ---- 2-D: the reference's own meshes and base flows (tests/golden/*_mesh.npz) ----------------------------
@pytest.mark.parametrize('name', ['cyl', 'bfs'])
def test_dealiased_convection_2d_reference_base_flows(ctx, name):
    """(U . grad) U of the committed base flow on the reference's curved 2-D mesh, lxd = 3 lx1 / 2."""
    import json
    from pathlib import Path
    import nekstab_next_b200 as nb
    gold = Path(__file__).resolve().parent / 'golden'
    N = json.loads((gold / 'known_answers.json').read_text())[name]['N']
    g = np.load(gold / f'{name}_mesh.npz')
    x, y, u, v, glo = g['x'], g['y'], g['u'], g['v'], g['glo'].astype(np.int64)
    geo = osem.geometry(N, x, y)
    dl = osem.dealias_setup(N, 3 * (N + 1) // 2, geo['rst'])
    cf = osem.set_convect([u, v], dl)
    ref = [osem.convect_dealiased(a, cf, dl) for a in (u, v)]
    sem = nb.Sem(ctx, N, x, y, None, mask=None, glo_num=glo)
    lay = nb.Layout(ctx, [x.size, x.size], [True, True])
    lay.set_weight([geo['bm1'], geo['bm1']])
    B = nb.Basis(lay, 2)
    sem.dealias_setup()
    B[0].upload([u, v])
    sem.set_convect(0, B[0])
    sem.convect(0, B[0], B[1], field0=0, nf=2)
    out, _ = B[1].download()
    for f in range(2):
        assert relerr(out[f].reshape(x.shape), ref[f]) <= 1e-12
    # accumulate with a scale on one field
    sem.convect(0, B[0], B[1], field0=1, nf=1, scale=2.5, accumulate=True)
    out2, _ = B[1].download()
    assert relerr(out2[1].reshape(x.shape), 3.5 * ref[1]) <= 1e-12
    assert np.array_equal(out2[0], out[0])
    B.close()
    sem.close()


def test_dealiased_convection_2d_box_other_orders(ctx):
    import nekstab_next_b200 as nb
    for N, lxd in ((3, 0), (6, 0), (5, 10)):
        x, y, glo = osem.box_mesh_2d(3, 2, N, deform=0.04)
        geo = osem.geometry(N, x, y)
        ld = lxd if lxd else 3 * (N + 1) // 2
        dl = osem.dealias_setup(N, ld, geo['rst'])
        rng = np.random.default_rng(N)
        vel = [rng.standard_normal(x.shape), np.cos(x) * y]
        w = rng.standard_normal(x.shape)
        ref = osem.convect_dealiased(w, osem.set_convect(vel, dl), dl)
        sem = nb.Sem(ctx, N, x, y, None, glo_num=glo)
        lay = nb.Layout(ctx, [x.size, x.size], [True, True])
        B = nb.Basis(lay, 3)
        sem.dealias_setup(lxd)
        B[0].upload(vel)
        sem.set_convect(1, B[0])
        B[1].upload([w, w])
        sem.convect(1, B[1], B[2], field0=1, nf=1)
        assert relerr(B[2].download()[0][1].reshape(x.shape), ref) <= 1e-12
        B.close()
        sem.close()


# ---- the device time-stepper as the Arnoldi operator (exponential_prop%matvec structure) --------------------
def _ramp_amplification(lam, dt, nsteps):
    """y_nsteps / y_0 of the BDF1 -> BDF2 -> BDF3 start-up sequence for y' = -lam y."""
    y = [1.0]
    for n in range(1, nsteps + 1):
        o = min(n, 3)
        rhs = sum(BD[o][i + 1] * y[-1 - i] for i in range(o)) / dt
        y.append(rhs / (BD[o][0] / dt + lam))
    return y[-1]


def test_stepper_operator_arnoldi_matches_oracle(ctx):
    """Arnoldi on Phi = (3 advection-diffusion steps): H against the oracle's Arnoldi on the same composition."""
    import nekstab_next_b200 as nb
    from oracle import krylov as okr
    N, K, nsteps, kappa, dt = 7, 4, 3, 0.05, 5e-3
    x, y, z, glo, geo = _mesh((2, 2, 2), N, 0.04)
    x0, y0, z0, _ = osem.box_mesh(2, 2, 2, N)
    mask = osem.boundary_mask_box(None, x0, y0, z0)
    vel = [np.sin(np.pi * x) * np.cos(np.pi * y), -np.cos(np.pi * x) * np.sin(np.pi * y), 0.2 * np.sin(np.pi * z)]
    rng = np.random.default_rng(3)
    q0 = osem.dssum(rng.standard_normal(x.shape), glo) / osem.multiplicity(glo) * mask
    c = okr.Ctx(bm1s=geo['bm1'], in_dot=[True], time_in_dot=False)

    def omatvec(q):
        return okr.KVec([_oracle_scalar_steps(x, y, z, glo, mask, geo, vel, q.f[0], kappa, dt, nsteps)], q.time)

    seed = okr.KVec([q0.copy()], 0.0)
    okr.k_normalize(c, seed)
    Qo = [okr.k_zero_like(seed) for _ in range(K + 1)]
    okr.k_copy(Qo[0], seed)
    Ho = np.zeros((K + 1, K))
    okr.arnoldi_factorization(c, omatvec, Qo, Ho, 1, K, K)

    sem = nb.Sem(ctx, N, x, y, z, mask=mask, glo_num=glo)
    lay3 = nb.Layout(ctx, [x.size] * 3, [True] * 3)
    Bv = nb.Basis(lay3, 1)
    sem.dealias_setup()
    Bv[0].upload(vel)
    sem.set_convect(0, Bv[0])
    lay = nb.Layout(ctx, [x.size], [True])
    lay.set_weight([geo['bm1']])
    Q = nb.Basis(lay, K + 1)
    op = nb.stepper_operator(sem, lay, 1, 0, kappa, dt, nsteps, tol=1e-13)
    Q[0].upload(seed.f)
    H = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(Q, H, 1, K, K, op)
    assert np.max(np.abs(H - Ho)) <= 1e-8 * np.max(np.abs(Ho))
    assert op.count() == K
    op.close(); Q.close(); Bv.close(); sem.close()


def test_stepper_operator_leading_growth_rate(ctx):
    """Pure diffusion: the leading Ritz value of the propagator is the BDF-ramp amplification of the slowest mode,
    and log(lambda)/tau recovers its decay rate -3 pi^2 kappa (what linear_stability_analysis prints,
    core/linear_stab.f90:72)."""
    import nekstab_next_b200 as nb
    N, kappa, dt, nsteps = 7, 1.0, 2.5e-3, 40
    x, y, z, glo, geo = _mesh((2, 2, 2), N, 0.0)
    mask = osem.boundary_mask_box(None, x, y, z)
    sem = nb.Sem(ctx, N, x, y, z, mask=mask, glo_num=glo)
    lay = nb.Layout(ctx, [x.size], [True])
    lay.set_weight([geo['bm1']])
    Q = nb.Basis(lay, 9)
    op = nb.stepper_operator(sem, lay, 1, -1, kappa, dt, nsteps, tol=1e-13)
    rng = np.random.default_rng(0)
    q0 = osem.dssum(rng.standard_normal(x.shape), glo) / osem.multiplicity(glo) * mask
    Q[0].upload([q0])
    nb.k_normalize(Q[0])
    vals, vecs, res, k, nconv, H = nb.eigs(Q, op, 8, nev=1, tol=1e-9)
    lead = vals[np.argmin(res)]
    lam = 3 * np.pi ** 2 * kappa
    assert nconv >= 1 and abs(lead.imag) <= 1e-10
    assert abs(lead.real - _ramp_amplification(lam, dt, nsteps)) <= 1e-7
    tau = dt * nsteps
    assert abs(np.log(lead.real) / tau + lam) <= 0.02 * lam      # time-discretisation error of the start-up steps
    op.close(); Q.close(); sem.close()


# ---- the adjoint side: transposed convection, discrete adjoint stepper, device-resident svds ------------------
@pytest.mark.parametrize('nel,N', [((2, 2, 2), 7), ((2, 3, 2), 5), ((3, 2), 5)])
def test_transposed_convection_matches_oracle_and_is_the_transpose(ctx, nel, N):
    """nsb_sem_convect_t vs the oracle's convect_dealiased_t, and sum v (C u) = sum u (C^T v) on the device."""
    import nekstab_next_b200 as nb
    dim = len(nel)
    if dim == 3:
        x, y, z, glo, geo = _mesh(nel, N, 0.04)
        coords = (x, y, z)
        vel = [np.sin(np.pi * x) * np.cos(np.pi * y), -np.cos(np.pi * x) * np.sin(np.pi * y), 0.2 * np.sin(np.pi * z)]
    else:
        x, y, glo = osem.box_mesh_2d(*nel, N, deform=0.04)
        geo = osem.geometry(N, x, y)
        coords = (x, y, None)
        vel = [1.0 + 0.3 * np.sin(np.pi * y), 0.4 * np.cos(np.pi * x)]
    dl = osem.dealias_setup(N, 3 * (N + 1) // 2, geo['rst'])
    cf = osem.set_convect(vel, dl)
    rng = np.random.default_rng(11)
    u, v = rng.standard_normal(x.shape), rng.standard_normal(x.shape)
    sem = nb.Sem(ctx, N, *coords, mask=None, glo_num=glo)
    nv = len(vel)
    lay = nb.Layout(ctx, [x.size] * nv, [True] * nv)
    lay.set_weight([geo['bm1']] * nv)
    B = nb.Basis(lay, 4)
    sem.dealias_setup()
    B[0].upload(vel)
    sem.set_convect(0, B[0])
    B[1].upload([u] + [0 * u] * (nv - 1))
    B[2].upload([v] + [0 * u] * (nv - 1))
    sem.convect(0, B[1], B[3], field0=0, nf=1)
    Cu = B[3].download()[0][0].reshape(x.shape)
    sem.convect_t(0, B[2], B[3], field0=0, nf=1)
    Ctv = B[3].download()[0][0].reshape(x.shape)
    assert relerr(Ctv, osem.convect_dealiased_t(v, cf, dl)) <= 1e-12
    assert abs(np.sum(v * Cu) - np.sum(u * Ctv)) <= 1e-12 * np.sqrt(np.sum(Cu * Cu) * np.sum(v * v))
    B.close(); sem.close()


def _bfs_problem():
    import json
    from pathlib import Path
    gold = Path(__file__).resolve().parent / 'golden'
    N = json.loads((gold / 'known_answers.json').read_text())['bfs']['N']
    g = np.load(gold / 'bfs_mesh.npz')
    x, y, glo = g['x'], g['y'], g['glo'].astype(np.int64)
    geo = osem.geometry(N, x, y)
    # Dirichlet on the walls / inflow like the reference's v1mask: here every node on the domain boundary
    mult = osem.multiplicity(glo)
    lx = N + 1
    edge = np.zeros(x.shape, dtype=bool)
    edge[:, 0, :] = edge[:, -1, :] = edge[:, :, 0] = edge[:, :, -1] = True
    # an element-edge node that belongs to one element only along a whole edge lies on the domain boundary
    bnd = np.zeros(x.shape)
    face_ids = [(slice(None), 0, slice(None)), (slice(None), lx - 1, slice(None)),
                (slice(None), slice(None), 0), (slice(None), slice(None), lx - 1)]
    for fi in face_ids:
        mid = mult[fi][:, lx // 2]                       # multiplicity of the mid-edge node: 1 = domain boundary
        sel = np.zeros(x.shape)
        sel[fi] = (mid == 1)[:, None]
        bnd = np.maximum(bnd, sel)
    mask = 1.0 - np.minimum(osem.dssum(bnd, glo), 1.0)
    return N, x, y, glo, geo, mask, [g['u'], g['v']]


@pytest.mark.parametrize('case', ['box3d', 'bfs'])
def test_adjoint_stepper_is_the_discrete_adjoint(ctx, case):
    """nsb_op_create_stepper_adjoint: A^+ v against the oracle's discrete adjoint, and <A u, v>_B = <u, A^+ v>_B to
    1e-10 on the device -- on a deformed 3-D box with a Taylor-Green-like flow and on the reference's
    backward-facing-step mesh with its own base flow (tests/golden/bfs_mesh.npz)."""
    import nekstab_next_b200 as nb
    if case == 'box3d':
        N, kappa, dt, nsteps = 7, 0.05, 5e-3, 4
        x, y, z, glo, geo = _mesh((2, 2, 2), N, 0.04)
        x0, y0, z0, _ = osem.box_mesh(2, 2, 2, N)
        mask = osem.boundary_mask_box(None, x0, y0, z0)
        vel = [np.sin(np.pi * x) * np.cos(np.pi * y), -np.cos(np.pi * x) * np.sin(np.pi * y), 0.2 * np.sin(np.pi * z)]
        coords = (x, y, z)
    else:
        N, x, y, glo, geo, mask, vel = _bfs_problem()
        kappa, dt, nsteps = 0.02, 2e-3, 4
        coords = (x, y, None)
    dl = osem.dealias_setup(N, 3 * (N + 1) // 2, geo['rst'])
    cf = osem.set_convect(vel, dl)
    rng = np.random.default_rng(8)
    vm = 1.0 / osem.multiplicity(glo)
    u = osem.dssum(rng.standard_normal(x.shape), glo) * vm * mask
    v = osem.dssum(rng.standard_normal(x.shape), glo) * vm * mask
    sem = nb.Sem(ctx, N, *coords, mask=mask, glo_num=glo)
    nv = len(vel)
    layv = nb.Layout(ctx, [x.size] * nv, [True] * nv)
    Bv = nb.Basis(layv, 1)
    sem.dealias_setup()
    Bv[0].upload(vel)
    sem.set_convect(0, Bv[0])
    lay = nb.Layout(ctx, [x.size], [True])
    lay.set_weight([geo['bm1']])
    Q = nb.Basis(lay, 4)
    fwd = nb.stepper_operator(sem, lay, 1, 0, kappa, dt, nsteps, tol=1e-14, maxit=3000)
    adj = nb.stepper_operator(sem, lay, 1, 0, kappa, dt, nsteps, tol=1e-14, maxit=3000, adjoint=True)
    Q[0].upload([u]); Q[1].upload([v])
    fwd.matvec(Q[0], Q[2])
    adj.matvec(Q[1], Q[3])
    Atv = Q[3].download()[0][0].reshape(x.shape)
    ref = osem.scalar_steps_adjoint(glo, mask, geo, N, cf, dl, v, kappa, dt, nsteps, tol=1e-14)
    assert relerr(Atv, ref) <= 1e-9
    lhs, rhs = nb.k_dot(Q[2], Q[1]), nb.k_dot(Q[0], Q[3])
    assert abs(lhs - rhs) <= 1e-10 * nb.k_norm(Q[0]) * nb.k_norm(Q[1])
    for o in (fwd, adj, Q, Bv, sem):
        o.close()


def test_svds_device_resident_transient_growth(ctx):
    """transient_growth_analysis (core/linear_stab.f90:112) with both A and A^+ on the device: nsb_svds over the
    forward / adjoint stepper pair against the oracle's svds over the oracle steppers; the leading singular value
    squared is the optimal energy gain G(tau) of the advection-diffusion problem."""
    import nekstab_next_b200 as nb
    from oracle import krylov as okr
    N, kappa, dt, nsteps, K = 5, 0.05, 5e-3, 3, 6
    x, y, z, glo, geo = _mesh((2, 2, 2), N, 0.04)
    x0, y0, z0, _ = osem.box_mesh(2, 2, 2, N)
    mask = osem.boundary_mask_box(None, x0, y0, z0)
    vel = [np.sin(np.pi * x) * np.cos(np.pi * y), -np.cos(np.pi * x) * np.sin(np.pi * y), 0.2 * np.sin(np.pi * z)]
    dl = osem.dealias_setup(N, 3 * (N + 1) // 2, geo['rst'])
    cf = osem.set_convect(vel, dl)
    c = okr.Ctx(bm1s=geo['bm1'], in_dot=[True], time_in_dot=False)
    rng = np.random.default_rng(2)
    u0 = okr.KVec([osem.dssum(rng.standard_normal(x.shape), glo) / osem.multiplicity(glo) * mask], 0.0)
    okr.k_normalize(c, u0)
    so, _, _, reso, ko, _ = okr.svds(
        c, lambda q: okr.KVec([osem.scalar_steps(glo, mask, geo, N, cf, dl, q.f[0], kappa, dt, nsteps)], q.time),
        lambda q: okr.KVec([osem.scalar_steps_adjoint(glo, mask, geo, N, cf, dl, q.f[0], kappa, dt, nsteps)], q.time),
        u0, K, 2, 1e-12)
    sem = nb.Sem(ctx, N, x, y, z, mask=mask, glo_num=glo)
    layv = nb.Layout(ctx, [x.size] * 3, [True] * 3)
    Bv = nb.Basis(layv, 1)
    sem.dealias_setup()
    Bv[0].upload(vel)
    sem.set_convect(0, Bv[0])
    lay = nb.Layout(ctx, [x.size], [True])
    lay.set_weight([geo['bm1']])
    U, V = nb.Basis(lay, K + 1), nb.Basis(lay, K)
    fwd = nb.stepper_operator(sem, lay, 1, 0, kappa, dt, nsteps, tol=1e-13)
    adj = nb.stepper_operator(sem, lay, 1, 0, kappa, dt, nsteps, tol=1e-13, adjoint=True)
    U[0].upload(u0.f)
    sig, uv, vv, res, k, nconv, Bm = nb.svds(U, V, fwd, adj, K, 2, 1e-12)
    assert k == ko
    assert np.max(np.abs(sig[:3] - so[:3])) <= 1e-8 * so[0]
    assert np.all(np.diff(sig) <= 1e-14) and sig[0] < 1.0          # a decaying problem: gain below one
    for o in (fwd, adj, U, V, Bv, sem):
        o.close()
