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
