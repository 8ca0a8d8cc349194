"""GPU parity: nek_dvector BLAS-1 set and BM1-weighted inner products vs the oracle
(tolerance 1e-12 relative, fp64 -- BASELINE.json north_star)."""
import numpy as np
import pytest

from helpers import BoxProblem, upload, download, relerr
from oracle import krylov as okr

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope='module')
def prob(ctx):
    P = BoxProblem(nel=(3, 2, 2), N=5, nfields=3, pressure=True, time_in_dot=True, seed=1)
    lay, basis, semg, op = P.gpu(ctx, 8)
    return P, lay, basis


def test_roundtrip_and_layout(prob):
    P, lay, B = prob
    a = P.random_kvec()
    upload(B[0], a)
    b = download(B[0])
    for x, y in zip(a.f, b.f):
        assert np.array_equal(x.ravel(), y)
    assert b.time == a.time
    assert lay.ld % 1024 == 0 and lay.ndot % 1024 == 0 and lay.ndof_dot == 3 * P.npts


def test_dot_norm(prob):
    import nekstab_next_b200 as nb
    P, lay, B = prob
    c = P.octx()
    a, b = P.random_kvec(), P.random_kvec()
    upload(B[0], a)
    upload(B[1], b)
    ref = okr.k_dot(c, a, b)
    assert abs(nb.k_dot(B[0], B[1]) - ref) <= TOL * abs(ref) + 1e-300
    # pressure is never in the inner product (core/krylov_subspace.f90:40-49)
    a2 = a.copy()
    a2.f[3][:] += 100.0
    upload(B[2], a2)
    assert abs(B[2].dot(B[1]) - ref) <= TOL * abs(ref)
    assert abs(nb.k_norm(B[0]) - okr.k_norm(c, a)) <= TOL * okr.k_norm(c, a)


def test_dot_semidefinite_weight(ctx):
    """bm1s may contain zeros (sponge, core/forcing.f90:102-104)."""
    import nekstab_next_b200 as nb
    P = BoxProblem(nel=(2, 2, 2), N=4, nfields=2, seed=5)
    w = P.bm1.copy()
    w[P.coords[0] > 0.7] = 0.0
    lay = nb.Layout(ctx, [P.npts, P.npts], [True, True])
    lay.set_weight([w, w])
    B = nb.Basis(lay, 2)
    a, b = P.random_kvec(), P.random_kvec()
    upload(B[0], a)
    upload(B[1], b)
    c = okr.Ctx(bm1s=w, in_dot=[True, True])
    ref = okr.k_dot(c, a, b)
    assert abs(B[0].dot(B[1]) - ref) <= TOL * abs(ref)
    B.close()
    lay.close()


def test_nan_is_an_error(prob):
    import nekstab_next_b200 as nb
    P, lay, B = prob
    a = P.random_kvec()
    a.f[0].ravel()[7] = np.nan
    upload(B[3], a)
    with pytest.raises(nb.NsbError) as e:
        B[3].dot(B[3])
    assert e.value.code == -4


def test_blas1(prob):
    import nekstab_next_b200 as nb
    P, lay, B = prob
    a, b, r = P.random_kvec(), P.random_kvec(), P.random_kvec()

    def check(vec, ref):
        got = download(vec)
        for x, y in zip(got.f, ref.f):
            assert relerr(x, y.ravel()) <= 1e-15
        assert abs(got.time - ref.time) <= 1e-15 * max(1.0, abs(ref.time))

    upload(B[0], a); upload(B[1], b); upload(B[2], r)
    # axpby: self <- alpha*self + beta*vec, %time untouched (quirk of real_axpby)
    B[0].axpby(0.3, B[1], -1.7)
    ra = a.copy(); okr.axpby(ra, 0.3, b, -1.7, skip_time=True); check(B[0], ra)
    B[0].axpby(2.0, B[1], 0.5, skip_time=False)
    okr.axpby(ra, 2.0, b, 0.5, skip_time=False); check(B[0], ra)
    B[0].scal(-0.25); okr.k_cmult(ra, -0.25); check(B[0], ra)
    nb.k_add2(B[0], B[1]); okr.k_add2(ra, b); check(B[0], ra)
    nb.k_sub2(B[0], B[2]); okr.k_sub2(ra, r); check(B[0], ra)
    nb.k_sub3(B[4], B[1], B[2]); rs = okr.k_zero_like(a); okr.k_sub3(rs, b, r); check(B[4], rs)
    nb.k_copy(B[5], B[0]); check(B[5], ra)
    nb.k_zero(B[5]); check(B[5], okr.k_zero_like(a))
    alpha = nb.k_normalize(B[0])
    ref_alpha = okr.k_normalize(P.octx(), ra)
    assert abs(alpha - ref_alpha) <= TOL * ref_alpha
    got = download(B[0])
    for x, y in zip(got.f, ra.f):
        assert relerr(x, y.ravel()) <= 1e-13


def test_empty_and_ragged_fields(ctx):
    """Zero-length fields and lengths that are not multiples of anything."""
    import nekstab_next_b200 as nb
    lens = [1, 0, 1023, 1025, 7]
    lay = nb.Layout(ctx, lens, [True, True, True, False, True], time_in_dot=True)
    rng = np.random.default_rng(0)
    ws = [rng.random(n) for n, d in zip(lens, [1, 1, 1, 0, 1]) if d]
    lay.set_weight(ws)
    B = nb.Basis(lay, 2)
    fa = [rng.standard_normal(n) for n in lens]
    fb = [rng.standard_normal(n) for n in lens]
    B[0].upload(fa, 2.0)
    B[1].upload(fb, -3.0)
    ref = sum(np.sum(fa[i] * w * fb[i]) for i, w in zip([0, 1, 2, 4], ws)) + 2.0 * -3.0
    assert abs(B[0].dot(B[1]) - ref) <= 1e-13 * abs(ref)
    out, t = B[1].download()
    assert all(np.array_equal(o, f) for o, f in zip(out, fb)) and t == -3.0
    B.close()
    lay.close()


def test_complex_vector_is_a_layout(ctx):
    """cmplx_nek_vector (core/nek_vectors.f90:33-43, 140-201) = two real vectors; its dot is
    real_dot(re, re) + real_dot(im, im), scal / axpby act on both parts: on the device that is simply a
    layout listing the fields of the real part followed by those of the imaginary part (resolvent path,
    core/linear_stab.f90:124-160)."""
    import nekstab_next_b200 as nb
    rng = np.random.default_rng(12)
    n, n2 = 3000, 1700                      # velocity points, pressure points
    bm1 = rng.random(n) + 0.1
    lens = [n, n, n2, n, n, n2]             # re: vx vy pr | im: vx vy pr
    dots = [True, True, False, True, True, False]
    lay = nb.Layout(ctx, lens, dots)
    lay.set_weight([bm1] * 4)
    B = nb.Basis(lay, 2)
    a = [rng.standard_normal(m) for m in lens]
    b = [rng.standard_normal(m) for m in lens]
    B[0].upload(a)
    B[1].upload(b)
    ref = sum(np.sum(a[i] * bm1 * b[i]) for i in (0, 1, 3, 4))
    assert abs(B[0].dot(B[1]) - ref) <= TOL * abs(ref)
    B[0].axpby(0.25, B[1], -3.0)
    got, _ = B[0].download()
    for i in range(6):
        assert relerr(got[i], 0.25 * a[i] - 3.0 * b[i]) <= 1e-15
    B.close()
