"""Error behaviour of the C ABI (codes instead of aborts) and randomised layouts."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

pytestmark = pytest.mark.gpu


def test_argument_errors_return_codes(ctx):
    import nekstab_next_b200 as nb
    lay = nb.Layout(ctx, [100, 50], [True, False])
    lay.set_weight([np.ones(100)])
    B = nb.Basis(lay, 4)
    other = nb.Layout(ctx, [100, 50], [True, False])
    B2 = nb.Basis(other, 2)
    for fn in (lambda: B[7].zero(), lambda: B[0].dot(nb.nek_dvector(B, 9)), lambda: nb.k_copy(B[0], B2[0]),
               lambda: nb.orthonormalize(B, 5, 3), lambda: nb.orthonormalize(B, 2, 1),
               lambda: nb.k_matmul(B[0], B, np.ones(2), 2), lambda: B.rotate(9, np.eye(9)),
               lambda: B[0].upload([np.ones(7), None])):
        with pytest.raises((nb.NsbError, ValueError)) as e:
            fn()
        if isinstance(e.value, nb.NsbError):
            assert e.value.code == -1 and str(e.value)
    with pytest.raises(ValueError):
        lay.set_weight([np.ones(99)])
    with pytest.raises(nb.NsbError):
        nb.Layout(ctx, [10, -1], [True, True])
    # the library is still usable after errors
    B[0].upload([np.arange(100.0), np.ones(50)])
    assert abs(B[0].dot(B[0]) - np.sum(np.arange(100.0) ** 2)) < 1e-9


def test_operator_errors(ctx):
    import nekstab_next_b200 as nb
    from helpers import BoxProblem
    P = BoxProblem(nel=(2, 2, 2), N=3, nfields=1)
    lay, B, S, op = P.gpu(ctx, 3)
    with pytest.raises(nb.NsbError):
        op.matvec(B[0], B[0])                       # in place
    wrong = nb.Layout(ctx, [P.npts + 1], [True])
    Bw = nb.Basis(wrong, 2)
    with pytest.raises(nb.NsbError):
        S.axhelm(Bw[0], Bw[1], 0, 1.0, 0.0)         # field length does not match the mesh
    with pytest.raises(nb.NsbError):
        nb.Sem(ctx, 12, *P.coords, mask=None, glo_num=P.glo)   # order not supported


def test_host_callback_failure_is_reported(ctx):
    import nekstab_next_b200 as nb
    lay = nb.Layout(ctx, [64], [True])
    lay.set_weight([np.ones(64)])
    B = nb.Basis(lay, 4)
    B[0].upload([np.ones(64) / 8.0])

    def bad(fields, t):
        raise RuntimeError('time-stepper blew up')

    op = nb.host_operator(lay, bad)
    H = np.zeros((4, 3), order='F')
    with pytest.raises(nb.NsbError):
        nb.arnoldi_factorization(B, H, 1, 2, 3, op)


def test_breakdown_is_reported(ctx):
    """f in span(Q): zero residual -> NSB_EBREAKDOWN instead of a division by zero."""
    import nekstab_next_b200 as nb
    lay = nb.Layout(ctx, [64], [True])
    lay.set_weight([np.ones(64)])
    B = nb.Basis(lay, 4)
    B[0].upload([np.ones(64) / 8.0])
    op = nb.host_operator(lay, lambda f, t: ([np.zeros(64)], t))
    H = np.zeros((4, 3), order='F')
    with pytest.raises(nb.NsbError) as e:
        nb.arnoldi_factorization(B, H, 1, 2, 3, op)
    assert e.value.code in (-7, -4)


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(lens=st.lists(st.integers(min_value=0, max_value=3000), min_size=1, max_size=6),
       dots=st.lists(st.booleans(), min_size=6, max_size=6), tdot=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_random_layouts_roundtrip_dot_axpby(ctx, lens, dots, tdot, seed):
    import nekstab_next_b200 as nb
    dots = dots[:len(lens)]
    if not any(d and n > 0 for d, n in zip(dots, lens)):
        dots[0] = True
        lens[0] = max(lens[0], 1)
    rng = np.random.default_rng(seed)
    lay = nb.Layout(ctx, lens, dots, time_in_dot=tdot)
    ws = [rng.random(n) for n, d in zip(lens, dots) if d]
    lay.set_weight(ws)
    B = nb.Basis(lay, 3)
    fa = [rng.standard_normal(n) for n in lens]
    fb = [rng.standard_normal(n) for n in lens]
    ta, tb = float(rng.standard_normal()), float(rng.standard_normal())
    B[0].upload(fa, ta)
    B[1].upload(fb, tb)
    ref = sum(np.sum(fa[i] * w * fb[i]) for i, w in zip([i for i, d in enumerate(dots) if d], ws)) + (ta * tb if tdot else 0.0)
    scale = sum(np.sum(np.abs(fa[i] * w * fb[i])) for i, w in zip([i for i, d in enumerate(dots) if d], ws)) + abs(ta * tb) + 1e-300
    assert abs(B[0].dot(B[1]) - ref) <= 1e-13 * scale
    B[0].axpby(0.5, B[1], -2.0, skip_time=False)
    out, t = B[0].download()
    for o, a, b in zip(out, fa, fb):
        assert np.allclose(o, 0.5 * a - 2.0 * b, rtol=1e-15, atol=1e-15)
    assert abs(t - (0.5 * ta - 2.0 * tb)) < 1e-14
    assert lay.ld % 1024 == 0 and lay.ndot % 1024 == 0 and lay.ndof_dot == sum(n for n, d in zip(lens, dots) if d)
    B.close()
    lay.close()


def test_errors_of_the_widened_entry_points(ctx):
    """svds / Helmholtz / EXT-BDF / time-stepper operator: bad arguments come back as NSB_EINVAL, and the
    library keeps working afterwards."""
    import nekstab_next_b200 as nb
    from helpers import BoxProblem
    P = BoxProblem(nel=(2, 1, 1), N=3, nfields=3)
    lay, B, S, op = P.gpu(ctx, 6)
    V = nb.Basis(lay, 2)
    cases = [
        lambda: nb.svds(B, V, op, op, 5, 1, 1e-8),                       # V has fewer than k_dim columns
        lambda: nb.svds(B, V, op, op, 0, 1, 1e-8),
        lambda: nb.eigs(B, op, 9, 1, 1e-8),                              # basis has 6 columns
        lambda: S.hmholtz_vec(B[0], B[1], 0, 4, 1.0, 1.0),               # more than three systems
        lambda: S.hmholtz_vec(B[0], B[1], 1, 3, 1.0, 1.0),               # fields 1..3 do not exist
        lambda: S.hmholtz(B[0], B[0], 0, 1.0, 1.0),                      # in place
        lambda: S.bdf_ext(B[0], B[1], B[1], [B[3]], [1, 0, 0], [1, 1], 1.0),          # e1 == e2
        lambda: S.bdf_ext(B[0], B[1], B[2], [B[0]], [1, 0, 0], [1, 1], 1.0),          # velocity aliases bf
        lambda: S.bdf_ext(B[0], B[1], B[2], [B[3]] * 4, [1, 0, 0], [1] * 5, 1.0),     # nbd > 3
        lambda: nb.stepper_operator(S, lay, 1, 0, 0.1, 1e-3, 3),         # slot 0 was never set
        lambda: nb.stepper_operator(S, lay, 1, -1, -0.1, 1e-3, 3),       # negative diffusivity
        lambda: nb.stepper_operator(S, lay, 4, -1, 0.1, 1e-3, 3),        # more fields than the layout has
        lambda: nb.stepper_operator(S, lay, 1, -1, 0.1, 1e-3, 0),        # no steps
    ]
    for fn in cases:
        with pytest.raises(nb.NsbError) as e:
            fn()
        assert e.value.code == -1 and str(e.value)
    step = nb.stepper_operator(S, lay, 2, -1, 0.1, 1e-3, 2)
    wrong = nb.Basis(nb.Layout(ctx, [P.npts] * 3, [True] * 3), 2)
    with pytest.raises(nb.NsbError):
        step.matvec(wrong[0], wrong[1])                                  # operator built for another layout
    q = P.random_kvec()
    from helpers import upload
    upload(B[0], q)
    step.matvec(B[0], B[1])                                              # still works; third field carried through
    out = B[1].download()[0]
    assert np.array_equal(out[2], q.f[2].ravel()) and not np.array_equal(out[0], q.f[0].ravel())
    assert np.all(np.isfinite(out[0])) and np.all(np.isfinite(out[1]))
    step.close()
