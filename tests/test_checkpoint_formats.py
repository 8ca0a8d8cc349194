"""Checkpoint / restart wire formats (HES*, Spectre_*, KRY* field files) against the reference's
own files and format statements."""
from pathlib import Path

import numpy as np
import pytest

from nekstab_next_b200 import checkpoint as ck
from oracle import nekfld

REF = Path('/root/reference/examples')


def test_fortran_e_format():
    # values as gfortran prints them with E15.7
    assert ck.fortran_e(1.0) == '  0.1000000E+01'
    assert ck.fortran_e(-0.5) == ' -0.5000000E+00'
    assert ck.fortran_e(0.0) == '  0.0000000E+00'
    assert ck.fortran_e(123456.789) == '  0.1234568E+06'
    assert ck.fortran_e(9.99999999) == '  0.1000000E+02'       # mantissa rounds up to 1.0
    assert ck.fortran_e(-3.1e-7) == ' -0.3100000E-06'
    assert all(len(ck.fortran_e(v)) == 15 for v in (1e-30, -1e30, 7.0))


def test_spectrum_roundtrip(tmp_path):
    vals = np.array([0.98 + 0.1j, 0.98 - 0.1j, -0.5 + 0j])
    res = np.array([1e-9, 1e-9, 3.2e-4])
    p = tmp_path / 'Spectre_Hd0100.dat'
    ck.write_spectrum(p, vals, res)
    lines = p.read_text().splitlines()
    assert all(len(l) == 45 for l in lines)
    v2, r2 = ck.read_spectrum(p)
    assert np.allclose(v2, vals, rtol=1e-6) and np.allclose(r2, res, rtol=1e-6)


def test_hessenberg_roundtrip_and_subsample(tmp_path):
    rng = np.random.default_rng(0)
    k = 7
    H = np.triu(rng.standard_normal((k + 1, k)), -1)
    p = tmp_path / ck.hessenberg_name('1cyl', k)
    assert p.name == 'HES1cyl0007'
    ck.write_hessenberg(p, H, k)
    H2 = ck.read_hessenberg(p, k_dim=10, mstart=k)
    assert H2.shape == (11, 10) and np.array_equal(H2[:k + 1, :k], H) and not H2[k + 1:].any()
    H3 = ck.read_hessenberg(p, k_dim=5, mstart=k)     # k_dim < mstart: subsample (core/eigensolvers.f90:248-254)
    assert np.array_equal(H3, H[:6, :5])
    with pytest.raises(ValueError):
        ck.read_hessenberg(p, k_dim=10, mstart=k + 1)


def test_field_file_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    nel, lx = 5, 4
    for nz in (1, lx):
        shp = (nel, nz, lx, lx) if nz > 1 else (nel, lx, lx)
        nd = 3 if nz > 1 else 2
        x = [rng.standard_normal(shp) for _ in range(nd)]
        u = [rng.standard_normal(shp) for _ in range(nd)]
        pr = rng.standard_normal(shp)
        p = tmp_path / ck.field_name('KRY', 'box', 12)
        assert p.name == 'KRYbox0.f00012'
        ck.write_fld(p, dict(x=x, u=u, p=pr), lx, lx, nz, time=3.5, istep=7)
        for reader in (ck.read_fld, nekfld.read_fld):          # product reader and oracle reader agree
            f = reader(p)
            assert f['rdcode'] == 'XUP' and f['nel'] == nel and f['time'] == 3.5 and f['istep'] == 7
            assert all(np.array_equal(a, b) for a, b in zip(f['x'], x))
            assert all(np.array_equal(a, b) for a, b in zip(f['u'], u)) and np.array_equal(f['p'], pr)
        ck.write_fld(p, dict(u=u), lx, lx, nz, wdsize=4)
        f = ck.read_fld(p)
        assert f['wdsize'] == 4 and np.allclose(f['u'][0], u[0], rtol=1e-6, atol=1e-6)


@pytest.mark.skipif(not REF.exists(), reason='reference tree not present (GPU box)')
def test_reads_reference_field_files_and_rewrites_them_identically(tmp_path):
    src = REF / 'cylinder/BF_1cyl0.f00001'
    f = ck.read_fld(src)
    g = nekfld.read_fld(src)
    assert f['nel'] == 1996 and f['rdcode'] == 'XUP'
    assert np.array_equal(f['x'][1], g['x'][1]) and np.array_equal(f['u'][0], g['u'][0]) and np.array_equal(f['p'], g['p'])
    out = tmp_path / 'BF_copy0.f00001'
    ck.write_fld(out, dict(x=f['x'], u=f['u'], p=f['p']), f['nx'], f['ny'], f['nz'], time=f['time'], istep=f['istep'],
                 elmap=f['elmap'])
    assert src.read_bytes() == out.read_bytes()               # header and payload bit-identical


def test_log_transform_and_ns_spectrum(tmp_path):
    """eigenvalue of the propagator -> growth rate / frequency of the linearised operator
    (core/eigensolvers.f90:547-548, 860-869)."""
    T = 2.5
    mu = np.array([0.1 + 0.7j, 0.1 - 0.7j, -0.3 + 0j, 0.05 + 0j])
    lam = np.exp(mu * T)
    lam[2] = -abs(lam[2])                       # a negative real Ritz value: the reference keeps only log|x|
    lt = ck.log_transform(lam)
    assert np.allclose(lt[:2] / T, mu[:2]) and lt[3].imag == 0 and abs(lt[3].real / T - 0.05) < 1e-14
    assert lt[2].imag == 0 and abs(lt[2].real - np.log(abs(lam[2]))) < 1e-14
    p = tmp_path / 'Spectre_NSd.dat'
    ck.write_ns_spectrum(p, lam, np.array([1e-8, 1e-8, 2e-3, 5e-7]), T)
    v, r = ck.read_spectrum(p)
    assert np.allclose(v[:2], mu[:2], atol=1e-6) and np.allclose(r, [1e-8, 1e-8, 2e-3, 5e-7])
    assert all(len(line) == 46 for line in open(p))           # (3E15.7) + newline


def test_singvals_roundtrip(tmp_path):
    """Spectrum_S*.dat of transient_growth_analysis: sigma**2 and residual, (2E15.7) (core/linear_stab.f90:113-114)."""
    sig = np.array([63151.984, 12.5, 0.75]) ** 0.5
    res = np.array([1e-9, 3e-7, 1e-2])
    p = tmp_path / 'Spectrum_Sp.dat'
    ck.write_singvals(p, sig ** 2, res)
    s2, r2 = ck.read_singvals(p)
    assert np.allclose(s2, sig ** 2, rtol=1e-7) and np.allclose(r2, res, rtol=1e-7)
    lines = open(p).read().splitlines()
    assert all(len(ln) == 30 for ln in lines) and lines[0].split()[0] == '0.6315198E+05'


def test_hessenberg_through_the_c_abi(tmp_path, lib):
    """nsb_hessenberg_write / nsb_hessenberg_read (host-only entry points) against the Python module and the
    reference's list-directed layout, incl. Fortran D exponents and the k_dim < mstart subsampling."""
    import nekstab_next_b200 as nb
    rng = np.random.default_rng(3)
    k = 9
    H = np.asfortranarray(np.triu(rng.standard_normal((k + 1, k)), -1))
    p = tmp_path / ck.hessenberg_name('1cyl', k)
    nb.hessenberg_write(p, H, k)
    assert np.array_equal(ck.read_hessenberg(p, k, k), H)                 # C writer -> Python reader, exact
    assert np.array_equal(nb.hessenberg_read(p, k, k), H)                 # C writer -> C reader, exact
    H12 = nb.hessenberg_read(p, 12, k)                                    # restart into a larger space
    assert H12.shape == (13, 12) and np.array_equal(H12[:k + 1, :k], H) and not H12[:, k:].any()
    H5 = nb.hessenberg_read(p, 5, k)                                      # subsampling, k_dim < mstart
    assert np.array_equal(H5, H[:6, :5])
    q = tmp_path / 'HESfortran0003'                                       # what write(67,*) can look like
    q.write_text('  1.5D+00 -2.25d-01, 3.0\n 4.0   5.0E0 6.0\n7 8 9\n 10 11 12\n')
    Hf = nb.hessenberg_read(q, 3, 3)
    assert np.array_equal(Hf, np.array([[1.5, -0.225, 3], [4, 5, 6], [7, 8, 9], [10, 11, 12]]))
    with pytest.raises(nb.NsbError):
        nb.hessenberg_read(q, 3, 4)                                       # wrong number of values


@pytest.mark.gpu
def test_restart_from_reference_wire_formats(ctx, tmp_path):
    """The restart branch of krylov_schur (core/eigensolvers.f90:240-285) through the C ABI: a factorisation
    interrupted after mstart steps, checkpointed as HES<session>%04d + KRY<session>0.f%05d (Nek field files,
    fp64, elements stored in a permuted order), reloaded into a fresh basis and continued gives the H of the
    uninterrupted run."""
    import sys
    sys.path.insert(0, str(Path(__file__).parent))
    import nekstab_next_b200 as nb
    from helpers import BoxProblem, upload, download
    from oracle import krylov as okr
    P = BoxProblem(nel=(3, 2, 2), N=4, nfields=3, conv=True, seed=17)
    c = P.octx()
    K, ms = 10, 6
    lay, B, S, op = P.gpu(ctx, K + 1)
    q0 = P.random_kvec()
    okr.k_normalize(c, q0)
    upload(B[0], q0)
    Hfull = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(B, Hfull, 1, K, K, op)
    nel, lx = P.shape[0], P.shape[-1]
    perm = np.random.default_rng(1).permutation(nel)            # file order != local order
    for i in range(ms + 1):                                      # outpost of the first ms + 1 Krylov vectors
        v = download(B[i], P.shape)
        ck.write_fld(tmp_path / ck.field_name('KRY', 'box', i + 1), dict(u=[f[perm] for f in v.f[:3]]), lx, lx, lx,
                     time=0.5 * (i + 1), elmap=perm + 1)
    nb.hessenberg_write(tmp_path / ck.hessenberg_name('box', ms), Hfull, ms)
    B2 = nb.Basis(lay, K + 1)
    t = nb.fld_read_into(B2[K], tmp_path / ck.field_name('KRY', 'box', 2), nel, lglel=np.arange(1, nel + 1))
    assert abs(t - 1.0) < 1e-12
    assert np.array_equal(B2[K].download()[0][1], B[1].download()[0][1])   # exact: fp64 file, elements re-ordered
    H, mnext = nb.restart_load(B2, tmp_path, 'box', ms, K, nel, lglel=np.arange(1, nel + 1))
    assert mnext == ms + 1 and np.array_equal(H[:ms + 1, :ms], Hfull[:ms + 1, :ms])
    nb.arnoldi_factorization(B2, H, mnext, K, K, op)
    assert np.max(np.abs(H - Hfull)) <= 1e-13 * np.max(np.abs(Hfull))
    with pytest.raises(nb.NsbError):                              # an element the file does not hold
        nb.fld_read_into(B2[K], tmp_path / ck.field_name('KRY', 'box', 2), nel, lglel=np.arange(2, nel + 2))
    for o in (B2, op, S, B, lay):
        o.close()
