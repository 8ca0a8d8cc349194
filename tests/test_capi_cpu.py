"""-m "not gpu": the C-ABI library loads, exports every declared symbol, host-only entry points
work, and compute entry points fail loudly (no CPU fallback)."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import krylov as okr, sem as osem

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    txt = (ROOT / 'include' / 'nekstab_b200.h').read_text()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(nsb_[a-z0-9_]+)\s*\(', txt)) - {'nsb_host_matvec_fn'})


def test_library_exports_every_header_symbol(lib):
    from nekstab_next_b200 import _capi
    syms = header_symbols()
    assert len(syms) > 50
    out = subprocess.run(['nm', '-D', '--defined-only', str(_capi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (nsb_[a-z0-9_]+)', out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f'header declares but library does not export: {missing}'
    undeclared = [s for s in syms if s not in _capi.PROTOTYPES]
    assert not undeclared, f'ctypes binding lacks: {undeclared}'
    assert lib.nsb_version() == 100


def test_ctypes_prototypes_match_the_header():
    """Every ctypes signature against its C prototype: argument count, by-value scalar kinds, pointer-ness."""
    import ctypes as C
    from nekstab_next_b200 import _capi
    text = re.sub(r'/\*.*?\*/', ' ', (ROOT / 'include' / 'nekstab_b200.h').read_text(), flags=re.S)
    byval = {'int': C.c_int, 'double': C.c_double, 'int64_t': C.c_int64, 'uint64_t': C.c_uint64}
    seen = 0
    for m in re.finditer(r'\b(int|int64_t|const char \*)\s*(nsb_\w+)\s*\(([^;{]*)\)\s*;', text):
        ret, name, args = m.group(1), m.group(2), ' '.join(m.group(3).split())
        params = [] if args in ('', 'void') else [a.strip() for a in args.split(',')]
        res, argtypes = _capi.PROTOTYPES[name]
        assert len(argtypes) == len(params), f'{name}: {len(argtypes)} ctypes arguments, {len(params)} C parameters'
        assert res is {'int': C.c_int, 'int64_t': C.c_int64, 'const char *': C.c_char_p}[ret], name
        for at, cp in zip(argtypes, params):
            if '*' in cp:
                assert at is C.c_void_p or at is C.c_char_p or hasattr(at, '_type_') and not issubclass(at, C._SimpleCData) \
                    or issubclass(at, C._Pointer), f'{name}: `{cp}` is a pointer, ctypes has {at}'
                continue
            ctype = ' '.join(cp.replace('const', '').split()[:-1])
            if re.fullmatch(r'nsb_\w+_t', ctype):
                assert at is C.c_void_p, f'{name}: handle `{cp}` must be c_void_p, ctypes has {at}'
            elif ctype == 'nsb_host_matvec_fn':
                assert at is _capi.HOST_MATVEC, name
            else:
                assert at is byval[ctype], f'{name}: `{cp}` needs {byval[ctype].__name__}, ctypes has {at}'
        seen += 1
    assert seen == len(_capi.PROTOTYPES), 'a ctypes prototype has no C declaration'


def test_library_is_sm100a_only(lib):
    from nekstab_next_b200 import _capi
    out = subprocess.run(['cuobjdump', '--list-elf', str(_capi.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_sass_shows_tma_mbarrier_and_fp64_tensor_cores(lib):
    """What the hot kernels claim to use is in the machine code: 2-D tensor TMA loads (fused orthogonalisation),
    bulk TMA copies and mbarrier waits (axhelm ring), fp64 tensor-core MMAs (axhelm contraction)."""
    from nekstab_next_b200 import _capi
    sass = subprocess.run(['cuobjdump', '-sass', str(_capi.LIB_PATH)], capture_output=True, text=True).stdout
    fn = None
    seen = {}
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            continue
        for op in ('UTMALDG', 'UBLKCP', 'DMMA', 'SYNCS.PHASECHK'):
            if op in line:
                seen.setdefault(op, set()).add(fn)
    assert any('fused_tma_reg' in f for f in seen.get('UTMALDG', ()))
    assert any('axhelm3d_dmma8' in f for f in seen.get('UBLKCP', ()))
    assert any('axhelm3d_dmma8' in f for f in seen.get('DMMA', ()))
    assert any('axhelm3d_dmma8' in f for f in seen.get('SYNCS.PHASECHK', ()))
    assert 'HGMMA' not in sass and 'HMMA' not in sass          # no legacy / Hopper tensor paths


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import nekstab_next_b200 as nb
    with pytest.raises(nb.NsbError) as e:
        nb.Context(device=0)
    assert e.value.code == -3 and 'no CPU fallback' in str(e.value)


def test_product_does_not_import_oracle():
    for p in (ROOT / 'nekstab_next_b200').rglob('*'):
        if p.suffix in ('.py', '.cu', '.cuh', '.h', '.cpp'):
            assert 'oracle' not in p.read_text(), f'{p} mentions the oracle'


def test_gll_matches_oracle(lib):
    import nekstab_next_b200 as nb
    for n in (1, 2, 3, 5, 7, 9, 11):
        z, w, D = nb.gll(n)
        z2, w2 = osem.gll(n)
        assert np.max(np.abs(z - z2)) < 1e-15 and np.max(np.abs(w - w2)) < 1e-14
        assert np.max(np.abs(D - osem.dgll(n))) < 1e-12


def test_dealias_matrices_match_oracle(lib):
    import nekstab_next_b200 as nb
    for N, lxd in ((7, 12), (5, 9), (4, 8), (3, 6), (7, 10)):
        zd, wd, J, Dg = nb.dealias_matrices(N, lxd)
        z2, w2 = osem.gl(lxd)
        assert np.max(np.abs(zd - z2)) < 1e-15 and np.max(np.abs(wd - w2)) < 1e-15
        zg, _ = osem.gll(N)
        assert np.max(np.abs(J - osem.interp_matrix(zg, z2))) < 1e-14
        assert np.max(np.abs(Dg - osem.deriv_matrix(z2))) < 1e-11


def test_lapack_wrapper_mirrors(lib):
    import nekstab_next_b200 as nb
    rng = np.random.default_rng(3)
    for n in (6, 12, 25):
        A = rng.standard_normal((n, n)) * (0.9 / np.sqrt(n))
        v1, l1 = nb.eig(A)
        v2, l2 = okr.eig(A)
        assert np.allclose(l1, l2, atol=1e-13) and np.allclose(v1, v2, atol=1e-12)
        assert np.all(np.diff(np.abs(l1)) <= 1e-14)          # sorted by decreasing magnitude
        assert np.allclose(A @ v1, v1 * l1[None, :], atol=1e-11)
        T1, Z1, s1 = nb.schur(A)
        T2, Z2, s2 = okr.schur(A)
        assert np.allclose(T1, T2, atol=1e-13) and np.allclose(Z1, Z2, atol=1e-13)
        assert np.allclose(Z1 @ T1 @ Z1.T, A, atol=1e-12)
        if n >= 8:
            sel1, c1 = nb.select_eigenvalues(s1, 0.1, 2)
            sel2, c2 = okr.select_eigenvalues(s2.copy(), 0.1, 2)
            assert c1 == c2 and np.array_equal(sel1, sel2) and c1 >= 6
            T1o, Z1o = nb.ordschur(T1, Z1, sel1)
            T2o, Z2o = okr.ordschur(T2, Z2, sel2)
            assert np.allclose(T1o, T2o, atol=1e-12) and np.allclose(Z1o @ T1o @ Z1o.T, A, atol=1e-11)
        B = rng.standard_normal((n + 1, n))
        b = rng.standard_normal(n + 1)
        assert np.allclose(nb.lstsq(B, b), okr.lstsq(B, b), atol=1e-12)
        assert np.allclose(nb.lstsq(B, b), np.linalg.lstsq(B, b, rcond=None)[0], atol=1e-10)


def test_svd_mirror(lib):
    import nekstab_next_b200 as nb
    rng = np.random.default_rng(5)
    for m, n in ((7, 7), (9, 5), (4, 8)):
        A = rng.standard_normal((m, n))
        U, S, V = nb.svd(A)
        assert np.allclose(S, np.linalg.svd(A, compute_uv=False), atol=1e-13)
        assert np.allclose(U @ np.diag(S) @ V.T, A, atol=1e-12)
        assert np.allclose(U.T @ U, np.eye(min(m, n)), atol=1e-12)


def test_select_eigenvalues_conjugate_pair_kept(lib):
    import nekstab_next_b200 as nb
    # the (nev+4)-th largest is half of a conjugate pair -> its partner is selected too (:747-749)
    vals = np.array([0.99, 0.95, 0.9 + 0.1j, 0.9 - 0.1j, 0.8, 0.7 + 0.2j, 0.7 - 0.2j, 0.5, 0.4, 0.3, 0.2, 0.1],
                    dtype=np.complex128)
    sel, cnt = nb.select_eigenvalues(vals, 0.05, 2)   # nev+4 = 6 largest
    sel2, cnt2 = okr.select_eigenvalues(vals.copy(), 0.05, 2)
    assert np.array_equal(sel, sel2) and cnt == cnt2
    assert sel[5] and sel[6]


def test_partition_ranges():
    from nekstab_next_b200.mesh import partition_range
    for total, gran in ((32768, 1024), (64, 16), (30, 6)):
        for P in (1, 2, 3, 4, 8):
            rs = [partition_range(total, r, P, gran) for r in range(P)]
            assert rs[0][0] == 0 and rs[-1][1] == total // gran * gran
            for (a, b), (c, d) in zip(rs[:-1], rs[1:]):
                assert b == c and (b - a) % gran == 0


def test_seed_noise_matches_reference_formula():
    """mth_rand / op_add_noise (core/utils.f90:297-359, 408-418): product vs oracle restatement."""
    from nekstab_next_b200 import seed
    x, y, z, glo = osem.box_mesh(2, 2, 2, 3, deform=0.02)
    a = seed.noise_fields((x, y, z))
    b = okr.op_add_noise((x, y, z), if3d=True)
    for p, q in zip(a, b):
        assert np.array_equal(p, q) and np.all(np.abs(p) <= 1.0)
    x2, y2, glo2 = osem.box_mesh_2d(3, 2, 4)
    a = seed.noise_fields((x2, y2))
    b = okr.op_add_noise((x2, y2), if3d=False)
    assert len(a) == 2 and all(np.array_equal(p, q) for p, q in zip(a, b))
    # spot value computed by hand from the formula: element 1, point (1,1,1) at the origin
    fc = seed.NOISE_FC[0]
    r = fc[0] * (1 + 0.0) + fc[1] * 1 + fc[2] * 1
    r = fc[0] * (1 + 0.0 * np.sin(r)) + fc[1] * 1 + fc[2] * 1
    assert abs(seed.noise_fields((x, y, z))[0][0, 0, 0, 0] - np.cos(1e3 * np.sin(1e3 * np.sin(r)))) < 1e-12


def test_pressure_mesh_matrices_match_oracle():
    """nsb_pressure_matrices (host-only): ixm12 / dxm12 of the P_N - P_N-2 pressure mesh vs the oracle's."""
    import nekstab_next_b200 as nb
    from oracle import sem as osem
    for N in (3, 4, 5, 7, 9):
        z2, w2, I12, D12 = nb.pressure_matrices(N)
        zo, wo = osem.gl(N - 1)
        z1, _ = osem.gll(N)
        Io = osem.interp_matrix(z1, zo)
        assert np.max(np.abs(z2 - zo)) <= 1e-14 and np.max(np.abs(w2 - wo)) <= 1e-14
        assert np.max(np.abs(I12 - Io)) <= 1e-13
        assert np.max(np.abs(D12 - Io @ osem.dgll(N))) <= 1e-11


def test_fdm_eigenpairs_match_scipy():
    """nsb_fdm_matrices (host-only; Cholesky + cyclic Jacobi in the library): the generalised eigenpairs of the 1-D
    operators of the element-wise pressure solves vs scipy.linalg.eigh on the oracle's matrices."""
    import scipy.linalg as sla
    import nekstab_next_b200 as nb
    from oracle import sem as osem
    for N in (3, 4, 5, 7, 9, 11):
        S, lam = nb.fdm_matrices(N)
        z1, w1 = osem.gll(N)
        z2, w2 = osem.gl(N - 1)
        I12 = osem.interp_matrix(z1, z2)
        D12 = I12 @ osem.dgll(N)
        b = w1.copy()
        b[0] *= 2.0
        b[-1] *= 2.0
        Eh = ((w2[:, None] * D12) / b) @ (w2[:, None] * D12).T
        Mh = ((w2[:, None] * I12) / b) @ (w2[:, None] * I12).T
        lo, _ = sla.eigh(Eh, Mh)
        assert np.max(np.abs(np.sort(lam) - lo)) <= 1e-11 * np.max(np.abs(lo))
        assert np.max(np.abs(S.T @ Mh @ S - np.eye(N - 1))) <= 1e-11
        assert np.max(np.abs(S.T @ Eh @ S - np.diag(lam))) <= 1e-11 * np.max(np.abs(lo))
