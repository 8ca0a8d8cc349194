"""There is no Fortran compiler in the image, so the ISO_C_BINDING layer cannot be compiled here.  This test
checks what a compiler would NOT even catch -- that every bind(C) interface matches the C prototype it binds
(argument count, by-value vs by-reference, C kind of every by-value scalar) -- plus the cheap syntax invariants
(balanced blocks, free-form line length)."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
F90 = ROOT / 'nekstab_next_b200' / 'fortran'
HDR = (ROOT / 'include' / 'nekstab_b200.h').read_text()

BYVAL_KIND = {'int': 'integer(c_int)', 'double': 'real(c_double)', 'int64_t': 'integer(c_int64_t)',
              'uint64_t': 'integer(c_int64_t)'}


def c_prototypes():
    text = re.sub(r'/\*.*?\*/', ' ', HDR, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(?:int|int64_t|const char \*)\s*(nsb_\w+)\s*\(([^;{]*)\)\s*;', text):
        name, args = m.group(1), ' '.join(m.group(2).split())
        params = [] if args in ('', 'void') else [a.strip() for a in args.split(',')]
        out[name] = params
    return out


def classify_c(param):
    """-> ('ref', None) for pointers, ('val', fortran type) for by-value arguments"""
    if '*' in param:
        return 'ref', None
    toks = param.replace('const', '').split()
    ctype = ' '.join(toks[:-1])
    if re.fullmatch(r'nsb_\w+_t', ctype):
        return 'val', 'type(c_ptr)'
    if ctype == 'nsb_host_matvec_fn':
        return 'val', 'type(c_funptr)'
    return 'val', BYVAL_KIND[ctype]


def fortran_interfaces(src):
    lines, buf = [], ''
    for raw in src.splitlines():
        line = raw.split('!')[0].rstrip() if "'" not in raw.split('!')[0] or raw.count("'") % 2 == 0 else raw.rstrip()
        line = raw if "name='" in raw else line
        line = line.split('!<')[0].rstrip()
        if line.rstrip().endswith('&'):
            buf += line.rstrip()[:-1] + ' '
            continue
        lines.append(buf + line)
        buf = ''
    out, i = {}, 0
    while i < len(lines):
        m = re.search(r"function\s+(\w+)\s*\(([^)]*)\)\s*.*bind\(C,\s*name='(\w+)'\)", lines[i])
        if not m:
            i += 1
            continue
        fname, args, cname = m.group(1), [a.strip() for a in m.group(2).split(',') if a.strip()], m.group(3)
        decl = {}
        i += 1
        while not re.match(r'\s*end function', lines[i]):
            d = re.match(r'\s*([\w() ,=*:]+?)\s*::\s*(.+)$', lines[i])
            if d and not lines[i].strip().startswith('import'):
                spec = d.group(1)
                base = spec.split(',')[0].strip().replace(' ', '')
                val = bool(re.search(r',\s*value\b', spec))
                for nm in re.split(r',\s*(?![^()]*\))', d.group(2)):
                    decl[re.sub(r'\(.*', '', nm).strip()] = (base, val)
            i += 1
        out[cname] = (fname, args, decl)
        i += 1
    return out


def test_every_binding_matches_its_c_prototype():
    protos = c_prototypes()
    assert len(protos) > 80
    src = (F90 / 'nekstab_b200.f90').read_text()
    ifs = fortran_interfaces(src)
    assert len(ifs) >= 60
    for cname, (fname, args, decl) in ifs.items():
        assert cname in protos, f'{cname} is bound but not declared in the header'
        assert fname == cname
        cpar = protos[cname]
        assert len(args) == len(cpar), f'{cname}: {len(args)} Fortran dummies, {len(cpar)} C parameters'
        for a, cp in zip(args, cpar):
            assert a in decl, f'{cname}: dummy {a} has no declaration'
            base, val = decl[a]
            kind, ftype = classify_c(cp)
            if kind == 'val':
                assert val, f'{cname}: `{cp}` is passed by value in C, dummy {a} lacks VALUE'
                assert base == ftype.replace(' ', ''), f'{cname}: `{cp}` needs {ftype}, dummy {a} is {base}'
            else:
                # a C pointer: by-reference dummy of any type, or a c_ptr / c_funptr handed over by value
                assert (not val) or base in ('type(c_ptr)', 'type(c_funptr)'), \
                    f'{cname}: `{cp}` is a pointer, dummy {a} is a by-value {base}'
        if cname != 'nsb_last_error' and cname != 'nsb_sem_npts':
            assert decl.get('ierr', ('', False))[0] == 'integer(c_int)', f'{cname}: result is not integer(c_int)'


def test_free_form_invariants():
    for f in F90.glob('*.f90'):
        src = f.read_text()
        for n, line in enumerate(src.splitlines(), 1):
            assert len(line) <= 132, f'{f.name}:{n} is {len(line)} characters long'
        code = '\n'.join(ln.split('!')[0] for ln in src.splitlines()).lower()
        for kw in ('function', 'subroutine', 'interface', 'module', 'type'):
            opens = len(re.findall(rf'^\s*(?:[\w()]+\s+)*{kw}\s+\w+', code, flags=re.M)) if kw != 'interface' \
                else len(re.findall(r'^\s*interface(?:\s+\w+)?\s*$', code, flags=re.M))
            ends = len(re.findall(rf'^\s*end\s+{kw}\b', code, flags=re.M))
            if kw == 'type':       # `type(c_ptr) :: x` declarations are not blocks
                opens = len(re.findall(r'^\s*type\s*(?:,[^:\n]*)?(?:::)?\s*\w+\s*$', code, flags=re.M))
            if kw == 'module':
                opens = len(re.findall(r'^\s*module\s+(?!procedure)\w+\s*$', code, flags=re.M))
            if kw == 'function':
                opens = len(re.findall(r'^\s*(?:[\w()]+\s+)*function\s+\w+\s*\(', code, flags=re.M))
            if kw == 'subroutine':
                opens = len(re.findall(r'^\s*subroutine\s+\w+', code, flags=re.M))
            assert opens == ends, f'{f.name}: {opens} `{kw}` blocks opened, {ends} closed'


def test_adapter_mirrors_the_reference_type_bound_names():
    src = (F90 / 'nekstab_b200_lightkrylov.f90').read_text()
    for proc in ('zero', 'dot', 'scal', 'axpby'):           # core/nek_vectors.f90:27-30
        assert re.search(rf'procedure, pass\(self\), public :: {proc} =>', src)
    for proc in ('matvec', 'rmatvec'):                       # core/linear_operators.f90:21-22
        assert re.search(rf'procedure, pass\(self\), public :: {proc} =>', src)
    assert 'extends(abstract_vector)' in src and 'extends(abstract_linop)' in src


def test_solver_entry_points_keep_the_reference_signatures():
    """The routines the reference's own call sites use (core/eigensolvers.f90:297,318; core/newton_krylov.f90:125,
    252; core/krylov_decomposition.f90:78) exist with the reference's names and argument lists, so those call
    sites compile unchanged against the module."""
    src = (F90 / 'nekstab_b200.f90').read_text().lower()
    ref = Path('/root/reference/core')
    want = {'arnoldi_factorization': 'q, h, mstart, mend, ksize',          # core/krylov_decomposition.f90:2
            'update_hessenberg_matrix': 'h, f, q, k',                        # core/krylov_decomposition.f90:103
            'schur_condensation': 'mstart, h, q, ksize',                     # core/eigensolvers.f90:363
            'ts_gmres': 'rhs, sol, maxiter, ksize, calls',                   # core/newton_krylov.f90:170
            'k_matmul_q': 'dq, q, yvec, k'}                                  # core/krylov_subspace.f90:163
    for name, args in want.items():
        m = re.search(rf'subroutine\s+{name}\s*\(([^)]*)\)', src)
        assert m, f'{name} is missing'
        assert [a.strip() for a in m.group(1).split(',')] == [a.strip() for a in args.split(',')], name
    assert re.search(r'interface\s+k_matmul\s+module procedure\s+k_matmul_q', src)
    if ref.exists():       # in the build container: cross-check against the reference text itself
        for fname, name, args in (('krylov_decomposition.f90', 'arnoldi_factorization', want['arnoldi_factorization']),
                                  ('krylov_decomposition.f90', 'update_hessenberg_matrix', want['update_hessenberg_matrix']),
                                  ('eigensolvers.f90', 'schur_condensation', want['schur_condensation']),
                                  ('newton_krylov.f90', 'ts_gmres', want['ts_gmres']),
                                  ('krylov_subspace.f90', 'k_matmul', want['k_matmul_q'])):
            txt = (ref / fname).read_text().lower()
            m = re.search(rf'subroutine\s+{name}\s*\(([^)]*)\)', txt)
            assert m and [a.strip() for a in m.group(1).split(',')] == [a.strip() for a in args.split(',')], name


def _joined_code(src):
    """Free-form source with comments stripped and continuation lines joined."""
    out, buf = [], ''
    for raw in src.splitlines():
        line, quote = '', None
        for ch in raw:                       # drop `!` comments outside character literals
            if quote:
                quote = None if ch == quote else quote
            elif ch in '\'"':
                quote = ch
            elif ch == '!':
                break
            line += ch
        line = line.rstrip()
        if buf and line.lstrip().startswith('&'):
            line = line.lstrip()[1:]
        if line.endswith('&'):
            buf += line[:-1] + ' '
            continue
        out.append(buf + line)
        buf = ''
    return out


def _call_args(text, start):
    """Top-level arguments of the parenthesised list opening at text[start] == '('."""
    depth, args, cur, quote = 0, [], '', None
    for i in range(start, len(text)):
        ch = text[i]
        if quote:
            cur += ch
            quote = None if ch == quote else quote
            continue
        if ch in '\'"':
            quote = ch
            cur += ch
        elif ch in '([':
            depth += 1
            cur += ch if depth > 1 else ''
        elif ch in ')]':
            depth -= 1
            if depth == 0:
                if cur.strip():
                    args.append(cur.strip())
                return args
            cur += ch
        elif ch == ',' and depth == 1:
            args.append(cur.strip())
            cur = ''
        else:
            cur += ch
    raise AssertionError('unbalanced parentheses')


def test_call_sites_pass_as_many_arguments_as_the_interfaces_take():
    """What the missing compiler would reject first: every reference to a bind(C) function inside the module's and the
    adapter's procedures hands over exactly the dummies its interface block declares."""
    src = (F90 / 'nekstab_b200.f90').read_text()
    arity = {c: len(a) for c, (_, a, _) in fortran_interfaces(src).items()}
    checked = 0
    for f in F90.glob('*.f90'):
        lines = _joined_code(f.read_text())
        inside_interface = False
        for ln in lines:
            low = ln.strip().lower()
            if re.match(r'interface\b', low):
                inside_interface = True
            elif re.match(r'end\s+interface\b', low):
                inside_interface = False
            if inside_interface or re.match(r'(public|private)\b', low):
                continue
            for m in re.finditer(r'\b(nsb_\w+)\s*\(', ln):
                name = m.group(1)
                if name not in arity:
                    continue                                   # module procedures such as nsb_check(ierr, where)
                got = _call_args(ln, m.end() - 1)
                assert len(got) == arity[name], f'{f.name}: `{ln.strip()}` passes {len(got)} arguments, ' \
                                                f'{name} takes {arity[name]}'
                checked += 1
    assert checked >= 40


def test_an_independent_parser_reads_the_module():
    """numpy.f2py's Fortran front end (the parser f2py builds wrappers from) reads the module: the same bind(C)
    functions with the same dummy lists as the regular-expression reader above finds, by-value attributes included."""
    from numpy.f2py import crackfortran
    import contextlib
    import io
    import os
    crackfortran.verbose = 0
    cwd = os.getcwd()
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            tree = crackfortran.crackfortran([str(F90 / 'nekstab_b200.f90')])
        finally:
            os.chdir(cwd)
    assert len(tree) == 1 and tree[0]['block'] == 'module' and tree[0]['name'] == 'nekstab_b200'
    found = {}
    for blk in tree[0]['body']:
        if blk['block'] != 'interface':
            continue
        for fn in blk['body']:
            if fn['block'] == 'function' and fn['name'].startswith('nsb_'):
                found[fn['name']] = fn
    ifs = fortran_interfaces((F90 / 'nekstab_b200.f90').read_text())
    assert set(found) == set(ifs)
    protos = c_prototypes()
    rename = lambda a: a[:-3] if a.endswith('_bn') else a          # f2py renames dummies that shadow intrinsics
    for name, fn in found.items():
        _, args, decl = ifs[name]
        assert [rename(a) for a in fn['args']] == [a.lower() for a in args], name
        assert len(fn['args']) == len(protos[name]), name
        for a, cp in zip(fn['args'], protos[name]):
            var = fn['vars'][a]
            by_value = 'value' in var.get('attrspec', [])
            if '*' not in cp:
                assert by_value, f'{name}: {a} must be passed by value'
                ctype = ' '.join(cp.replace('const', '').split()[:-1])
                want = {'int': ('integer', 'c_int'), 'double': ('real', 'c_double'), 'int64_t': ('integer', 'c_int64_t'),
                        'uint64_t': ('integer', 'c_int64_t')}.get(ctype)
                if want:
                    assert var['typespec'] == want[0] and var['kindselector']['kind'] == want[1], (name, a, var)
                else:
                    assert var['typespec'] == 'type' and var['typename'] in ('c_ptr', 'c_funptr'), (name, a, var)
