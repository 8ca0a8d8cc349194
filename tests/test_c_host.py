"""The boundary is a C ABI: a C99 host program (gcc -std=c99 -pedantic -Werror) must be able to include the
header, reference every declared entry point, link against the library and get status codes back."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_c99_host_program_links_every_entry_point(lib, tmp_path):
    from nekstab_next_b200 import _capi
    hdr = (ROOT / 'include' / 'nekstab_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', ' ', hdr, flags=re.S)
    names = sorted(set(re.findall(r'\b(nsb_[a-z0-9_]+)\s*\(', text)))
    assert len(names) >= 85
    refs = '\n'.join(f'  n += (nsb_fn)(&{n}) != (nsb_fn)0;' for n in names)
    src = tmp_path / 'host.c'
    src.write_text(f'''#include "nekstab_b200.h"
#include <stdio.h>
#include <string.h>
typedef void (*nsb_fn)(void);
int main(void) {{
  int n = 0;
{refs}
  nsb_context_t ctx = 0;
  int rc = nsb_init(0, 0, 1, 0, &ctx);
  printf("%d %d %d\\n", nsb_version(), n, rc);
  if (rc == NSB_OK) return nsb_finalize(ctx);
  /* no device: a code and a message, never an abort */
  return (rc == NSB_ENODEVICE && strstr(nsb_last_error(), "no CPU fallback")) ? 0 : 1;
}}
''')
    exe = tmp_path / 'host'
    r = subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Wextra', '-Werror', '-I', str(ROOT / 'include'),
                        str(src), '-o', str(exe), '-L', str(_capi.LIB_PATH.parent), '-lnekstab_b200',
                        f'-Wl,-rpath,{_capi.LIB_PATH.parent}'], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    version, nrefs, _ = run.stdout.split()
    assert int(version) == 100 and int(nrefs) == len(names)


@pytest.mark.gpu
def test_compiled_c_host_runs_arnoldi_and_gmres(lib, tmp_path):
    """The path a Fortran caller takes, from a compiled host: tests/c_host/arnoldi_host.c (C99) runs a complete
    nsb_arnoldi and nsb_ts_gmres with a C callback operator and a dlopen'ed LAPACK, and its H / solution are
    compared with the oracle's committed golden vector (tests/golden/c_host_arnoldi.json)."""
    import glob
    import json
    import os

    import numpy as np
    import scipy
    from nekstab_next_b200 import _capi
    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), 'scipy.libs', 'libscipy_openblas*.so'))
    assert cands, 'no LAPACK shared library found (scipy.libs/libscipy_openblas*.so)'
    exe = tmp_path / 'arnoldi_host'
    r = subprocess.run(['gcc', '-std=c99', '-D_POSIX_C_SOURCE=200809L', '-O1', '-Wall', '-Wextra', '-Werror',
                        '-I', str(ROOT / 'include'), str(ROOT / 'tests' / 'c_host' / 'arnoldi_host.c'), '-o', str(exe),
                        '-L', str(_capi.LIB_PATH.parent), '-lnekstab_b200', f'-Wl,-rpath,{_capi.LIB_PATH.parent}',
                        '-ldl', '-lm'], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    run = subprocess.run([str(exe), cands[0], 'scipy_'], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    gold = json.loads((ROOT / 'tests' / 'golden' / 'c_host_arnoldi.json').read_text())
    Hg = np.array(gold['H'])
    H = np.zeros_like(Hg)
    vals = {}
    for line in run.stdout.splitlines():
        t = line.split()
        if t[0] == 'H':
            H[int(t[1]), int(t[2])] = float(t[3])
        else:
            vals[t[0]] = [float(x) for x in t[1:]]
    assert abs(vals['seed_norm'][0] - gold['seed_norm']) <= 1e-12 * gold['seed_norm']
    assert np.max(np.abs(H - Hg)) <= 1e-10 * np.max(np.abs(Hg))
    assert vals['matvec_calls_arnoldi'][0] == Hg.shape[1]
    assert vals['gmres_calls'][0] == gold['gmres_calls'] and vals['gmres_restarts'][0] == gold['gmres_restarts']
    assert abs(vals['sol_norm'][0] - gold['sol_norm']) <= 1e-8 * gold['sol_norm']
    assert np.max(np.abs(np.array(vals['sol']) - np.array(gold['sol']))) <= 1e-8 * gold['sol_norm']
