"""The boundary is a C ABI: a C99 host program (gcc -std=c99 -pedantic -Werror) must be able to include the
header, reference every declared entry point, link against the library and get status codes back."""
import re
import subprocess
from pathlib import Path

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
