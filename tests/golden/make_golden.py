"""Generate the committed golden fixtures from the reference's own field files.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Only mesh coordinates, the base-flow velocity and its pressure (as the file holds it: on the velocity mesh,
where Nek's output routine interpolates it to) are extracted (data, not source code); the known
answers are recomputed by the oracle and cross-checked against SURVEY.md section 8c.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import nekfld, sem  # noqa: E402

REF = Path('/root/reference/examples')
CASES = {
    'cyl': (REF / 'cylinder/BF_1cyl0.f00001', 2111.214601758201, 2129.932531914025, 71856, 50420),
    'bfs': (REF / 'back_fstep/baseflow/BF_bfs0.f00001', 110.0, 21.904316009768, 60120, 42341),
}

known = {}
for name, (path, sum_bm1, uu, nloc, nuniq) in CASES.items():
    f = nekfld.read_fld(path)
    x, y = f['x']
    u, v = f['u']
    n = f['nx'] - 1
    geo = sem.geometry(n, x, y)
    glo = sem.glo_num_from_coords((x, y))
    got = dict(sum_bm1=float(geo['bm1'].sum()),
               uu=sem.glsc3(u, u, geo['bm1']) + sem.glsc3(v, v, geo['bm1']),
               nlocal=int(x.size), nunique=int(glo.max()) + 1, N=n, nel=int(f['nel']),
               jac_min=float(geo['jac'].min()), jac_max=float(geo['jac'].max()))
    assert abs(got['sum_bm1'] - sum_bm1) < 1e-9 * sum_bm1, (name, got)
    assert abs(got['uu'] - uu) < 1e-9 * uu, (name, got)
    assert got['nlocal'] == nloc and got['nunique'] == nuniq, (name, got)
    known[name] = got
    np.savez_compressed(Path(__file__).parent / f'{name}_mesh.npz', x=x, y=y, u=u, v=v, p=f['p'],
                        glo=glo.astype(np.int32))
    print(name, got)
(Path(__file__).parent / 'known_answers.json').write_text(json.dumps(known, indent=1))

# boundary-condition tables of the same cases from their .re2 files (element, side, 5 parameters, type), with the
# element order of the field file the mesh fixture was taken from: what dirichlet_mask builds v1mask from
from nekstab_next_b200 import mesh  # noqa: E402

RE2 = {'cyl': REF / 'cylinder/1cyl.re2', 'bfs': REF / 'back_fstep/baseflow/bfs.re2'}
for name, path in RE2.items():
    r = mesh.read_re2(path)
    bcs = r['bcs'][0]
    elmap = nekfld.read_fld(CASES[name][0])['elmap']
    np.savez_compressed(Path(__file__).parent / f'{name}_bc.npz', nel=r['nel'], elmap=elmap.astype(np.int32),
                        elem=np.array([b[0] for b in bcs], dtype=np.int32),
                        side=np.array([b[1] for b in bcs], dtype=np.int8),
                        params=np.array([b[2] for b in bcs]), type=np.array([b[3] for b in bcs]))
    print(name, 'boundary faces', len(bcs))
