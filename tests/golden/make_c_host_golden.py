"""Writes tests/golden/c_host_arnoldi.json: the oracle's (numpy restatement of core/krylov_decomposition.f90 /
core/newton_krylov.f90) results for the inputs the compiled C host tests/c_host/arnoldi_host.c builds.  CPU only:

    python tests/golden/make_c_host_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import krylov as okr  # noqa: E402

NPF, NF, K, KS = 5000, 2, 12, 15


def matvec(q):
    i = np.arange(NPF)
    d = 0.9 - 0.8 * (i / NPF)
    out = []
    for f in range(NF):
        a = q.f[f]
        out.append(d * a + 0.05 * np.roll(a, -1) - 0.03 * np.roll(a, 1) + 0.02 * q.f[1 - f])
    return okr.KVec(out, q.time)


def main():
    i = np.arange(NPF, dtype=float)
    w = 1.0 + 0.5 * np.cos(0.01 * i)
    ctx = okr.Ctx(bm1s=w, in_dot=[True, True], time_in_dot=False)
    seed = okr.KVec([np.sin(0.37 * i + 1.3 * f) + 0.25 * np.cos(0.011 * i * (f + 1)) for f in range(NF)], 0.0)
    rhs = okr.KVec([np.cos(0.21 * i - 0.7 * f) for f in range(NF)], 0.0)
    nrm = okr.k_normalize(ctx, seed)
    Q = [okr.k_zero_like(seed) for _ in range(K + 1)]
    okr.k_copy(Q[0], seed)
    H = np.zeros((K + 1, K))
    okr.arnoldi_factorization(ctx, matvec, Q, H, 1, K, K)
    sol, hist, calls = okr.ts_gmres(ctx, matvec, rhs, 4, KS, 1e-24)
    out = dict(seed_norm=nrm, H=H.tolist(), gmres_calls=calls, gmres_restarts=len(hist), sol_norm=okr.k_norm(ctx, sol),
               sol=[float(sol.f[0][0]), float(sol.f[0][NPF // 2]), float(sol.f[1][NPF - 1])], last_residual=hist[-1])
    (Path(__file__).parent / 'c_host_arnoldi.json').write_text(json.dumps(out, indent=1))
    print('seed_norm', nrm, 'calls', calls, 'restarts', len(hist), 'res', hist)


if __name__ == '__main__':
    main()
