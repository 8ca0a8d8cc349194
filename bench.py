#!/usr/bin/env python
"""Benchmark of the nekStab Arnoldi hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Workload (BASELINE.json configs[3]): 3-D box, 32^3 = 32768 spectral elements of
order N=7 (16.8 M GLL points per component), smooth deformation so the general G1..G6 path runs,
three velocity components (50.3 M dof), k_dim = 100, operator M = I - tau B^-1 mask QQ^T (A + 0.1 B).

One bench "step" is one complete k_dim-step Arnoldi factorisation (100 x {matvec, two-pass
BM1-weighted orthogonalisation, normalisation}), so that the metric does not depend on K:
  value  = Arnoldi steps/s = k_dim * K / t       (device-resident, max over ranks)
  e2e    = the same metric through the C ABI with the operator on the HOST (the reference's
           deployment: nek_advance runs on the CPU), i.e. every Arnoldi step downloads q_m, calls
           the host matvec, uploads f from pinned memory and reads H(:,m) back.
N > 1 partitions the same mesh by element slabs (strong scaling, like Nek's MPI ranks).

--impl reference times the CPU restatement of the reference's MPI Fortran path (oracle/ref_cpu.c,
OpenMP threads for ranks) -- the reference itself cannot be built here (no Fortran/MPI/Nek5000).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'arnoldi_steps_per_s'
UNIT = 'Arnoldi steps/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--kdim', type=int, default=100)
    ap.add_argument('--nelx', type=int, default=32, help='elements per direction (E = nelx^3)')
    ap.add_argument('--order', type=int, default=7)
    ap.add_argument('--ncomp', type=int, default=3)
    ap.add_argument('--deform', type=float, default=0.05)
    ap.add_argument('--conv', action='store_true', help='add the Taylor-Green convective term (M2)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-dgks', action='store_true')
    ap.add_argument('--no-c0', action='store_true', help='skip the C0 (unique-node) storage layout section')
    ap.add_argument('--no-single-rank-check', action='store_true')
    ap.add_argument('--cpu-nelx', type=int, default=16, help='elements per direction of the CPU sample')
    return ap.parse_args()


def workload_name(a):
    return (f'box{a.nelx}^3 E={a.nelx ** 3} N={a.order} deform={a.deform} ncomp={a.ncomp} '
            f'k_dim={a.kdim} op={"M2 conv" if a.conv else "M1 helmholtz"}')


def peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    power_w_max=float(max(pw)), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# CPU arm: C restatement of the reference path on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
_CPU_SETUP = {}


def cpu_setup(a, nes, kmax):
    """Mesh, geometry, gather-scatter lists and kmax+2 distinct vectors of the CPU sample (cached per size)."""
    key = (nes, a.order, a.ncomp, a.deform)
    st = _CPU_SETUP.get(key)
    if st is not None and st['kmax'] >= kmax:
        return st
    from oracle import cref, sem as osem
    N, nc = a.order, a.ncomp
    x, y, z, glo = osem.box_mesh(nes, nes, nes, N, deform=a.deform)
    geo = osem.geometry(N, x, y, z)
    g = np.ascontiguousarray(geo['g'])
    bm1 = np.ascontiguousarray(geo['bm1'])
    off, idx = cref.gs_lists(glo)
    binv = 1.0 / osem.dssum(bm1, glo)
    x0, y0, z0, _ = osem.box_mesh(nes, nes, nes, N)
    mask = osem.boundary_mask_box(None, x0, y0, z0)
    npts = x.size
    n = nc * npts
    del x, y, z, x0, y0, z0, geo, glo
    Q = np.empty((kmax + 2, n))
    base = np.random.default_rng(0).standard_normal(1 << 16)
    for i in range(kmax + 2):                      # timing sample: content only has to be finite and distinct
        Q[i] = np.resize(np.roll(base, i), n) * 1e-3
    st = dict(kmax=kmax, Q=Q, g=g, bm1=bm1, binv=binv, mask=mask, D=osem.dgll(N), off=off, idx=idx, nes=nes,
              H=np.zeros((kmax + 2, kmax + 1), order='F'))
    _CPU_SETUP.clear()
    _CPU_SETUP[key] = st
    return st


def cpu_step_seconds(a, st, k):
    """One Arnoldi step of the reference algorithm at Krylov index k: matvec (ax + dssum + mask + binvm1) and the
    two MGS sweeps of update_hessenberg_matrix over k vectors (core/krylov_decomposition.f90:155-180)."""
    from oracle import cref
    t0 = time.perf_counter()
    cref.arnoldi(st['Q'], st['H'], k - 1, k - 1, st['bm1'], st['g'], st['bm1'], st['binv'], st['mask'], st['D'],
                 a.order + 1, st['nes'] ** 3, a.ncomp, st['off'], st['idx'], 1.0, 0.1, 1.0, -1e-4)
    return time.perf_counter() - t0


def cpu_arnoldi_rate(a, sample_ks=(10, 40, 70), verbose=False):
    """Arnoldi steps/s of the reference algorithm (MGS2 sweeps + ax/dssum matvec) on the host cores.

    Preferred sample: Arnoldi steps at Krylov index k in sample_ks on the FULL mesh of the workload (the
    configuration itself, nothing scaled in the element count); the step time is linear in k (MGS2 reads
    20 k n words), so the mean over k = 1..k_dim is the least-squares line evaluated at (k_dim+1)/2.
    When the host lacks the memory for max(k)+2 full-size vectors, the k-fit runs on a cpu_nelx^3 mesh and the
    element scaling is MEASURED with one full-mesh step at k = min(sample_ks) instead of assumed.
    """
    from oracle import cref
    N, nc = a.order, a.ncomp
    n_full = nc * a.nelx ** 3 * (N + 1) ** 3
    kmax = max(sample_ks)
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:  # noqa: BLE001
        avail = 0
    need = lambda k, n: 8.0 * n * (k + 2 + 14) * 1.25          # vectors + mesh arrays + transient copies
    if avail > need(kmax, n_full) + (8 << 30):
        st = cpu_setup(a, a.nelx, kmax)
        ts = [cpu_step_seconds(a, st, k) for k in sample_ks]
        scale, how = 1.0, (f'Arnoldi steps at k={list(sample_ks)} on the full mesh box{a.nelx}^3 '
                           f'(E={a.nelx ** 3}, N={N}, ncomp={nc}); linear-in-k fit evaluated at k={(a.kdim + 1) / 2:g}')
    else:
        nes = a.cpu_nelx
        st = cpu_setup(a, nes, kmax)
        ts = [cpu_step_seconds(a, st, k) for k in sample_ks]
        k0 = min(sample_ks)
        t_small = ts[list(sample_ks).index(k0)]
        st = None
        stf = cpu_setup(a, a.nelx, k0)
        t_full = cpu_step_seconds(a, stf, k0)
        scale = t_full / t_small
        how = (f'Arnoldi steps at k={list(sample_ks)} on box{nes}^3 (E={nes ** 3}, N={N}, ncomp={nc}); linear-in-k fit '
               f'evaluated at k={(a.kdim + 1) / 2:g}; element scaling measured with one step at k={k0} on the full '
               f'box{a.nelx}^3 mesh: x{scale:.2f} (element ratio {(a.nelx / nes) ** 3:g})')
    A = np.vstack([np.ones(len(sample_ks)), np.array(sample_ks, dtype=float)]).T
    coef, *_ = np.linalg.lstsq(A, np.array(ts), rcond=None)
    mean_t = (coef[0] + coef[1] * (a.kdim + 1) / 2.0) * scale
    rate = 1.0 / mean_t
    info = dict(sample=how, step_seconds_sample=[float(t) for t in ts], cores=cref.num_threads(),
                host_cpus=os.cpu_count())
    return rate, info


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        # torchrun pins OMP_NUM_THREADS=1 for multi-rank launches; rank 0 alone works here and may use
        # every host core, like the single-process launch
        from oracle import cref
        cref.set_num_threads(os.cpu_count() or 1)
    rates, info = [], None
    for _ in range(max(1, a.warmup)):
        cpu_arnoldi_rate(a, sample_ks=(10, 70))
    t0 = time.perf_counter()
    for _ in range(a.steps):
        r, info = cpu_arnoldi_rate(a)
        rates.append(r)
    wall = time.perf_counter() - t0
    value = float(np.mean(rates))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=1e3 * a.kdim / value, higher_is_better=True, scaling='strong', vs_baseline=None,
                dtype='f64', data='synthetic', impl='reference',
                config=dict(workload=workload_name(a), note='one bench step = one k_dim-step factorisation; '
                            'CPU figure = mean over k from a bounded sample of steps (see cpu_baseline.sample)'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=info['cores'], kind='port',
                                  sample=info['sample'], host_cpus=info['host_cpus']),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, wall_s=wall)
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def pinned_array(lib, n):
    import ctypes as C
    p = C.c_void_p()
    from nekstab_next_b200._capi import check
    check(lib.nsb_host_alloc(C.byref(p), int(n) * 8))
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(int(n),)), p


def run_native(a):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit('launch with torch.distributed.run --nproc-per-node N for --gpus N > 1')
    if rank == 0:
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):   # stdout carries the one JSON line only
            ge.build()
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        dist.barrier()
    import nekstab_next_b200 as nb
    if world > 1:
        box = [nb.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    ctx = nb.Context(device=local, rank=rank, nranks=world, unique_id=uid)
    p2p = False
    if world > 1 and os.environ.get('NSB_NO_P2P', '0') != '1':
        def allgather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        p2p = ctx.connect_peers(allgather)      # NVLink peer-memory all-reduce / halo exchange

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def maxall(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumall(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    K, N, nc = a.kdim, a.order, a.ncomp
    m = nb.mesh.box_mesh(a.nelx, a.nelx, a.nelx, N, deform=a.deform, rank=rank, nranks=world)
    sem = nb.Sem(ctx, N, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
    if world > 1:
        sem.setup_exchange()
    npts = sem.npts
    bm1 = sem.get('bm1')
    lay = nb.Layout(ctx, [npts] * nc, [True] * nc)
    lay.set_weight([bm1] * nc)
    Q = nb.Basis(lay, K + 1)
    conv = nb.mesh.taylor_green(m['x'], m['y'], m['z']) if a.conv else None
    rng = np.random.default_rng(1234 + rank)

    # seed: pseudo-random in the GLOBAL node id (continuous by construction, independent of the partition,
    # so every N starts from the same vector), masked, unit norm
    def set_seed(sem_, vec, glo):
        vec.upload([nb.seed.hashed_field(glo, c) for c in range(nc)])
        for f in range(nc):
            sem_.col2(vec, f, 'mask')
        nb.k_normalize(vec)

    set_seed(sem, Q[0], m['glo'])
    # spectral radius of L = B^-1 mask QQ^T (A + 0.1 B) by power iteration -> M = I - L / (1.05 rho)
    Lop = nb.sem_operator(sem, nc, 0.0, 1.0, 1.0, 0.1, conv=conv)
    nb.k_copy(Q[1], Q[0])
    rho = 1.0
    for _ in range(15):
        Lop.matvec(Q[1], Q[2])
        rho = nb.k_normalize(Q[2])
        nb.k_copy(Q[1], Q[2])
    Lop.close()
    beta_op = -1.0 / (1.05 * rho)
    op = nb.sem_operator(sem, nc, 1.0, beta_op, 1.0, 0.1, conv=conv)
    glo_local = m['glo']
    mask_local = m['mask']
    del m
    H = np.zeros((K + 1, K), order='F')

    def factorise(mode=nb.ORTH_CGS2):
        nb.arnoldi_factorization(Q, H, 1, K, K, op, mode)

    def timed(mode, warm, steps, sample_clocks=False):
        for _ in range(warm):
            factorise(mode)
        sampler = ClockSampler(local)
        barrier()
        if rank == 0 and sample_clocks:
            sampler.start()
        l0 = ctx.launch_count()
        ctx.timer_start()
        for _ in range(steps):
            factorise(mode)
        t = ctx.timer_stop()
        barrier()
        clk = sampler.stop() if (rank == 0 and sample_clocks) else None
        return maxall(t), sumall(float(ctx.launch_count() - l0)), clk

    ms, launches, clocks = timed(nb.ORTH_CGS2, a.warmup, a.steps, sample_clocks=True)
    value = K * a.steps / (ms * 1e-3)
    ndof = sumall(float(nc * npts))

    # ---- parity block: computed on the factorisation the timed region just produced ----------------
    #   orth        max |V^T B V - I| over all k_dim+1 columns (the reference's orthonormality.dat check,
    #               core/eigensolvers.f90:335-345; north-star bound 1e-10)
    #   arnoldi_res ||M q_j - V h_j|| / ||h_j|| for j in {0, k_dim/2, k_dim-1}
    #   H_replicated  max over ranks of |H - H_rank0| (must be exactly 0: rank-ordered all-reduce sums)
    #   H8_vs_single_rank  (N > 1) relative difference of H(:, 0:8) against a 1-rank run of the same seed on
    #               the whole mesh, run on rank 0's GPU
    parity = {}
    G = Q.gram(K + 1)
    parity['orth'] = float(np.max(np.abs(G - np.eye(K + 1))))
    Wk = nb.Basis(lay, 2)
    res = []
    for j in (0, K // 2, K - 1):
        op.matvec(Q[j], Wk[0])
        nb.k_matmul(Wk[1], Q, H[:j + 2, j], j + 2)
        nb.k_sub2(Wk[0], Wk[1])
        res.append(nb.k_norm(Wk[0]) / float(np.linalg.norm(H[:j + 2, j])))
    parity['arnoldi_res'] = [float(r) for r in res]
    Wk.close()
    if world > 1:
        Hall = [None] * world
        dist.all_gather_object(Hall, H)
        parity['H_replicated'] = max(float(np.max(np.abs(h - Hall[0]))) for h in Hall)
    else:
        parity['H_replicated'] = 0.0
    if world > 1 and not a.no_single_rank_check:
        k8 = min(8, K)
        if rank == 0:
            c1 = nb.Context(device=local)
            m1 = nb.mesh.box_mesh(a.nelx, a.nelx, a.nelx, N, deform=a.deform)
            s1 = nb.Sem(c1, N, m1['x'], m1['y'], m1['z'], mask=m1['mask'], glo_num=m1['glo'])
            l1 = nb.Layout(c1, [s1.npts] * nc, [True] * nc)
            l1.set_weight([s1.get('bm1')] * nc)
            Q1 = nb.Basis(l1, k8 + 1)
            set_seed(s1, Q1[0], m1['glo'])
            conv1 = nb.mesh.taylor_green(m1['x'], m1['y'], m1['z']) if a.conv else None
            o1 = nb.sem_operator(s1, nc, 1.0, beta_op, 1.0, 0.1, conv=conv1)
            H1 = np.zeros((k8 + 1, k8), order='F')
            nb.arnoldi_factorization(Q1, H1, 1, k8, k8, o1, nb.ORTH_CGS2)
            parity['H8_vs_single_rank'] = float(np.max(np.abs(H1 - H[:k8 + 1, :k8])) / np.max(np.abs(H1)))
            for o in (o1, Q1, l1, s1, c1):
                o.close()
            del m1
        barrier()
    parity['bounds'] = dict(orth=1e-10, arnoldi_res=1e-10, H_replicated=0.0, H8_vs_single_rank=1e-10)
    parity['ok'] = bool(parity['orth'] < 1e-10 and max(parity['arnoldi_res']) < 1e-10 and
                        parity['H_replicated'] == 0.0 and parity.get('H8_vs_single_rank', 0.0) < 1e-10)

    # ---- DGKS: the same factorisation with the second projection decided on the device -------------
    dgks = None
    if not a.no_dgks:
        set_seed(sem, Q[0], glo_local)
        d_ms, _, _ = timed(nb.ORTH_DGKS, 1, a.steps)
        passes = nb.arnoldi_passes(Q, 1, K, nb.ORTH_DGKS)
        Gd = Q.gram(K + 1)
        dgks = dict(value_dgks=K * a.steps / (d_ms * 1e-3), ms_per_arnoldi_step=d_ms / a.steps / K,
                    passes_mean=float(np.mean(passes)), orth=float(np.max(np.abs(Gd - np.eye(K + 1)))),
                    note='second projection only where |w\'| < |w|/sqrt 2 (device-side predicate); the headline '
                         'value keeps the reference\'s unconditional second pass')
        set_seed(sem, Q[0], glo_local)
        factorise()                                  # basis of the headline mode again for what follows

    # ---- the same factorisation on the C0 (unique-node) storage layout ------------------------------
    # Krylov vectors are continuous, so storing a shared GLL node once instead of once per element changes no
    # result (the BM1 inner product over distinct nodes with the assembled weight is the same sum) but every
    # sweep moves 32 % fewer bytes.  Reported next to the headline, which stays on the reference's element-local
    # layout; same seed, same operator, same parity block plus the difference of H against the headline run.
    c0 = None
    if not a.no_c0:
        H_local = H.copy()
        layC = nb.Layout(ctx, [npts] * nc, [True] * nc, c0_sem=sem, n_c0=nc)
        layC.set_weight([bm1] * nc)
        QC = nb.Basis(layC, K + 1)
        HC = np.zeros((K + 1, K), order='F')

        def seed_c0():
            QC[0].upload([nb.seed.hashed_field(glo_local, c) * mask_local for c in range(nc)])
            nb.k_normalize(QC[0])

        seed_c0()
        for _ in range(a.warmup):
            nb.arnoldi_factorization(QC, HC, 1, K, K, op, nb.ORTH_CGS2)
        barrier()
        l0 = ctx.launch_count()
        ctx.timer_start()
        for _ in range(a.steps):
            nb.arnoldi_factorization(QC, HC, 1, K, K, op, nb.ORTH_CGS2)
        c_ms = maxall(ctx.timer_stop())
        barrier()
        c_launch = sumall(float(ctx.launch_count() - l0))
        GC = QC.gram(K + 1)
        WC = nb.Basis(layC, 2)
        resC = []
        for j in (0, K // 2, K - 1):
            op.matvec(QC[j], WC[0])
            nb.k_matmul(WC[1], QC, HC[:j + 2, j], j + 2)
            nb.k_sub2(WC[0], WC[1])
            resC.append(float(nb.k_norm(WC[0]) / np.linalg.norm(HC[:j + 2, j])))
        ctx.timer_start()
        for i in range(20):
            op.matvec(QC[i % K], WC[0])
        mvC = maxall(ctx.timer_stop()) / 20
        WC.close()
        ctx.prof_enable(True)
        nb.arnoldi_factorization(QC, HC, 1, K, K, op, nb.ORTH_CGS2)
        repC = ctx.prof_report()
        ctx.prof_enable(False)
        pkC, _ = peaks()
        kernC = {n: dict(ms=round(v['ms'], 3), launches=v['launches'],
                         achieved_gbs=round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1) if v['ms'] > 0 else 0.0,
                         frac=round(v['bytes'] / (v['ms'] * 1e-3) / 1e9 / pkC, 4) if v['ms'] > 0 else 0.0)
                 for n, v in repC.items()}
        rows = sumall(float(layC.c0_rows))
        c0 = dict(value=K * a.steps / (c_ms * 1e-3), arnoldi_ms_per_step=c_ms / a.steps / K, gpu_launches=int(c_launch),
                  stored_rows_per_component=int(rows), element_local_rows_per_component=int(ndof / nc),
                  matvec_ms=mvC, kernels=kernC,
                  parity=dict(orth=float(np.max(np.abs(GC - np.eye(K + 1)))), arnoldi_res=resC,
                              H_vs_element_local=float(np.max(np.abs(HC - H_local)) / np.max(np.abs(H_local))),
                              H8_vs_element_local=float(np.max(np.abs(HC[:9, :8] - H_local[:9, :8])) / np.max(np.abs(H_local)))),
                  note='nsb_layout_create_c0: one row per distinct GLL node (shared nodes are not duplicated); '
                       'host interface, operator and inner product unchanged; opt-in, for continuous fields')
        c0['parity']['ok'] = bool(c0['parity']['orth'] < 1e-10 and max(resC) < 1e-10 and
                                  c0['parity']['H8_vs_element_local'] < 1e-10)
        QC.close()
        layC.close()

    # ---- matvec alone (GDOF/s) ---------------------------------------------------------------
    barrier()
    ctx.timer_start()
    nmv = 20
    for i in range(nmv):
        op.matvec(Q[i % K], Q[K])
    mv_ms = maxall(ctx.timer_stop()) / nmv
    matvec_gdofs = ndof / (mv_ms * 1e-3) / 1e9
    # whole `ax` (axhelm + dssum + mask + binvm1) against its algorithmic minimum: u and w of every component,
    # G1..G6, bm1 and bmask once per point
    pk, _ = peaks()
    mv_alg = 8.0 * (2 * nc + 8) * (ndof / nc)
    matvec_roofline = dict(algorithmic_bytes=mv_alg, achieved_gbs=mv_alg / (mv_ms * 1e-3) / 1e9,
                           frac=mv_alg / (mv_ms * 1e-3) / 1e9 / (pk * world),
                           structure='axhelm of element slab s+1 overlapped with the L2-resident gather-scatter of slab s'
                           if os.environ.get('NSB_AX_SLAB_MB', '0') not in ('0', '0.0') else 'one axhelm + one gather-scatter launch')

    # ---- per-kernel device times for the roofline (same factorisation, events around launches) --
    Q[0].download()  # keeps column 0 intact; nothing to do, just a sync point
    ctx.prof_enable(True)
    ctx.timer_start()
    factorise()
    prof_ms = ctx.timer_stop()
    # BLAS-1 set of the nek_dvector type (K4): not part of the fused Arnoldi step, timed here for the table
    Wb = nb.Basis(lay, 3)
    nb.k_copy(Wb[0], Q[1]); nb.k_copy(Wb[1], Q[2])
    for _ in range(5):
        Wb[0].axpby(0.5, Wb[1], 0.25, skip_time=False)
        nb.k_sub3(Wb[2], Wb[0], Wb[1])
        Wb[2].scal(1.0001)
    rep = ctx.prof_report()
    ctx.prof_enable(False)
    Wb.close()
    peak, peak_src = peaks()
    tot = sum(v['ms'] for v in rep.values())
    kernels = {}
    for name, v in rep.items():
        gbs = v['bytes'] / (v['ms'] * 1e-3) / 1e9 if v['ms'] > 0 else 0.0
        kernels[name] = dict(ms=round(v['ms'], 3), launches=v['launches'], share=round(v['ms'] / tot, 4),
                             achieved_gbs=round(gbs, 1), frac=round(gbs / peak, 4))
    dom = max((n for n in rep if n != 'blas1'), key=lambda n: rep[n]['ms'])
    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
    # capture, kept as a committed file (profiles/ncu_traffic.json: per kernel the measured bytes of the captured
    # launch and its algorithmic bytes); scaled to this run's mean algorithmic bytes per launch.  null when the
    # file has no entry for the kernel -- never a hard-coded ratio.
    traffic, traffic_src = None, None
    tf = ROOT / 'profiles' / 'ncu_traffic.json'
    if tf.exists():
        ent = json.loads(tf.read_text()).get(dom)
        if ent and ent.get('algorithmic_bytes'):
            ratio = ent['dram_bytes'] / ent['algorithmic_bytes']
            traffic = rep[dom]['bytes'] / rep[dom]['launches'] * ratio
            traffic_src = (f"{ent.get('source', 'profiles/ncu_traffic.json')}: dram read+write {ent['dram_bytes']:.4g} B "
                           f"for {ent['algorithmic_bytes']:.4g} algorithmic B at k={ent.get('k')} (ratio {ratio:.3f}), "
                           'applied to the mean algorithmic bytes per launch of this run')
    roofline = dict(bound='hbm', kernel=dom, achieved=kernels[dom]['achieved_gbs'], peak=peak, unit='GB/s',
                    frac=kernels[dom]['frac'], traffic=traffic, traffic_source=traffic_src,
                    peak_source=peak_src,
                    avg_launch_ms=round(rep[dom]['ms'] / rep[dom]['launches'], 4),
                    algorithmic_bytes_per_launch=rep[dom]['bytes'] / rep[dom]['launches'],
                    share_of_step=kernels[dom]['share'], kernels=kernels,
                    profiled_pass_ms=round(prof_ms, 2))

    # ---- e2e: operator on the host, vectors cross PCIe every Arnoldi step -----------------------
    e2e = None
    if not a.no_e2e:
        nbuf = 3
        bufs = []
        for b in range(nbuf):
            fs = []
            for f in range(nc):
                arr, _p = pinned_array(ctx.lib, npts)
                arr[:] = rng.standard_normal(npts)
                fs.append(arr)
            bufs.append(fs)
        calls = [0]

        noise = rng.standard_normal((K + 8, 4096))

        def host_matvec(fields, t):
            # stand-in for the reference's host time-stepper: returns host-resident vectors.  The three pinned
            # buffers are recycled, but every call rewrites a 4096-entry window with fresh numbers, so each returned
            # vector has a component outside the span of the earlier ones (a non-degenerate Arnoldi process; the
            # first version returned the same three vectors in turn and every step from the fourth on was an exact
            # breakdown that only the measured norm of rounding noise kept going).
            calls[0] += 1
            out = bufs[calls[0] % nbuf]
            j = calls[0] % noise.shape[0]
            for f in range(nc):
                out[f][j * 4096:(j + 1) * 4096] = noise[j] * (f + 1)
            return out, t

        hop = nb.host_operator(lay, host_matvec, linear=True)   # linearised time-stepper: un-normalised hand-over allowed
        He = np.zeros((K + 1, K), order='F')
        nb.arnoldi_factorization(Q, He, 1, 3, K, hop, nb.ORTH_CGS2)        # warm-up
        barrier()
        ctx.timer_start()
        nb.arnoldi_factorization(Q, He, 1, K, K, hop, nb.ORTH_CGS2)
        e_ms = maxall(ctx.timer_stop())
        barrier()
        bytes_step = float(nc * npts * 8)
        e2e = dict(value=K / (e_ms * 1e-3), unit=UNIT,
                   h2d_bytes_per_step=sumall(bytes_step) * K, d2h_bytes_per_step=sumall(bytes_step + 8 * (K + 1)) * K,
                   note='per bench step (= k_dim Arnoldi steps): each Arnoldi step downloads q_m, calls the host '
                        'matvec stand-in, uploads f from pinned host memory, orthonormalises on the GPU and '
                        'reads H(:,m) back; the operator is declared linear (nsb_op_set_linear), so the download of '
                        'q_m+1 overlaps the third sweep of step m')
        hop.close()

    line = None
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                    ms_per_step=ms / a.steps, higher_is_better=True, scaling='strong', vs_baseline=None,
                    dtype='f64', data='synthetic',
                    config=dict(workload=workload_name(a), ndof=int(ndof), elements=a.nelx ** 3,
                                orthogonalisation='CGS2 (two fused passes, H = h1 + h2)',
                                bench_step='one k_dim-step Arnoldi factorisation',
                                l2='inputs (basis >= 0.4 GB per column) exceed the 126 MB L2; no flush needed',
                                partition=f'{world} z-slab(s) of elements',
                                collectives=('NVLink peer-memory kernels (one-shot all-reduce, halo stores)' if p2p
                                             else 'NCCL' if world > 1 else 'none')),
                    arnoldi_ms_per_step=ms / a.steps / K, matvec_gdof_per_s=matvec_gdofs,
                    matvec_ms=mv_ms, matvec_roofline=matvec_roofline, roofline=roofline, clocks=clocks, gpu_launches=int(launches),
                    parity=parity, build_id=nb.build_id(),
                    build_mode='in-tree nvcc -gencode arch=compute_100a,code=sm_100a; content hash of csrc/ + '
                               'include/ embedded in the binary (nsb_build_id) and compared by build.stale()',
                    launch_mode=os.environ.get('NSB_GRAPH', '1') != '0' and 'one CUDA graph per Arnoldi step' or 'plain launches')
        if dgks:
            line.update(value_dgks=dgks['value_dgks'], passes_mean=dgks['passes_mean'], dgks=dgks)
        if c0:
            line.update(value_c0=c0['value'], c0_layout=c0)
        line['config']['layout'] = ('headline value: element-local layout of the reference (shared nodes duplicated); '
                                    'value_c0: the same factorisation with the basis stored once per distinct node')
        if e2e:
            line['e2e'] = e2e
    if rank == 0 and world == 1 and not a.no_cpu:
        rate, info = cpu_arnoldi_rate(a)
        line['cpu_baseline'] = dict(value=rate, unit=UNIT, cores=info['cores'], kind='port',
                                    sample=info['sample'], host_cpus=info['host_cpus'])
    if rank == 0:
        print(json.dumps(line))
    op.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    args = parse()
    # stdout must carry exactly one JSON line: libraries (NCCL banner, build logs) that write to
    # fd 1 are diverted to stderr until the result is printed.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _line = []
    _print = print

    def print(*a, **k):  # noqa: A001  (the run_* functions print the JSON line last)
        _line.append(' '.join(str(x) for x in a))

    import builtins
    builtins.print = print
    try:
        rc = run_reference(args) if args.impl == 'reference' else run_native(args)
    finally:
        builtins.print = _print
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
    for ln in _line:
        if ln.startswith('{'):
            _print(ln, flush=True)
        else:
            _print(ln, file=sys.stderr, flush=True)
    sys.exit(rc)
