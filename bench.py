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
def cpu_arnoldi_rate(a, sample_ks=(10, 50, 90), verbose=False):
    """Arnoldi steps/s of the reference algorithm (MGS2 sweeps + ax/dssum matvec) on the host cores.

    Sample: Arnoldi steps at Krylov index k in sample_ks on a cpu_nelx^3 mesh (same N, ncomp);
    the step time is linear in k (MGS2 reads 20 k n words) and proportional to the number of
    elements, so the mean over k = 1..k_dim on the full mesh is the least-squares line evaluated at
    (k_dim+1)/2, scaled by E / E_sample.
    """
    from oracle import cref, sem as osem
    nes, N, nc = a.cpu_nelx, a.order, a.ncomp
    lx = N + 1
    x, y, z, glo = osem.box_mesh(nes, nes, nes, N, deform=a.deform)
    geo = osem.geometry(N, x, y, z)
    g = np.ascontiguousarray(geo['g'])
    bm1 = np.ascontiguousarray(geo['bm1'])
    off, idx = cref.gs_lists(glo)
    binv = 1.0 / osem.dssum(bm1, glo)
    x0, y0, z0, _ = osem.box_mesh(nes, nes, nes, N)
    mask = osem.boundary_mask_box(None, x0, y0, z0)
    D = osem.dgll(N)
    npts = x.size
    n = nc * npts
    kmax = max(sample_ks)
    Q = np.empty((kmax + 2, n))
    rng = np.random.default_rng(0)
    base = rng.standard_normal(1 << 16)
    for i in range(kmax + 2):                      # timing sample: content only has to be finite
        Q[i] = np.resize(np.roll(base, i), n) * 1e-3
    H = np.zeros((kmax + 2, kmax + 1), order='F')
    ts = []
    for k in sample_ks:
        t0 = time.perf_counter()
        cref.arnoldi(Q, H, k - 1, k - 1, bm1, g, bm1, binv, mask, D, lx, nes ** 3, nc, off, idx, 1.0, 0.1,
                     1.0, -1e-4)
        ts.append(time.perf_counter() - t0)
    A = np.vstack([np.ones(len(sample_ks)), np.array(sample_ks, dtype=float)]).T
    coef, *_ = np.linalg.lstsq(A, np.array(ts), rcond=None)
    mean_t = coef[0] + coef[1] * (a.kdim + 1) / 2.0
    scale = (a.nelx / nes) ** 3
    rate = 1.0 / (mean_t * scale)
    info = dict(sample=f'Arnoldi steps at k={list(sample_ks)} on box{nes}^3 (E={nes ** 3}, N={N}, ncomp={nc}); '
                       f'linear-in-k fit evaluated at k={(a.kdim + 1) / 2:g}, scaled x{scale:g} to E={a.nelx ** 3}',
                step_seconds_sample=[float(t) for t in ts], cores=cref.num_threads(),
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
        cpu_arnoldi_rate(a, sample_ks=(10,))
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
                            'CPU figure extrapolated from a bounded sample (see cpu_baseline.sample)'),
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
    # seed: random, made C0 (dssum * vmult) and masked, unit norm
    rng = np.random.default_rng(1234 + rank)
    Q[0].upload([rng.standard_normal(npts) for _ in range(nc)])
    for f in range(nc):
        sem.dssum(Q[0], f)
        sem.col2(Q[0], f, 'vmult')
        sem.col2(Q[0], f, 'mask')
    nb.k_normalize(Q[0])
    # spectral radius of L = B^-1 mask QQ^T (A + 0.1 B) by power iteration -> M = I - L / (1.05 rho)
    Lop = nb.sem_operator(sem, nc, 0.0, 1.0, 1.0, 0.1, conv=conv)
    nb.k_copy(Q[1], Q[0])
    rho = 1.0
    for _ in range(15):
        Lop.matvec(Q[1], Q[2])
        rho = nb.k_normalize(Q[2])
        nb.k_copy(Q[1], Q[2])
    Lop.close()
    op = nb.sem_operator(sem, nc, 1.0, -1.0 / (1.05 * rho), 1.0, 0.1, conv=conv)
    del m
    H = np.zeros((K + 1, K), order='F')

    def factorise():
        nb.arnoldi_factorization(Q, H, 1, K, K, op, nb.ORTH_CGS2)

    for _ in range(a.warmup):
        factorise()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(a.steps):
        factorise()
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sumall(float(ctx.launch_count() - l0))
    ms = maxall(ms)
    value = K * a.steps / (ms * 1e-3)
    ndof = sumall(float(nc * npts))

    # ---- matvec alone (GDOF/s) ---------------------------------------------------------------
    barrier()
    ctx.timer_start()
    nmv = 20
    for i in range(nmv):
        op.matvec(Q[i % K], Q[K])
    mv_ms = maxall(ctx.timer_stop()) / nmv
    matvec_gdofs = ndof / (mv_ms * 1e-3) / 1e9

    # ---- per-kernel device times for the roofline (same factorisation, events around launches) --
    Q[0].download()  # keeps column 0 intact; nothing to do, just a sync point
    ctx.prof_enable(True)
    ctx.timer_start()
    factorise()
    prof_ms = ctx.timer_stop()
    rep = ctx.prof_report()
    ctx.prof_enable(False)
    peak, peak_src = peaks()
    tot = sum(v['ms'] for v in rep.values())
    kernels = {}
    for name, v in rep.items():
        gbs = v['bytes'] / (v['ms'] * 1e-3) / 1e9 if v['ms'] > 0 else 0.0
        kernels[name] = dict(ms=round(v['ms'], 3), launches=v['launches'], share=round(v['ms'] / tot, 4),
                             achieved_gbs=round(gbs, 1), frac=round(gbs / peak, 4))
    dom = max(rep, key=lambda n: rep[n]['ms'])
    # DRAM traffic / algorithmic bytes measured once with `ncu --set full` at k = 100
    # (profiles/ncu_r01_final_kernels_k100.md): the sweeps move exactly their algorithmic bytes
    ncu_ratio = dict(multidot=1.000, fused_update_dot=1.000, update=0.999)
    traffic = (rep[dom]['bytes'] / rep[dom]['launches'] * ncu_ratio[dom]) if dom in ncu_ratio else None
    roofline = dict(bound='hbm', kernel=dom, achieved=kernels[dom]['achieved_gbs'], peak=peak, unit='GB/s',
                    frac=kernels[dom]['frac'], traffic=traffic,
                    traffic_source='ncu dram read+write / algorithmic = %.3f at k=100, applied to the mean '
                                   'algorithmic bytes per launch' % ncu_ratio[dom] if traffic else None,
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

        def host_matvec(fields, t):
            # stand-in for the reference's host time-stepper: returns host-resident vectors
            calls[0] += 1
            return bufs[calls[0] % nbuf], t

        hop = nb.host_operator(lay, host_matvec)
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
                        'reads H(:,m) back')
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
                    matvec_ms=mv_ms, roofline=roofline, clocks=clocks, gpu_launches=int(launches))
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
