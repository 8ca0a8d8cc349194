"""Model check of the producer / consumer protocol of the axhelm TMA ring (nsb_sem.cu, axhelm3d_dmma8_kernel):
element j lives in buffer j % NBUF and is processed by warp group j % NG, synchronised by full[] / empty[]
mbarriers that are waited on BY PARITY.  A parity wait is ambiguous when the waiter is two phases ahead of the
barrier; with NBUF > NG consecutive uses of a buffer belong to different warps and that can happen.  The
kernel therefore waits for the previous user's release before waiting for its own data; this test replays both
variants under random timings (TMA latency, per-element compute time) and checks that the shipped one never
reads a buffer that does not hold its element, and that the naive one does (the bug seen on the GPU as launch
failures with NF = 1 on large meshes)."""
import random

import pytest


def simulate(NG, NBUF, nit, seed, guarded):
    rnd = random.Random(seed)
    full_done, empty_done = [0] * NBUF, [0] * NBUF     # completed phases of every mbarrier
    content, inflight = [None] * NBUF, [None] * NBUF
    tma = []                                           # (completion time, buffer, element)
    p_it, p_time = 0, 0.0
    g_it, g_time, g_busy = list(range(NG)), [0.0] * NG, [False] * NG
    time = 0.0
    for _ in range(400000):
        time += 1.0
        for x in [x for x in tma if x[0] <= time]:
            tma.remove(x)
            full_done[x[1]] += 1
            content[x[1]], inflight[x[1]] = x[2], None
        if p_it < nit and p_time <= time:              # producer: wait empty[s] by parity, arm, issue the copies
            s, u = p_it % NBUF, p_it // NBUF
            if (empty_done[s] & 1) != ((u & 1) ^ 1):
                if empty_done[s] != u or inflight[s] is not None or full_done[s] != u:
                    return 'producer armed a buffer that was not released'
                inflight[s] = p_it
                tma.append((time + rnd.uniform(1, 60), s, p_it))
                p_it += 1
                p_time = time + rnd.uniform(0.1, 2)
        for g in range(NG):
            if g_it[g] >= nit or g_time[g] > time:
                continue
            it = g_it[g]
            s, u = it % NBUF, it // NBUF
            if not g_busy[g]:
                if guarded and u >= 1 and (empty_done[s] & 1) == ((u - 1) & 1):
                    continue                           # mbar_wait(empty[s], (u - 1) & 1) still spinning
                if (full_done[s] & 1) != (u & 1):      # mbar_wait(full[s], u & 1) passes
                    if content[s] != it or full_done[s] != u + 1:
                        return 'consumer passed its wait on a buffer that does not hold its element'
                    g_busy[g] = True
                    g_time[g] = time + rnd.choice([rnd.uniform(1, 5), rnd.uniform(20, 200)])
            else:
                empty_done[s] += 1                     # mbar_arrive(empty[s])
                g_busy[g] = False
                g_it[g] += NG
        if p_it >= nit and all(x >= nit for x in g_it):
            return 'ok'
    return 'deadlock'


# (warp groups, buffers) of every instantiation the library ships: NF = 3 / 2 / 1 without and with convection
@pytest.mark.parametrize('NG,NBUF', [(3, 4), (4, 5), (5, 6), (3, 3), (4, 4), (5, 5)])
def test_shipped_protocol_never_aliases(NG, NBUF):
    for seed in range(60):
        assert simulate(NG, NBUF, 90, seed, guarded=True) == 'ok'


def test_parity_only_protocol_aliases_when_buffers_outnumber_groups():
    bad = [simulate(3, 4, 90, seed, guarded=False) for seed in range(20)]
    assert any(r != 'ok' for r in bad)
    # one buffer per group (the same warp owns consecutive uses) is safe without the guard
    assert all(simulate(3, 3, 90, seed, guarded=False) == 'ok' for seed in range(20))


def test_more_groups_than_buffers_is_unsafe_even_with_the_guard():
    """Why Dmma8Cfg caps the warp groups at the number of buffers that fit (the single-component kernel with
    convection fits 4 buffers, so it runs 4 groups, not 5)."""
    assert any(simulate(5, 4, 90, seed, guarded=True) != 'ok' for seed in range(20))
