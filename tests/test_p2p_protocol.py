"""Model check of the NVLink peer-memory all-reduce (nsb_comm.cu, p2p_allreduce_kernel): collective `seq` uses
slot seq & 1 of every mailbox; a rank stores its vector into slot [seq & 1][rank] of EVERY mailbox, publishes
`seq` in the matching flag, waits until all flags of that slot in its own mailbox equal `seq`, then sums its own
mailbox in rank order.  Claim: two slots suffice, because a rank can be at most one collective ahead of the
slowest one.  The test replays P ranks with random delays between every remote store and checks that every rank
gets the exact sum of every collective -- and that ONE slot is not enough."""
import random


def run(P, ncoll, nslots, seed):
    rnd = random.Random(seed)
    data = [[[None] * P for _ in range(nslots)] for _ in range(P)]    # data[mailbox][slot][source rank]
    flag = [[[0] * P for _ in range(nslots)] for _ in range(P)]
    contrib = [[rnd.randrange(1, 1000) for _ in range(P)] for _ in range(ncoll + 1)]   # contrib[seq][rank]
    # per-rank program counter: (seq, phase, index); phases: 0 store data to peer i, 1 store flag to peer i,
    # 2 wait for the flags, 3 read and sum
    pc = [[1, 0, 0] for _ in range(P)]
    results = [[None] * (ncoll + 1) for _ in range(P)]
    for _ in range(200000):
        live = [r for r in range(P) if pc[r][0] <= ncoll]
        if not live:
            break
        r = rnd.choice(live)                       # any interleaving of the ranks' steps
        seq, ph, i = pc[r]
        s = seq % nslots
        if ph == 0:
            data[i][s][r] = (seq, contrib[seq][r])
            pc[r] = [seq, 0, i + 1] if i + 1 < P else [seq, 1, 0]
        elif ph == 1:
            flag[i][s][r] = seq
            pc[r] = [seq, 1, i + 1] if i + 1 < P else [seq, 2, 0]
        elif ph == 2:
            if all(flag[r][s][q] == seq for q in range(P)):
                pc[r] = [seq, 3, 0]
        else:
            got = [data[r][s][q] for q in range(P)]
            if any(g is None or g[0] != seq for g in got):
                return f'rank {r} read a slot that holds another collective at seq {seq}'
            results[r][seq] = sum(g[1] for g in got)
            pc[r] = [seq + 1, 0, 0]
    else:
        return 'deadlock'
    for seq in range(1, ncoll + 1):
        want = sum(contrib[seq])
        if any(results[r][seq] != want for r in range(P)):
            return f'wrong sum at collective {seq}'
    return 'ok'


def test_two_slots_suffice():
    for P in (2, 3, 8):
        for seed in range(40):
            assert run(P, 12, 2, seed) == 'ok'


def test_one_slot_is_not_enough():
    outcomes = [run(4, 12, 1, seed) for seed in range(40)]
    assert any(o != 'ok' for o in outcomes)


def run_halo(P, nex, nslots, seed):
    """Pairwise halo exchange of a 1-D chain of ranks (nsb_comm.cu, halo_exchange_p2p): per exchange `seq` a rank
    packs its interface data into slot seq % nslots of each neighbour's mailbox, raises the matching flag there,
    then -- one neighbour after the other -- waits for that neighbour's flag in its own mailbox and reads.
    No global synchronisation between exchanges (finish_assembled runs two dssums back to back)."""
    rnd = random.Random(seed)
    nb = [[q for q in (r - 1, r + 1) if 0 <= q < P] for r in range(P)]
    data = [[[None] * P for _ in range(nslots)] for _ in range(P)]
    flag = [[[0] * P for _ in range(nslots)] for _ in range(P)]
    pc = [[1, 0, 0] for _ in range(P)]          # seq, phase (0 pack, 1 flags, 2 wait+read), neighbour index
    for _ in range(400000):
        live = [r for r in range(P) if pc[r][0] <= nex]
        if not live:
            return 'ok'
        r = rnd.choice(live)
        seq, ph, i = pc[r]
        s = seq % nslots
        if ph == 0:
            data[nb[r][i]][s][r] = seq
            pc[r] = [seq, 0, i + 1] if i + 1 < len(nb[r]) else [seq, 1, 0]
        elif ph == 1:
            flag[nb[r][i]][s][r] = seq
            pc[r] = [seq, 1, i + 1] if i + 1 < len(nb[r]) else [seq, 2, 0]
        else:
            q = nb[r][i]
            if flag[r][s][q] == seq:
                if data[r][s][q] != seq:
                    return f'rank {r} read exchange {data[r][s][q]} from {q} while expecting {seq}'
                pc[r] = [seq, 2, i + 1] if i + 1 < len(nb[r]) else [seq + 1, 0, 0]
    return 'deadlock'


def test_halo_exchange_two_slots():
    for P in (2, 3, 5, 8):
        for seed in range(30):
            assert run_halo(P, 10, 2, seed) == 'ok'
    assert any(run_halo(4, 10, 1, seed) != 'ok' for seed in range(30))
