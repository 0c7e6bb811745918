"""Model check of the schedule planned for the ON-CHIP pair-symmetric passes: one persistent kernel,
work items handed out in z-plane order from a global counter, A(z) writes the pair scalars of plane z into ring slot z mod R,
B(z) gathers from planes z-2 .. z, completion counters per plane, every wait pointing BACKWARDS in the queue. Under random
interleavings of W resident CTAs: B never reads a slot that is incomplete or already reused, nothing deadlocks -- provided
the ring holds R >= lag + 3 planes; with a smaller ring and few CTAs the model deadlocks, which is the bound the kernel must
respect. CPU only, no product code involved: this pins the design before the kernel exists."""
import random

import pytest

LO = -2   # ghost planes below the box run pass A too (their upper neighbours are owned), eam_sym.cuh sym_lo


def queue(n_planes, ua, ub, lag):
    """A(LO) .. A(lag-1), then [A(z+lag), B(z)] for z = 0 .. n-1; each plane item split into units."""
    q = []
    for z in range(LO, min(lag, n_planes)):
        q += [("A", z, u) for u in range(ua)]
    for z in range(n_planes):
        if lag <= z + lag < n_planes:
            q += [("A", z + lag, u) for u in range(ua)]
        q += [("B", z, u) for u in range(ub)]
    return q


def run(n_planes, ua, ub, lag, ring, workers, seed):
    rng = random.Random(seed)
    q = queue(n_planes, ua, ub, lag)
    assert sorted(i for i in q if i[0] == "A") == [("A", z, u) for z in range(LO, n_planes) for u in range(ua)]
    done_a = {z: 0 for z in range(LO, n_planes)}
    done_b = {z: 0 for z in range(n_planes)}
    slot = [dict() for _ in range(ring)]           # slot -> {unit: plane that wrote it}
    nxt = 0
    held = [None] * workers                        # the item a CTA fetched and has not finished
    finished = 0

    def ready(item):
        kind, z, _ = item
        if kind == "A":
            p = z - ring                           # previous occupant of the slot; its consumers are B(p), B(p+1), B(p+2) --
            if p < LO:                             # ALL three: B items finish in any order (the first draft of the plan waited
                return True                        # for B(p+2) only and this model caught B(p+1) still reading)
            return all(done_b[c] == ub for c in (p, p + 1, p + 2) if 0 <= c < n_planes)
        return all(done_a[p] == ua for p in (z - 2, z - 1, z))

    while finished < len(q):
        moves = []
        for w in range(workers):
            if held[w] is None:
                if nxt < len(q):
                    moves.append(("fetch", w))
            elif ready(held[w]):
                moves.append(("exec", w))
        if not moves:
            return "deadlock"
        kind, w = rng.choice(moves)
        if kind == "fetch":
            held[w] = q[nxt]
            nxt += 1
            continue
        k, z, u = held[w]
        if k == "A":
            slot[z % ring][u] = z
            done_a[z] += 1
        else:
            for p in (z - 2, z - 1, z):            # every unit of the three planes must be there and still be theirs
                assert all(slot[p % ring].get(v) == p for v in range(ua)), ("B(%d) found slot of plane %d overwritten or incomplete" % (z, p))
            done_b[z] += 1
        held[w] = None
        finished += 1
    return "ok"


@pytest.mark.parametrize("lag", [0, 1, 3])
@pytest.mark.parametrize("workers", [1, 2, 7, 40])
def test_ring_of_lag_plus_three_planes_is_safe_and_live(lag, workers):
    for seed in range(60):
        assert run(n_planes=11, ua=5, ub=4, lag=lag, ring=lag + 3, workers=workers, seed=seed) == "ok"
        assert run(n_planes=11, ua=5, ub=4, lag=lag, ring=lag + 5, workers=workers, seed=seed) == "ok"


def test_a_smaller_ring_deadlocks_with_few_ctas():
    """A(z + R) then waits for B(p) .. B(p + 2), which sit LATER in the queue: with few CTAs every one of them ends up holding such
    an item. (It never corrupts data -- the wait is what protects the slot -- it just cannot be allowed.)"""
    outcomes = {run(n_planes=11, ua=5, ub=4, lag=3, ring=4, workers=2, seed=s) for s in range(40)}
    assert "deadlock" in outcomes
