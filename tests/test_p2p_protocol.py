"""Model check of the flag protocol of the direct ghost push (misa_md_b200/csrc/p2p.cuh + misa_b200.cu:step_pipelined):
monotonic epochs, READY receiver -> origin, ARRIVE origin -> receiver, READY posted two exchanges ahead after the force
kernel when another step follows in the same call. Ranks are run under random interleavings (every GPU's stream is
sequential, different GPUs are unordered); a ghost read must always see the value of ITS step -- never a neighbour's
next push, never a stale one -- and nothing may deadlock. CPU only: this checks the protocol, the GPU tests check the code
(tests/test_gpu_multi.py::test_two_gpus_direct_push_equals_staged_nccl_exchange)."""
import random

import pytest


class Rank:
    def __init__(self, r, n, early_on_last_step):
        self.r, self.n = r, n
        self.peers = sorted({(r - 1) % n, (r + 1) % n})      # ring: every peer is origin and destination
        self.ready = {p: 0 for p in self.peers}              # my flag words, written by peers
        self.arrive = {p: 0 for p in self.peers}
        self.ghost_x = {p: -1 for p in self.peers}            # step of the neighbour's data in my ghost shell
        self.ghost_df = {p: -1 for p in self.peers}
        self.epoch = 0
        self.ready_sent = 0
        self.early_on_last_step = early_on_last_step
        self.pc = 0
        self.prog = []

    def build(self, calls):
        """calls: steps per misa_b200_step call; a host-side reader of the ghosts (thermo / dump) runs between calls."""
        step = 0
        for n_steps in calls:
            for s in range(n_steps):
                step += 1
                last = s == n_steps - 1
                self.prog += [("push", "x", step), ("read", "x", step), ("push", "df", step), ("read", "xdf", step)]
                if not last or self.early_on_last_step:
                    self.prog.append(("post_ready", 2, step))
            self.prog.append(("read", "xdf", step))              # between calls: must still be this step's ghosts


def enabled(world, k):
    op = k.prog[k.pc]
    if op[0] == "push":
        e = k.epoch + 1
        # the in-line READY of p2p_push is sent by the same stream op, so it cannot block; the push head waits for the
        # destinations' READY >= e
        return all(k.ready[p] >= e for p in k.peers)
    if op[0] == "read":
        return all(k.arrive[p] >= k.epoch for p in k.peers)
    return True


def execute(world, k):
    op = k.prog[k.pc]
    if op[0] == "push":
        e = k.epoch + 1
        k.epoch = e
        for p in k.peers:                                       # body + tail of k_p2p_push_*
            tgt = world[p]
            (tgt.ghost_x if op[1] == "x" else tgt.ghost_df)[k.r] = op[2]
            tgt.arrive[k.r] = e
    elif op[0] == "read":
        for p in k.peers:
            assert k.ghost_x[p] == op[2], ("ghost x of step %d read in step %d" % (k.ghost_x[p], op[2]))
            if op[1] == "xdf":
                assert k.ghost_df[p] == op[2], ("ghost df of step %d read in step %d" % (k.ghost_df[p], op[2]))
    elif op[0] == "post_ready":
        send_ready(world, k, k.epoch + op[1])
    k.pc += 1


def send_ready(world, k, e):
    if k.ready_sent >= e:
        return
    k.ready_sent = e
    for p in k.peers:
        world[p].ready[k.r] = max(world[p].ready[k.r], e)


def run(n_ranks, calls, seed, early_on_last_step=False):
    rng = random.Random(seed)
    world = [Rank(r, n_ranks, early_on_last_step) for r in range(n_ranks)]
    for k in world:
        k.build(calls)
    while any(k.pc < len(k.prog) for k in world):
        # p2p_push sends its own READY in line when it was not posted ahead: model it as part of becoming runnable
        for k in world:
            if k.pc < len(k.prog) and k.prog[k.pc][0] == "push" and k.ready_sent < k.epoch + 1:
                send_ready(world, k, k.epoch + 1)
        runnable = [k for k in world if k.pc < len(k.prog) and enabled(world, k)]
        assert runnable, "deadlock"
        execute(world, rng.choice(runnable))
    return world


@pytest.mark.parametrize("n_ranks", [2, 3, 5])
def test_protocol_never_reads_a_wrong_step_and_never_deadlocks(n_ranks):
    for seed in range(300):
        run(n_ranks, calls=[3, 1, 4, 2], seed=seed)


def test_model_has_teeth_ready_ahead_on_the_last_step_of_a_call_is_a_race():
    """Posting READY ahead on the LAST step of a call lets a neighbour's next push overtake the host-side readers that run
    between calls -- the reason step_pipelined posts it only when another step follows (defer_out)."""
    with pytest.raises(AssertionError):
        for seed in range(300):
            run(3, calls=[2, 2, 2], seed=seed, early_on_last_step=True)
