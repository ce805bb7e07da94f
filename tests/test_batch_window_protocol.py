"""The batch-window protocol of svsb_batch_peer (csrc/engine.cu "BXchg", csrc/batch.cu batch_publish / wait_flags_kernel,
DESIGN.md section 6c) as an interleaving model, checked by random and exhaustive scheduling -- CPU evidence for the slot-reuse
argument (compute-sanitizer's racecheck is not available on the GPU pool).

Model.  Every rank runs, per batch j = 1, 2, ... (sequence number j, window slot j % SLOTS), strictly in order (one stream):
  A   store its sample maxima of batch j into EVERY rank's window (one step per destination), then release-store j into
      every rank's phase-A flag for (slot, this rank);
  [M(j-1) if that batch's merge was deferred: see M]
  wA  wait until its own window's phase-A flags of the slot are all >= j; read all ranks' maxima of the slot (one step per
      source): each must carry tag j                                    -- wait_flags_kernel + union_threshold_kernel
  B   store its records of batch j into every rank's window, then release-store j into the phase-B flags;
  M   wait until its own phase-B flags of the slot are all >= j; read all ranks' records: each must carry tag j
      -- either right away, or (pipelined form) behind phase A of the next batch / at the final flush.
A step of a rank is enabled when its wait condition holds; the scheduler picks any enabled rank.  A violation is a read that
sees another batch's data -- i.e. a peer overwrote a region before this rank was done with it."""
import random

import pytest


def _program(rank, world, batches, slots, deferred):
    """The rank's step list: tuples (op, ...)."""
    prog = []

    def merge(j):
        s = j % slots
        prog.append(("wait", "B", s, j))
        for src in range(world):
            prog.append(("read", "recs", s, src, j))

    for j in range(1, batches + 1):
        s = j % slots
        for dst in range(world):
            prog.append(("store", "tops", dst, s, j))
        for dst in range(world):
            prog.append(("flag", "A", dst, s, j))
        if deferred and j > 1:
            merge(j - 1)
        prog.append(("wait", "A", s, j))
        for src in range(world):
            prog.append(("read", "tops", s, src, j))
        for dst in range(world):
            prog.append(("store", "recs", dst, s, j))
        for dst in range(world):
            prog.append(("flag", "B", dst, s, j))
        if not deferred:
            merge(j)
    if deferred:
        merge(batches)
    return prog


class _World:
    def __init__(self, world, batches, slots, deferred):
        self.world, self.slots = world, slots
        self.prog = [_program(r, world, batches, slots, deferred) for r in range(world)]
        self.pc = [0] * world
        # window[owner][kind][slot][src] = tag of the batch whose data sits there; flags[owner][phase][slot][src]
        self.win = [{k: [[0] * world for _ in range(slots)] for k in ("tops", "recs")} for _ in range(world)]
        self.flags = [{p: [[0] * world for _ in range(slots)] for p in ("A", "B")} for _ in range(world)]

    def enabled(self, r):
        if self.pc[r] >= len(self.prog[r]):
            return False
        st = self.prog[r][self.pc[r]]
        if st[0] == "wait":
            _, phase, s, j = st
            return all(f >= j for f in self.flags[r][phase][s])
        return True

    def step(self, r):
        st = self.prog[r][self.pc[r]]
        self.pc[r] += 1
        if st[0] == "store":
            _, kind, dst, s, j = st
            self.win[dst][kind][s][r] = j
        elif st[0] == "flag":
            _, phase, dst, s, j = st
            self.flags[dst][phase][s][r] = j
        elif st[0] == "read":
            _, kind, s, src, j = st
            got = self.win[r][kind][s][src]
            if got != j:
                return f"rank {r} read {kind} of rank {src} in slot {s}: batch {got}, expected {j}"
        return None

    def done(self):
        return all(self.pc[r] >= len(self.prog[r]) for r in range(self.world))

    def key(self):
        return tuple(self.pc)


def _random_runs(world, batches, slots, deferred, trials, seed):
    rng = random.Random(seed)
    for _ in range(trials):
        w = _World(world, batches, slots, deferred)
        # a biased scheduler: one rank is "fast" (runs ahead as far as the protocol lets it), which is the dangerous shape
        fast = rng.randrange(world)
        while not w.done():
            en = [r for r in range(world) if w.enabled(r)]
            assert en, "deadlock: no rank can make progress"
            r = fast if fast in en and rng.random() < 0.7 else rng.choice(en)
            err = w.step(r)
            if err:
                return err
    return None


def _exhaustive(world, batches, slots, deferred):
    """Every interleaving, memoised on the program counters (the window state is a function of them: each location's
    last writer is fixed by the counters because every rank's stores are in program order)."""
    seen = set()
    stack = [[0] * world]
    while stack:
        pcs = stack.pop()
        if tuple(pcs) in seen:
            continue
        seen.add(tuple(pcs))
        # rebuild the state by replaying each rank's prefix (stores / flags commute across ranks: distinct locations)
        w = _World(world, batches, slots, deferred)
        for r in range(world):
            for i in range(pcs[r]):
                st = w.prog[r][i]
                if st[0] == "store":
                    w.win[st[2]][st[1]][st[3]][r] = st[4]
                elif st[0] == "flag":
                    w.flags[st[2]][st[1]][st[3]][r] = st[4]
            w.pc[r] = pcs[r]
        if w.done():
            continue
        en = [r for r in range(world) if w.enabled(r)]
        assert en, f"deadlock at {pcs}"
        for r in en:
            st = w.prog[r][w.pc[r]]
            if st[0] == "read":
                got = w.win[r][st[1]][st[2]][st[3]]
                if got != st[4]:
                    return f"rank {r} read {st[1]} of rank {st[3]} in slot {st[2]}: batch {got}, expected {st[4]}"
            nxt = list(pcs)
            nxt[r] += 1
            stack.append(nxt)
    return None


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_slots_never_expose_another_batchs_data(world, deferred):
    assert _random_runs(world, batches=6, slots=2, deferred=deferred, trials=300 if world == 8 else 1500, seed=world) is None


@pytest.mark.parametrize("deferred", [False, True])
def test_two_slots_every_interleaving_of_two_ranks(deferred):
    assert _exhaustive(2, batches=3, slots=2, deferred=deferred) is None


def test_one_slot_is_enough_without_the_pipelined_merge_but_not_with_it():
    """The model must be able to FIND a violation: with a single slot the immediate form is still safe (the argument of DESIGN
    6c: a rank can only start batch j+1 after its merge of batch j saw every peer's records, which they pushed after reading all
    maxima of batch j), the pipelined form is not (phase A of batch j+1 overwrites maxima a slower peer has yet to read)."""
    assert _exhaustive(2, batches=3, slots=1, deferred=False) is None
    assert _random_runs(3, batches=5, slots=1, deferred=False, trials=1500, seed=5) is None
    assert _exhaustive(2, batches=3, slots=1, deferred=True) is not None
