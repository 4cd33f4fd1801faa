"""Model check of the producer/consumer protocol of k_spmm_staged (csrc/spmm_staged.cu) on the CPU.

The kernel's warps are modelled as coroutines that follow the kernel's control flow statement by statement
(stage ring, parity toggling, full/empty mbarriers with transaction counting, asynchronous copies that land
at arbitrary later times); a random scheduler interleaves them.  Checked: no deadlock, every consumer reads
tile t's data from the stage (never a stale or overwritten tile), a producer never writes a stage a consumer
still reads.  mbarrier semantics modelled as documented in the PTX ISA: a phase completes when the pending
arrival count AND the transaction count reach zero; try_wait.parity(P) succeeds once the phase of parity P
has completed (immediately for P = 1 on a freshly initialised barrier)."""
import random

import pytest

LANES = 4          # lanes per modelled producer warp (ldgsts mode: every lane copies a piece of every row)


class MBar:
    def __init__(self, count):
        self.init, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.init

    def arrive(self, expect_tx=0):
        self.tx += expect_tx
        assert self.pending > 0, "more arrivals than the barrier was initialised for"
        self.pending -= 1
        self._check()

    def complete_tx(self, n):
        self.tx -= n
        self._check()

    def done(self, parity):            # try_wait.parity
        return self.phase != parity


def simulate(n_tiles, S, W, NP, mode, seed, rows_per_tile=5, empty_count=None):
    """Returns the number of operand-row reads performed.  A step of a coroutine yields True when it made
    progress and False when it only polled a barrier."""
    rng = random.Random(seed)
    full = [MBar(NP if mode == "bulk" else NP * LANES) for _ in range(S)]
    empty = [MBar(W if empty_count is None else empty_count) for _ in range(S)]
    stage_tile = [[None] * rows_per_tile for _ in range(S)]    # which tile each row slot of a stage holds
    readers = [0] * S                                          # consumers currently reading the stage
    pending_ops = []                                           # asynchronous operations not yet performed (closures)
    reads = [0]

    def producer(pj):
        stage, parity = 0, 0
        for t in range(n_tiles):
            while not empty[stage].done(parity ^ 1):
                yield False
            rows = [r for r in range(rows_per_tile) if r % NP == pj]
            bar = full[stage]
            if mode == "bulk":
                bar.arrive(expect_tx=len(rows) * 16)           # lane 0: arrive.expect_tx, then __syncwarp
                yield True
                for r in rows:
                    assert readers[stage] == 0, "producer overwrites a stage that is being read"

                    def land(stage=stage, r=r, t=t, bar=bar):
                        stage_tile[stage][r] = t
                        bar.complete_tx(16)
                    pending_ops.append(land)
                    yield True
            else:
                left = [len(rows)] * LANES                     # pieces of this tile each lane still has in flight
                fired = [False] * LANES
                for r in rows:
                    assert readers[stage] == 0, "producer overwrites a stage that is being read"
                    row_left = [LANES]
                    for lane in range(LANES):
                        def land(stage=stage, r=r, t=t, lane=lane, row_left=row_left, left=left, fired=fired, bar=bar):
                            row_left[0] -= 1
                            if row_left[0] == 0:
                                stage_tile[stage][r] = t
                            left[lane] -= 1
                            if left[lane] == 0 and fired[lane]:
                                bar.arrive()
                        pending_ops.append(land)
                    yield True
                for lane in range(LANES):                      # cp.async.mbarrier.arrive.noinc, every lane
                    fired[lane] = True
                    if left[lane] == 0:
                        bar.arrive()
                yield True
            stage += 1
            if stage == S:
                stage, parity = 0, parity ^ 1

    def consumer(w):
        stage, parity = 0, 0
        for t in range(n_tiles):
            yield True                                         # header load
            while not full[stage].done(parity):
                yield False
            readers[stage] += 1
            for _ in range(rng.randint(0, 3)):                 # entries of this warp in the tile
                r = rng.randrange(rows_per_tile)
                assert stage_tile[stage][r] == t, f"consumer {w} read tile {stage_tile[stage][r]} instead of {t}"
                reads[0] += 1
                yield True
            readers[stage] -= 1
            empty[stage].arrive()
            yield True
            stage += 1
            if stage == S:
                stage, parity = 0, parity ^ 1

    alive = [producer(j) for j in range(NP)] + [consumer(w) for w in range(W)]
    stalled = 0
    while alive:
        if pending_ops and rng.random() < 0.4:                 # an asynchronous copy lands, in any order
            pending_ops.pop(rng.randrange(len(pending_ops)))()
            stalled = 0
            continue
        th = rng.choice(alive)
        try:
            progressed = next(th)
        except StopIteration:
            alive.remove(th)
            progressed = True
        stalled = 0 if progressed else stalled + 1
        if stalled > 50 * (len(alive) + 1) and not pending_ops:
            raise AssertionError("deadlock")
    return reads[0]


@pytest.mark.parametrize("mode", ["bulk", "ldgsts"])
def test_protocol_never_deadlocks_or_reads_stale_tiles(mode):
    total = 0
    for seed in range(60):
        rng = random.Random(1000 + seed)
        total += simulate(n_tiles=rng.randint(0, 23), S=rng.randint(2, 5), W=rng.randint(1, 6), NP=rng.randint(1, 4),
                          mode=mode, seed=seed)
    assert total > 0


def test_model_detects_a_broken_protocol():
    """The model is sensitive: with the empty barrier initialised one arrival short (a consumer not waited for),
    a stale read, an overwrite of a stage that is being read, or a surplus arrival shows up."""
    failures = 0
    for seed in range(40):
        try:
            simulate(n_tiles=12, S=2, W=4, NP=1, mode="bulk", seed=seed, empty_count=3)
        except AssertionError:
            failures += 1
    assert failures > 0
