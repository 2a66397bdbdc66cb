"""Randomised model check of the synchronisation protocol of the fused K-upd kernel
(`ruvector_b200/csrc/gemm_upd_fused.cu`) -- no GPU needed.

The kernel's ten warps (TMA weight producer, MMA issuer, four A producers, four epilogue warps), the in-order
tensor pipe behind `tcgen05.mma` / `tcgen05.commit` and the TMA engine are agents of a discrete-event simulation;
a seeded adversarial scheduler interleaves them (agents get random, changing speeds: any warp may stall for a long
time, asynchronous work completes arbitrarily late).  The mbarriers follow the hardware rule the kernel relies on
(`try_wait.parity P` succeeds iff the phase of parity P has completed).  Shared-memory weight stages and TMEM
column blocks carry content tags.  Checked on every schedule:

  * every MMA finds in its A / B / D operands exactly the k-block, tile and accumulation state it was issued for;
  * nothing (tcgen05.st, TMA, an MMA's accumulator write) overwrites a block a queued MMA still has to read;
  * every epilogue read-out sees the completed accumulator of its own tile and phase;
  * a parity wait never passes on another phase than the one it was written for (no phase aliasing), nobody deadlocks.

The model mirrors the role loops of the kernel line by line (ring depths, barrier counts, wait order, the
`ring_free` / `pq_free` / `r_done` / drain gates, the opt-in stage-1 gate); a change of the kernel's protocol has to be
made here too.
What it does not model: the tensor pipe reordering MMAs on different accumulators (the kernel drains before C1
for that) and anything about arithmetic."""
import random

import pytest


class Violation(AssertionError):
    pass


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.tx, self.phases = name, count, count, 0, 0

    def _settle(self):
        if self.pending == 0 and self.tx == 0:
            self.phases += 1
            self.pending = self.count

    def arrive(self):
        if self.pending <= 0:
            raise Violation(f"{self.name}: arrival beyond the barrier's count")
        self.pending -= 1
        self._settle()

    def arrive_expect_tx(self, n):
        self.tx += n
        self.arrive()

    def complete_tx(self, n):
        self.tx -= n
        self._settle()

    def passed(self, parity):
        return (self.phases & 1) != parity


class Cell:
    def __init__(self, name):
        self.name, self.tag, self.pending_reads = name, None, 0

    def write(self, tag, who):
        if self.pending_reads:
            raise Violation(f"{who} overwrites {self.name} ({self.tag}) while {self.pending_reads} queued MMA(s) still read it")
        self.tag = tag


class Sim:
    """nk1: streamed k-blocks of stage 1; tiles: row tiles this CTA processes."""

    S_STAGES, M_STAGES, B_STAGES = 4, 2, 4

    def __init__(self, nk1, tiles, seed, gate=True, adversarial=True):
        self.nk1, self.tiles, self.rng = nk1, tiles, random.Random(seed)
        self.gate, self.adversarial = gate, adversarial
        B = lambda n, c: Bar(n, c)
        self.full_b = [B(f"full_b{s}", 1) for s in range(self.B_STAGES)]
        self.empty_b = [B(f"empty_b{s}", 1) for s in range(self.B_STAGES)]
        self.full_s = [B(f"full_s{s}", 4) for s in range(self.S_STAGES)]
        self.empty_s = [B(f"empty_s{s}", 1) for s in range(self.S_STAGES)]
        self.full_m = [B(f"full_m{s}", 4) for s in range(self.M_STAGES)]
        self.empty_m = [B(f"empty_m{s}", 1) for s in range(self.M_STAGES)]
        self.acc_bar, self.epi_bar = B("acc_bar", 1), B("epi_bar", 4)
        self.ring_free, self.drain_bar = B("ring_free", 4), B("drain_bar", 1)
        self.pq_free, self.r_done = B("pq_free", 1), B("r_done", 1)
        self.smem = [Cell(f"B{s}") for s in range(self.B_STAGES)]
        # TMEM, per lane quadrant: X, Y0 Y1 (message-row slots / candidate accumulator), PQ0..PQ3 (stage-1 slots /
        # resident operand).  Quadrant q is written by producer warp q and epilogue warp q only.
        self.tmem = {(n, q): Cell(f"{n}.q{q}") for n in ("X", "Y0", "Y1", "PQ0", "PQ1", "PQ2", "PQ3") for q in range(4)}
        self.pipe = []      # tensor pipe: in-order queue of ("mma", ...) / ("commit", bar)
        self.tma = []       # TMA loads in flight: (stage, tag)
        self.checked_mmas = 0
        self.readouts = 0

    # ---- asynchronous engines -------------------------------------------------------
    def cells(self, names):
        return [self.tmem[(n, q)] for n in names for q in range(4)]

    def issue_mma(self, a_names, a_tag, b_stage, b_tag, d_names, d_expect, d_new):
        reads = [(c, a_tag) for c in self.cells(a_names)] + [(self.smem[b_stage], b_tag)]
        for c, _ in reads:
            c.pending_reads += 1
        self.pipe.append(("mma", reads, self.cells(d_names), d_expect, d_new))

    def step_pipe(self):
        op = self.pipe.pop(0)
        if op[0] == "commit":
            op[1].arrive()
            return
        _, reads, d_cells, d_expect, d_new = op
        for c, tag in reads:
            if c.tag != tag:
                raise Violation(f"MMA reads {c.name}: found {c.tag}, issued for {tag}")
            c.pending_reads -= 1
        for c in d_cells:
            if d_expect is not None and c.tag != d_expect:
                raise Violation(f"MMA accumulates into {c.name}: found {c.tag}, expected {d_expect}")
            c.write(d_new, "an MMA's accumulator write")
        self.checked_mmas += 1

    def step_tma(self):
        stage, tag = self.tma.pop(self.rng.randrange(len(self.tma)))
        self.smem[stage].write(tag, "TMA")
        self.full_b[stage].complete_tx(1)

    # ---- warps ------------------------------------------------------------------------
    # a warp yields ("wait", bar, parity, nth): resume once the nth completion of bar (1-based) is visible
    def weight_order(self, tile):
        order = [("F", j, 0) for j in range(self.nk1)]
        for half in range(2):
            for t in range(8):
                order.append(("ZR", (t >> 1) + (t & 1) * 4, half))
        order += [("H", j, 0) for j in range(8)]
        return [(tile,) + o for o in order]

    def warp_tma(self):
        n = 0
        for tile in range(self.tiles):
            for tag in self.weight_order(tile):
                stage, use = n % self.B_STAGES, n // self.B_STAGES
                yield ("wait", self.empty_b[stage], (use & 1) ^ 1, use)
                self.full_b[stage].arrive_expect_tx(1)
                self.tma.append((stage, tag))
                n += 1

    def warp_mma(self):
        nb = ns = nm = 0
        n_epi = n_drain = 0

        def kblock(a_names, a_tag, d_names, d_expect, d_new, b_tag):
            nonlocal nb
            stage, use = nb % self.B_STAGES, nb // self.B_STAGES
            yield ("wait", self.full_b[stage], use & 1, use + 1)
            self.issue_mma(a_names, a_tag, stage, b_tag, d_names, d_expect, d_new)
            self.pipe.append(("commit", self.empty_b[stage]))
            nb += 1

        for tile in range(self.tiles):
            weights = iter(self.weight_order(tile))
            for kb in range(self.nk1):  # S1
                slot, use = ns % self.S_STAGES, ns // self.S_STAGES
                yield ("wait", self.full_s[slot], use & 1, use + 1)
                yield from kblock([f"PQ{slot}"], ("S", tile, kb), ["X"], None if kb == 0 else ("acc-a", tile, kb),
                                  ("acc-a", tile, kb + 1), next(weights))
                self.pipe.append(("commit", self.empty_s[slot]))
                ns += 1
            if self.gate and tile > 0:  # RUVGNN_FUSED_GATE: every epilogue warp has observed the previous C2
                yield ("wait", self.ring_free, (tile - 1) & 1, tile)
            self.pipe.append(("commit", self.acc_bar))
            for half, gate in enumerate(("z", "r")):
                n_epi += 1
                yield ("wait", self.epi_bar, (n_epi - 1) & 1, n_epi)  # E1 resp. Ez
                cnt = 0
                for kb in range(4):
                    yield from kblock(["PQ0", "PQ1", "PQ2", "PQ3"], ("a", tile), ["X"],
                                      None if cnt == 0 else (f"acc-{gate}", tile, cnt), (f"acc-{gate}", tile, cnt + 1),
                                      next(weights))
                    cnt += 1
                    slot, use = nm % self.M_STAGES, nm // self.M_STAGES
                    yield ("wait", self.full_m[slot], use & 1, use + 1)
                    yield from kblock([f"Y{slot}"], ("m", tile, half, kb), ["X"], (f"acc-{gate}", tile, cnt),
                                      (f"acc-{gate}", tile, cnt + 1), next(weights))
                    cnt += 1
                    self.pipe.append(("commit", self.empty_m[slot]))
                    nm += 1
                self.pipe.append(("commit", self.acc_bar if half == 0 else self.r_done))
            self.pipe.append(("commit", self.drain_bar))
            n_drain += 1
            yield ("wait", self.drain_bar, (n_drain - 1) & 1, n_drain)
            for kb in range(4):  # C1
                yield from kblock(["PQ0", "PQ1", "PQ2", "PQ3"], ("a", tile), ["Y0", "Y1"],
                                  None if kb == 0 else ("acc-c", tile, kb), ("acc-c", tile, kb + 1), next(weights))
            self.pipe.append(("commit", self.acc_bar))
            n_epi += 1
            yield ("wait", self.epi_bar, (n_epi - 1) & 1, n_epi)  # Er
            for kb in range(4):  # C2
                yield from kblock(["PQ0", "PQ1", "PQ2", "PQ3"], ("rh", tile), ["Y0", "Y1"], ("acc-c", tile, 4 + kb),
                                  ("acc-c", tile, 5 + kb), next(weights))
            self.pipe.append(("commit", self.acc_bar))
            self.pipe.append(("commit", self.pq_free))

    def warp_producer(self, q):
        ns = nm = 0
        for tile in range(self.tiles):
            for i in range(self.nk1 + 8):
                if i < self.nk1:
                    slot, use = ns % self.S_STAGES, ns // self.S_STAGES
                    yield ("wait", self.empty_s[slot], (use & 1) ^ 1, use)
                    if i == 0 and tile > 0:
                        yield ("wait", self.pq_free, (tile - 1) & 1, tile)
                    self.tmem[(f"PQ{slot}", q)].write(("S", tile, i), f"producer {q}")
                    yield ("pause",)  # tcgen05.wait::st, fence, __syncwarp
                    self.full_s[slot].arrive()
                    ns += 1
                else:
                    j = i - self.nk1
                    slot, use = nm % self.M_STAGES, nm // self.M_STAGES
                    yield ("wait", self.empty_m[slot], (use & 1) ^ 1, use)
                    if j == 0 and tile > 0:
                        yield ("wait", self.ring_free, (tile - 1) & 1, tile)
                    self.tmem[(f"Y{slot}", q)].write(("m", tile, j >> 2, j & 3), f"producer {q}")
                    yield ("pause",)
                    self.full_m[slot].arrive()
                    nm += 1

    def warp_epilogue(self, q):
        n_acc = 0

        def read(names, tag):
            for n in names:
                c = self.tmem[(n, q)]
                if c.tag != tag:
                    raise Violation(f"epilogue {q} reads {c.name}: found {c.tag}, expected {tag}")
            self.readouts += 1

        def write(names, tag):
            for n in names:
                self.tmem[(n, q)].write(tag, f"epilogue {q}")

        PQ = ["PQ0", "PQ1", "PQ2", "PQ3"]
        for tile in range(self.tiles):
            n_acc += 1
            yield ("wait", self.acc_bar, (n_acc - 1) & 1, n_acc)  # E1
            read(["X"], ("acc-a", tile, self.nk1))
            write(PQ, ("a", tile))
            yield ("pause",)
            self.epi_bar.arrive()
            n_acc += 1
            yield ("wait", self.acc_bar, (n_acc - 1) & 1, n_acc)  # Ez: copy out, hand X back, sigmoid on the copy
            read(["X"], ("acc-z", tile, 8))
            yield ("pause",)
            self.epi_bar.arrive()
            yield ("pause",)
            yield ("wait", self.r_done, tile & 1, tile + 1)  # Er, first half: gate math under C1, rh parked in X
            read(["X"], ("acc-r", tile, 8))
            write(["X"], ("rh-parked", tile))
            yield ("pause",)
            n_acc += 1
            yield ("wait", self.acc_bar, (n_acc - 1) & 1, n_acc)  # C1 complete: a is dead
            read(["X"], ("rh-parked", tile))
            write(PQ, ("rh", tile))
            yield ("pause",)
            self.epi_bar.arrive()
            n_acc += 1
            yield ("wait", self.acc_bar, (n_acc - 1) & 1, n_acc)  # E3
            read(["Y0", "Y1"], ("acc-c", tile, 8))
            yield ("pause",)
            self.ring_free.arrive()
            yield ("pause",)  # LayerNorm, copy-out

    # ---- scheduler --------------------------------------------------------------------
    def run(self):
        warps = {"tma": self.warp_tma(), "mma": self.warp_mma()}
        for q in range(4):
            warps[f"prod{q}"] = self.warp_producer(q)
            warps[f"epi{q}"] = self.warp_epilogue(q)
        waiting = {}
        for name, g in list(warps.items()):
            waiting[name] = next(g, None)
            if waiting[name] is None:
                del warps[name]
        steps = 0
        speed = {}
        while warps or self.pipe or self.tma:
            ready = []
            for name in warps:
                w = waiting[name]
                if w[0] == "pause" or w[1].passed(w[2]):
                    ready.append(name)
            if self.pipe:
                ready.append("@pipe")
            if self.tma:
                ready.append("@tma")
            if not ready:
                stuck = {n: (w[1].name, w[2], w[3], w[1].phases) for n, w in waiting.items() if n in warps}
                raise Violation(f"deadlock after {steps} steps: {stuck}")
            # adversarial scheduling: every agent gets a speed that is re-drawn now and then, so that single warps or
            # engines fall far behind (or run far ahead of) the rest
            if steps % 150 == 0:
                speed = {}
            if self.adversarial:
                pick = self.rng.choices(ready, weights=[speed.setdefault(n, self.rng.choice((0.01, 0.2, 1.0, 1.0, 8.0))) for n in ready])[0]
            else:  # fair: every ready agent is equally likely, nobody stalls for long
                pick = self.rng.choice(ready)
            steps += 1
            if pick == "@pipe":
                self.step_pipe()
                continue
            if pick == "@tma":
                self.step_tma()
                continue
            w = waiting[pick]
            if w[0] == "wait" and w[1].phases != w[3]:
                raise Violation(f"{pick}: wait on {w[1].name} for completion {w[3]} passed at completion {w[1].phases}")
            nxt = next(warps[pick], None)
            if nxt is None:
                del warps[pick]
            waiting[pick] = nxt
        return steps


@pytest.mark.parametrize("nk1", [8, 12, 20])
def test_protocol_with_the_stage1_gate_holds_on_adversarial_schedules(nk1):
    """RUVGNN_FUSED_GATE=1: the MMA warp completes stage 1 of a tile only after every epilogue warp has observed
    the previous tile's last GEMM.  No schedule breaks it."""
    for seed in range(80):
        sim = Sim(nk1, tiles=4, seed=seed * 7919 + nk1, gate=True, adversarial=True)
        sim.run()
        assert sim.checked_mmas == 4 * (nk1 + 24) and sim.readouts == 4 * 4 * 5


@pytest.mark.parametrize("nk1", [8, 12, 20])
def test_default_protocol_holds_when_no_warp_stalls_for_a_whole_stage(nk1):
    """The round's default (gate off, the instruction stream that was validated on the GPU): correct on every fair
    schedule ..."""
    for seed in range(60):
        sim = Sim(nk1, tiles=4, seed=seed * 104729 + nk1, gate=False, adversarial=False)
        sim.run()
        assert sim.checked_mmas == 4 * (nk1 + 24) and sim.readouts == 4 * 4 * 5


def test_default_protocol_aliases_a_phase_if_an_epilogue_warp_stalls_for_a_whole_stage():
    """... but not on all of them: stage 1 of tile n+1 does not depend on the epilogue, so acc_bar can complete
    twice (C2 of tile n, S1 of tile n+1) while a stalled epilogue warp still waits for the first -- it then waits on
    the wrong phase (in the kernel: the 4 s trap of mbar_wait).  Needs a warp that does not poll for ~19 us; the
    stage-1 gate closes it."""
    hits = []
    for seed in range(200):
        try:
            Sim(8, tiles=4, seed=seed, gate=False, adversarial=True).run()
        except Violation as e:
            hits.append(str(e))
    assert hits and all("acc_bar" in h for h in hits), hits[:3]


def test_the_checker_catches_a_missing_gate():
    """Sanity of the model itself: without the pq_free gate the next tile's gather-state k-blocks land in P | Q
    while the candidate GEMM still reads rh there -- some schedule must expose it."""

    class NoGate(Sim):
        def warp_producer(self, q):
            for step in super().warp_producer(q):
                if step[0] == "wait" and step[1] is self.pq_free:
                    continue
                yield step

    caught = 0
    for seed in range(40):
        try:
            NoGate(20, tiles=3, seed=seed, gate=True).run()
        except Violation:
            caught += 1
    assert caught > 0


def test_the_checker_catches_a_missing_drain():
    """Without waiting for ring_free the next tile's message rows overwrite the candidate accumulator in Y."""

    class NoRingFree(Sim):
        def warp_producer(self, q):
            for step in super().warp_producer(q):
                if step[0] == "wait" and step[1] is self.ring_free:
                    continue
                yield step

    hits = []
    for seed in range(60):
        try:
            NoRingFree(8, tiles=3, seed=seed, gate=True).run()
        except Violation as e:
            hits.append(str(e))
    assert any("Y0" in h or "Y1" in h for h in hits), hits[:3]
