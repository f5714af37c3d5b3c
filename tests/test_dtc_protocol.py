"""Host-side model of the synchronisation protocol of the streamed dense kernel (pycqed_b200/csrc/dense_tiled.cuh,
dtc_stream / dtc_tridiag_kernel): one producer (bulk copies into a BYTE ring, up to NBAR pieces in flight, running ahead
into the next pass), CW consumers (every consumer takes every piece in order, hands its column part over once per tile
column) and one folder, synchronised by mbarriers (full / empty per piece, tfull / tempty per hand-over slot) and a block
barrier between passes.

The model restates the control flow of the three roles line by line (piece enumeration, ring placement rule on both
sides, barrier indices and phase parities, the producer's run-ahead rule) with the arithmetic left out, runs the roles as
coroutines under a seeded random scheduler with randomly delayed copy completions, and checks what the GPU tests cannot
see reliably because it depends on timing:

  * no deadlock (some role or copy can always make progress until every pass is done);
  * every parity wait is unambiguous: the barrier is in the phase waited for or the one before (never two behind or ahead);
  * a piece is never overwritten in the ring before all consumers released it, and a consumer always reads the piece it
    expects; the folder always finds the parts of the column it expects from every consumer.

CPU only; it guards edits of the protocol (the first spanning-piece variant of this kernel deadlocked on the GPU: a warp
without tiles in the last pieces of a column blocked in the hand-over while holding a ring stage)."""
import random

import pytest

TILE = 512


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.ph = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.ph += 1
            self.pending = self.count

    def passed(self, use):
        """try_wait.parity((use) & 1): true once phase `use` completed.  The parity test is only unambiguous while the
        barrier is in phase `use` (not yet) or `use + 1` (done)."""
        assert self.ph in (use, use + 1), "ambiguous parity wait: barrier in phase %d, waiting for use %d" % (self.ph, use)
        return (self.ph & 1) != (use & 1)


class Model:
    def __init__(self, NT, NB, piece, ringb, nbar, cw, seed, matrices=2):
        assert ringb >= 2 * piece * TILE
        self.NT, self.NB, self.PIECE, self.ringb, self.NBAR, self.CW = NT, NB, piece, ringb, nbar, cw
        self.rng = random.Random(seed)
        self.full = [Bar(1) for _ in range(nbar)]          # completed by the copy (expect_tx arrival + bytes)
        self.empty = [Bar(cw) for _ in range(nbar)]
        self.tfull = [Bar(cw) for _ in range(2)]
        self.tempty = [Bar(1) for _ in range(2)]
        self.ring = {}                                      # byte offset of a tile slot -> piece index held there
        self.live = {}                                      # piece -> (off, sz) while not released by all consumers
        self.released = {}                                  # piece -> number of consumers that released it
        self.copies = []                                    # [ticks left, piece, off, sz]
        self.tpart = [[None] * cw for _ in range(2)]
        self.block = Bar(cw + 2)                            # __syncthreads between passes
        self.passes = self._passes(matrices)
        self.done = 0

    def _passes(self, matrices):
        n = 8 * self.NT
        out = []
        for _ in range(matrices):
            for j in range(n - 1):
                fused = j % self.NB == 0 and j > 0
                nxt = -1 if fused else ((j + 2) >> 3 if j + 1 < n - 1 else -1)
                out.append(((j + 1) >> 3, nxt))
        return out

    # ---- roles (generators yield a predicate: "resume me when this holds") -----------------------------------------------
    def producer(self):
        NT, PIECE, ringb, NBAR = self.NT, self.PIECE, self.ringb, self.NBAR
        P = dict(g=0, off=0, tail=0, infl=0, pre=0, pC=0, pR0=0)
        szs = [0] * NBAR

        def issue(C, R0, ahead, first):
            while C < NT:
                cnt = min(PIECE, NT - R0)
                sz = cnt * TILE
                off, skip = P["off"], 0
                if off + sz > ringb:
                    skip, off = ringb - off, 0
                need = sz + skip
                while P["infl"] + need > ringb or P["g"] - P["tail"] >= NBAR:
                    if ahead and P["tail"] >= first:
                        return C, R0, False
                    t = P["tail"]
                    yield lambda t=t: self.empty[t % NBAR].passed(t // NBAR)
                    P["infl"] -= szs[t % NBAR]
                    P["tail"] += 1
                g = P["g"]
                szs[g % NBAR] = need
                for p, (o, s) in self.live.items():         # nothing a consumer may still read is overwritten
                    assert off + sz <= o or o + s <= off, "piece %d overwrites live piece %d" % (g, p)
                self.live[g] = (off, sz)
                self.released[g] = 0
                self.copies.append([self.rng.randint(0, 12), g, off, sz])
                P["infl"] += need
                P["off"] = off + sz
                P["g"] += 1
                R0 += PIECE
                if R0 >= NT:
                    C += 1
                    R0 = C
            return C, R0, True

        for T0, nxt in self.passes:
            C, R0 = (P["pC"], P["pR0"]) if P["pre"] else (T0, T0)
            C, R0, ok = yield from issue(C, R0, False, 0)
            assert ok
            P["pre"] = 0
            if 0 <= nxt < NT:
                P["pC"], P["pR0"], ok = yield from issue(nxt, nxt, True, P["g"])
                P["pre"] = 1
            yield from self.sync()

    def consumer(self, w):
        NT, PIECE, ringb, NBAR = self.NT, self.PIECE, self.ringb, self.NBAR
        g, off, tc = 0, 0, 0
        for T0, _ in self.passes:
            for C in range(T0, NT):
                for R0 in range(C, NT, PIECE):
                    cnt = min(PIECE, NT - R0)
                    sz = cnt * TILE
                    if off + sz > ringb:
                        off = 0
                    yield lambda g=g: self.full[g % NBAR].passed(g // NBAR)
                    for q in range(cnt):
                        assert self.ring.get(off + q * TILE) == g, "consumer %d expected piece %d at %d" % (w, g, off)
                    self.empty[g % NBAR].arrive()
                    self.released[g] += 1
                    if self.released[g] == self.CW:
                        del self.live[g]
                    off += sz
                    g += 1
                slot, use = tc & 1, tc >> 1
                if use >= 1:
                    yield lambda slot=slot, use=use: self.tempty[slot].passed(use - 1)
                assert self.tpart[slot][w] is None
                self.tpart[slot][w] = tc
                self.tfull[slot].arrive()
                tc += 1
            yield from self.sync()

    def folder(self):
        tc = 0
        for T0, _ in self.passes:
            for C in range(T0, self.NT):
                slot, use = tc & 1, tc >> 1
                yield lambda slot=slot, use=use: self.tfull[slot].passed(use)
                assert self.tpart[slot] == [tc] * self.CW
                self.tpart[slot] = [None] * self.CW
                self.tempty[slot].arrive()
                tc += 1
            yield from self.sync()

    def sync(self):
        ph = self.block.ph
        self.block.arrive()
        yield lambda: self.block.ph > ph
        self.done += 1

    # ---- scheduler ---------------------------------------------------------------------------------------------------------
    def run(self):
        tasks = [self.producer(), self.folder()] + [self.consumer(w) for w in range(self.CW)]
        waiting = {}
        for t in tasks:
            waiting[t] = next(t)
        steps = 0
        while waiting:
            steps += 1
            # the copy engine: a copy whose delay ran out lands (ring content tagged, full barrier phase completes)
            for c in self.copies:
                c[0] -= 1
            for c in [c for c in self.copies if c[0] <= 0]:
                self.copies.remove(c)
                _, g, off, sz = c
                for q in range(sz // TILE):
                    self.ring[off + q * TILE] = g
                self.full[g % self.NBAR].arrive()
            ready = [t for t, cond in waiting.items() if cond()]
            if not ready:
                assert self.copies, "deadlock after %d steps (%d block barriers passed)" % (steps, self.block.ph)
                continue
            t = self.rng.choice(ready)
            try:
                waiting[t] = t.send(None)
            except StopIteration:
                del waiting[t]
        assert self.block.ph == len(self.passes)


@pytest.mark.parametrize("NT,NB,piece,ring_tiles,nbar", [
    (3, 4, 4, 8, 4), (5, 4, 2, 4, 2), (5, 8, 2, 5, 4), (7, 4, 4, 9, 16), (7, 2, 3, 7, 4), (9, 8, 4, 8, 2), (6, 4, 64, 128, 16),
])
def test_dtc_protocol_is_deadlock_free_and_unambiguous(NT, NB, piece, ring_tiles, nbar):
    for seed in range(6):
        Model(NT, NB, piece, ring_tiles * TILE, nbar, cw=4, seed=seed).run()


def test_dtc_protocol_model_detects_a_held_stage():
    """The model is not vacuous: a consumer that blocks in the hand-over BEFORE releasing the piece it holds (the bug of the
    first spanning-piece variant) is reported as a deadlock."""
    class Broken(Model):
        def consumer(self, w):
            if w != 0:
                yield from Model.consumer(self, w)
                return
            # consumer 0 never releases its pieces
            yield lambda: False

    with pytest.raises(AssertionError):
        Broken(5, 4, 2, 4 * TILE, 2, cw=4, seed=0, matrices=1).run()
