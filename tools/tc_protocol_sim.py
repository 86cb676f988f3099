#!/usr/bin/env python
"""Discrete-event model of the synchronisation protocol of point_tc_kernel (disn_b200/csrc/point_tc.cu).

Why: on the GPU a protocol bug is a 4-second trap or silent corruption and costs a scarce GPU run; most of them (dead-
locks, a parity wait by an agent that skipped a phase, a ring slot overwritten before it was consumed, an accumulator
overwritten while it is still being drained) are properties of the *protocol*, not of the hardware.  This model runs
the same agents -- weight producers, MMA issuer, two epilogue groups, front end -- as Python generators over mbarrier
objects with the hardware's semantics (arrival counts, phases, parity waits that can alias), under randomised operation
latencies, and checks:
  * no deadlock (every agent terminates);
  * every successful parity wait observed exactly the phase the agent meant (no aliasing, no overrun);
  * every MMA reads the weight stage / activation slice it expects (ring contents are versioned);
  * no accumulator region is written while an overlapping older accumulator still has undrained slices.
One CTA of the pair is modelled (arrival counts that come from both CTAs are halved; the peer adds latency only).

    python tools/tc_protocol_sim.py [tiles] [schedules]
"""
import heapq
import random
import sys

NW, NX, NG = 3, 3, 2
SLICES = {0: 1, 1: 4, 2: 8, 3: 8}          # K slices per tensor-core layer (L0..L3 = fold1/conv2 .. fold2/conv2)
NNB = {0: 1, 1: 2, 2: 2, 3: 1}
XS = 20                                     # ring slices per stream (X3: 4, X4: 8, X5: 8); X2 has its own slot


def acc_cols(layer, q):
    """[start, end) TMEM columns of the accumulator written by `layer` of a stream with parity q (acc_col() + width)."""
    base = (128 if layer == 0 else (0 if layer == 2 else 256)) ^ (q << 8)
    return base, base + (128 if layer in (0, 3) else 256)


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals on " + self.name
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def parity_done(self, parity):          # mbarrier.try_wait.parity semantics
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, tiles, seed, mutate=None):
        self.mutate = mutate
        self.rng = random.Random(seed)
        self.t, self.q, self.n, self.progress_t = 0, [], 0, 0
        self.T, self.S = tiles, 2 * tiles
        self.mma_tail = 0                   # completion time of the last MMA in the tensor-pipe FIFO
        B = Bar
        self.wfull = [B("wfull%d" % i, 1) for i in range(NW)]      # {expect_tx arrival + bytes}: one event here
        self.wempty = [B("wempty%d" % i, 1) for i in range(NW)]
        self.xfull = [B("xfull%d" % i, 1) for i in range(NX)]       # per group-slice (4 warps x 2 CTAs in hardware)
        self.xempty = [B("xempty%d" % i, 1) for i in range(NX)]
        self.x2full, self.x2empty = B("x2full", 1), B("x2empty", 1)
        self.gfull = [B("gfull%d" % i, 1) for i in range(NG)]
        self.gempty = [B("gempty%d" % i, 1) for i in range(NG)]
        self.acc_full = [[B("acc%d_%d" % (l, nb), 1) for nb in range(2)] for l in range(4)]
        self.acc5_free = B("acc5_free", 2)                          # both epilogue groups
        self.w_slot, self.x_slot, self.x2_slot, self.g_slot = [None] * NW, [None] * NX, None, [None] * NG
        self.live = []                       # accumulators with undrained slices: [cols, remaining, tag]
        self.blocked = {}

    # ---- event loop -------------------------------------------------------------------------------
    def lat(self, lo, hi):
        return self.rng.randint(lo, hi)

    def at(self, dt, fn):
        self.n += 1
        heapq.heappush(self.q, (self.t + dt, self.n, fn))

    def spawn(self, name, gen):
        def step(val=None):
            try:
                op = gen.send(val)
            except StopIteration:
                self.blocked.pop(name, None)
                return
            kind = op[0]
            self.progress_t = self.t
            if kind == "delay":
                self.at(op[1], step)
            elif kind == "wait":
                _, bar, parity, want = op
                def poll():
                    if bar.parity_done(parity):
                        assert bar.phase == want + 1, "%s: wait on %s meant phase %d but the barrier has completed %d" % (
                            name, bar.name, want, bar.phase)
                        self.blocked.pop(name, None)
                        self.at(self.lat(20, 200), step)
                    else:
                        self.blocked[name] = (bar.name, want, bar.phase)
                        self.at(100, poll)
                poll()
            else:
                raise ValueError(kind)
        self.blocked[name] = ("start", 0, 0)
        self.at(0, step)

    def run(self):
        while self.q:
            self.t, _, fn = heapq.heappop(self.q)
            if self.t - self.progress_t > 300_000:      # far beyond any latency in the model: nobody can move
                raise AssertionError("no progress: " + repr(self.blocked))
            fn()
        assert not self.blocked, "deadlock: " + repr(self.blocked)

    # ---- tensor pipe ------------------------------------------------------------------------------
    def issue_mmas(self, dur):
        self.mma_tail = max(self.mma_tail, self.t) + dur

    def commit(self, bar):                   # tcgen05.commit: arrive when everything issued so far has retired
        self.at(max(0, self.mma_tail - self.t) + self.lat(10, 100), bar.arrive)

    # ---- agents -----------------------------------------------------------------------------------
    def stage_of(self, g):
        """consumption index -> (stream, layer, slice, nb) in the MMA warp's issue order."""
        return self.order[g]

    def build_order(self):
        order = []
        def layer(sn, l):
            for t in range(SLICES[l]):
                for nb in range(NNB[l]):
                    order.append((sn, l, t, nb))
        layer(0, 0)
        for sn in range(self.S):
            layer(sn, 1); layer(sn, 2)
            if sn + 1 < self.S:
                layer(sn + 1, 0)
            layer(sn, 3)
        self.order = order

    def producer(self, pw):
        g = pw
        while g < len(self.order):
            use = g // NW
            if self.mutate != "no_wempty_wait":
                yield ("wait", self.wempty[pw], (use & 1) ^ 1, use - 1)
            yield ("delay", self.lat(50, 300))
            def land(g=g, pw=pw):
                self.w_slot[pw] = g
                self.wfull[pw].arrive()
            self.at(self.lat(600, 2500), land)          # bulk copy + peer relay
            yield ("wait", self.wfull[pw], use & 1, use)
            yield ("delay", self.lat(50, 300))
            g += NW

    def mma(self):
        g = 0
        xseq = 0
        for (sn, l, t, nb) in self.order:
            q = sn & 1
            if t == 0 and nb == 0:
                if l == 2 and sn > 0 and self.mutate != "no_acc5_wait":
                    yield ("wait", self.acc5_free, (sn - 1) & 1, sn - 1)
                cols = acc_cols(l, q)
                for c, rem, tag in self.live:           # write-after-read hazard on TMEM
                    assert rem == 0 or c[1] <= cols[0] or cols[1] <= c[0], \
                        "L%d of stream %d overwrites %s with %d undrained slices" % (l, sn, tag, rem)
                self.live = [e for e in self.live if e[1] > 0]
            if nb == 0:                                  # activation slice
                if l == 0:
                    yield ("wait", self.x2full, sn & 1, sn)
                    assert self.x2_slot == sn, "X2 slot holds stream %r, wanted %d" % (self.x2_slot, sn)
                else:
                    slot = xseq % NX
                    yield ("wait", self.xfull[slot], (xseq // NX) & 1, xseq // NX)
                    want = sn * XS + {1: 0, 2: 4, 3: 12}[l] + t
                    assert self.x_slot[slot] == want, "X slot %d holds %r, wanted %d" % (slot, self.x_slot[slot], want)
            st = g % NW
            yield ("wait", self.wfull[st], (g // NW) & 1, g // NW)
            assert self.w_slot[st] == g, "W slot %d holds stage %r, wanted %d" % (st, self.w_slot[st], g)
            yield ("delay", self.lat(100, 400))
            self.issue_mmas(512)
            self.commit(self.wempty[st])
            if t == SLICES[l] - 1:                       # this N-block (128 columns = 4 drain slices) is final
                self.commit(self.acc_full[l][nb])
                c0 = acc_cols(l, q)[0] + 128 * nb
                self.at(max(0, self.mma_tail - self.t),
                        lambda c=(c0, c0 + 128), tag="acc of L%d/%d stream %d" % (l, nb, sn): self.live.append([c, 4, tag]))
            if nb == NNB[l] - 1:
                if l == 0:
                    self.commit(self.x2empty)
                else:
                    self.commit(self.xempty[xseq % NX])
                    xseq += 1
            g += 1

    def drained(self, l, q, t):              # 32-column slice t of the accumulator written by layer l was read
        c0 = acc_cols(l, q)[0] + 128 * (t // 4)
        cols = (c0, c0 + 128)
        for e in self.live:
            if e[0] == cols and e[1] > 0:
                e[1] -= 1
                return
        raise AssertionError("drain of an accumulator that is not complete: L%d parity %d" % (l, q))

    def epilogue(self, eg):
        gcount = 0
        def drain(sn, l_acc, t, seq, gather):
            """read slice t of the accumulator written by layer l_acc, write activation slice `seq`"""
            nonlocal gcount
            yield ("delay", self.lat(100, 300))          # tcgen05.ld + bias
            if gather:
                yield ("wait", self.gfull[eg], gcount & 1, gcount)
                assert self.g_slot[eg] == (sn >> 1) * 8 + t, "gather slot %d holds %r" % (eg, self.g_slot[eg])
                gcount += 1
            self.drained(l_acc, sn & 1, t)
            slot = seq % NX
            if self.mutate != "no_xempty_wait":
                yield ("wait", self.xempty[slot], ((seq // NX) & 1) ^ 1, seq // NX - 1)
            yield ("delay", self.lat(300, 900))          # convert + store
            self.x_slot[slot] = seq
            self.xfull[slot].arrive()
            if gather:
                self.gempty[eg].arrive()
        def x3(sn):
            for t in range(eg, 4, 2):
                if t == eg:
                    yield ("wait", self.acc_full[0][0], sn & 1, sn)
                yield from drain(sn, 0, t, sn * XS + t, False)
        if self.S:
            yield from x3(0)
        for sn in range(self.S):
            for l_acc, off in ((1, 4), (2, 12)):
                for t in range(eg, 8, 2):
                    if t == eg and self.mutate != "no_acc_wait":
                        yield ("wait", self.acc_full[l_acc][0], sn & 1, sn)
                    if t == eg + 4:
                        yield ("wait", self.acc_full[l_acc][1], sn & 1, sn)
                    yield from drain(sn, l_acc, t, sn * XS + off + t, l_acc == 2 and (sn & 1) == 1)
            if sn + 1 < self.S and self.mutate != "x3_after_final":
                yield from x3(sn + 1)
            yield ("wait", self.acc_full[3][0], sn & 1, sn)
            for t in (2 * eg, 2 * eg + 1):
                yield ("delay", self.lat(100, 300))
                self.drained(3, sn & 1, t)
            self.acc5_free.arrive()
            if sn + 1 < self.S and self.mutate == "x3_after_final":
                yield from x3(sn + 1)

    def front_end(self):
        def stage_x2(use):
            yield ("wait", self.x2empty, (use & 1) ^ 1, use - 1)
            yield ("delay", self.lat(200, 600))
            self.x2_slot = use
            self.x2full.arrive()
        if self.T:
            yield ("delay", self.lat(200, 800))          # points of tile 0
            yield from stage_x2(0)
        for it in range(self.T):
            yield from stage_x2(2 * it + 1)
            if it + 1 < self.T:
                yield ("delay", self.lat(200, 800))      # points of the next tile
                yield from stage_x2(2 * it + 2)
            for t in range(8):
                gsq = it * 8 + t
                gs = gsq % NG
                yield ("wait", self.gempty[gs], ((gsq // NG) & 1) ^ 1, gsq // NG - 1)
                yield ("delay", self.lat(800, 2500))
                self.g_slot[gs] = gsq
                self.gfull[gs].arrive()



def simulate(tiles=3, seed=0, mutate=None):
    """mutate: None (the kernel's protocol); 'no_xempty_wait', 'no_wempty_wait', 'no_acc_wait' are deliberately broken variants
    (tests/test_protocol_cpu.py uses them to show that the checks bite); 'x3_after_final' (the un-skewed epilogue order) and
    'no_acc5_wait' are legal-but-slower / redundant-in-this-model variants."""
    s = Sim(tiles, seed, mutate)
    s.build_order()
    for pw in range(NW):
        s.spawn("producer%d" % pw, s.producer(pw))
    s.spawn("mma", s.mma())
    for eg in range(2):
        s.spawn("epilogue%d" % eg, s.epilogue(eg))
    s.spawn("front", s.front_end())
    s.run()
    return s.t


if __name__ == "__main__":
    tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    ts = [simulate(tiles, seed) for seed in range(n)]
    print("%d schedules x %d tiles: no deadlock, no aliased wait, no ring/TMEM hazard; model time %d..%d cycles/tile"
          % (n, tiles, min(ts) // tiles, max(ts) // tiles))
