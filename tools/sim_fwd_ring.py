"""Discrete-event model of the barrier protocol of conv_stack_fwd_kernel<MODE, RING=1> (csrc/conv_fwd_bf16.cuh).

The ring variant was written without access to a GPU; this model replays its producer / consumer protocol -- the MMA
warp, the tensor pipe (in-order, asynchronous commits), the two front-end groups and the back-end -- under random
scheduling with mbarrier semantics (arrival counts, phase parity, `try_wait.parity`), and checks that
  * nothing deadlocks,
  * no TMEM region (the three accumulator slots, the layer-2 half region) is overwritten before it was drained or
    read before it was produced, and every unit is drained exactly once, in order,
  * the A1 / A2 shared-memory tile of an item is not rewritten while MMAs that read it are outstanding.
Each actor is a generator transcribed from the kernel; `yield ("wait", bar, parity)` blocks, everything else is a step.

    python tools/sim_fwd_ring.py            # a few thousand random schedules over ragged item sizes
"""
from __future__ import annotations

import random
import sys


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def done(self, parity):                 # try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


class Sim:
    def __init__(self, items, nchunk, rng):
        # items: list of NT (padded point counts, multiples of 16, <= 256)
        self.items, self.nchunk, self.rng = items, nchunk, rng
        B = self.bars = {}
        B["a1_full"] = MBar(2)              # 256 threads = 2 groups in the model
        B["d2_full"], B["d2_full1"] = MBar(1), MBar(1)
        B["d2_empty"] = MBar(1)             # group 0 (128 threads)
        for i in range(2):
            B[f"a2_full{i}"], B[f"a2_empty{i}"] = MBar(2), MBar(1)
        for i in range(3):
            B[f"acc_full{i}"], B[f"acc_empty{i}"] = MBar(1), MBar(1)   # back-end = one actor
        B["bar_sync1"] = MBar(2)            # the named barrier `bar.sync 1, 256` of the front end, once per item
        self.pipe = []                      # tensor pipe queue: ("mma", reads, writes) / ("commit", bar)
        self.slot = [None] * 3              # accumulator slot content: None = free, else (li, j, h) written & not drained
        self.d2 = None                      # content of the layer-2 region: (li, half) or None
        self.buf_reads = [0, 0]             # outstanding MMAs reading smem buffer b
        self.buf_owner = [None, None]       # ("a1", li) / ("a2", li): what the buffer holds (rows of item li)
        self.drained = []
        self.errors = []

    @staticmethod
    def halves(NT):
        N0 = min(NT, ((NT >> 1) + 15) & ~15)
        return N0, NT - N0

    # ---- actors ------------------------------------------------------------------------------------------
    def front(self, g):
        ph_d2, ph_a2e = 0, [0, 0]
        for li, NT in enumerate(self.items):
            b = li & 1
            N0, N1 = self.halves(NT)
            if li >= 2:
                yield ("wait", f"a2_empty{b}", ph_a2e[b]); ph_a2e[b] ^= 1
            self.bars["bar_sync1"].arrive()                  # raw points staged: bar.sync 1, 256
            yield ("wait", "bar_sync1", li & 1)
            # layer 1 writes the A1 tile into buffer b
            if self.buf_reads[b]:
                self.errors.append(f"front {g}: A1 of item {li} written while MMAs read buffer {b}")
            self.buf_owner[b] = ("a1", li)
            yield ("step",)
            self.bars["a1_full"].arrive()
            has = (N0 > 0) if g == 0 else (N1 > 0)
            if has:
                yield ("wait", "d2_full1" if g else "d2_full", ph_d2); ph_d2 ^= 1
                if self.d2 != (li, g):
                    self.errors.append(f"front {g}: item {li} reads layer-2 region holding {self.d2}")
                yield ("step",)             # TMEM loads + A2 stores of this group's rows
                self.d2 = None
            if g == 0 and N1 > 0:
                self.bars["d2_empty"].arrive()
            self.bars[f"a2_full{b}"].arrive()

    def back(self):
        unit = 0
        for li, NT in enumerate(self.items):
            N0, N1 = self.halves(NT)
            for j in range(self.nchunk):
                for h, n in enumerate((N0, N1)):
                    if n == 0:
                        continue
                    slot, par = unit % 3, (unit // 3) & 1
                    unit += 1
                    yield ("wait", f"acc_full{slot}", par)
                    if self.slot[slot] != (li, j, h):
                        self.errors.append(f"back: expects {(li, j, h)} in slot {slot}, found {self.slot[slot]}")
                    yield ("step",)
                    self.drained.append((li, j, h))
                    self.slot[slot] = None
                    self.bars[f"acc_empty{slot}"].arrive()

    def mma(self):
        ph_a1 = ph_d2e = 0
        ph_a2f = [0, 0]
        unit = 0
        state = {"stage": 0}

        def l2_ring(li, block):
            nonlocal ph_a1, ph_d2e
            N0, N1 = self.halves(self.items[li])
            b = li & 1
            if state["stage"] == 0:
                if not block and not self.bars["a1_full"].done(ph_a1):
                    return
                yield ("wait", "a1_full", ph_a1); ph_a1 ^= 1
                self.pipe.append(("mma_l2", li, 0, b)); self.buf_reads[b] += 1
                self.pipe.append(("commit", "d2_full"))
                state["stage"] = 1 if N1 > 0 else 2
            if state["stage"] == 1:
                if not block and not self.bars["d2_empty"].done(ph_d2e):
                    return
                yield ("wait", "d2_empty", ph_d2e); ph_d2e ^= 1
                self.pipe.append(("mma_l2", li, 1, b)); self.buf_reads[b] += 1
                self.pipe.append(("commit", "d2_full1"))
                state["stage"] = 2

        yield from l2_ring(0, True)
        n = len(self.items)
        for li, NT in enumerate(self.items):
            N0, N1 = self.halves(NT)
            b = li & 1
            state["stage"] = 0
            yield ("wait", f"a2_full{b}", ph_a2f[b]); ph_a2f[b] ^= 1
            for j in range(self.nchunk):
                for h, nn in enumerate((N0, N1)):
                    if nn == 0:
                        continue
                    slot, par = unit % 3, ((unit // 3) & 1) ^ 1
                    unit += 1
                    yield ("wait", f"acc_empty{slot}", par)
                    self.pipe.append(("mma_l3", li, j, h, slot, b)); self.buf_reads[b] += 1
                    self.pipe.append(("commit", f"acc_full{slot}"))
                yield ("step",)
                if li + 1 < n and state["stage"] < 2:
                    yield from l2_ring(li + 1, j == self.nchunk - 1)
            self.pipe.append(("commit", f"a2_empty{b}"))

    def tensor_pipe(self):
        while True:
            if not self.pipe:
                yield ("idle",)
                continue
            op = self.pipe.pop(0)
            if op[0] == "commit":
                self.bars[op[1]].arrive()
            elif op[0] == "mma_l2":
                _, li, half, b = op
                if self.buf_owner[b] != ("a1", li):
                    self.errors.append(f"pipe: layer 2 of item {li} reads buffer {b} holding {self.buf_owner[b]}")
                if self.d2 is not None:
                    self.errors.append(f"pipe: layer-2 half {(li, half)} overwrites undrained {self.d2}")
                self.d2 = (li, half)
                self.buf_reads[b] -= 1
            else:
                _, li, j, h, slot, b = op
                if self.buf_owner[b] != ("a1", li):
                    self.errors.append(f"pipe: layer 3 of item {li} reads buffer {b} holding {self.buf_owner[b]}")
                self.buf_reads[b] -= 1
                if self.slot[slot] is not None:
                    self.errors.append(f"pipe: unit {(li, j, h)} overwrites undrained {self.slot[slot]} in slot {slot}")
                self.slot[slot] = (li, j, h)
            yield ("step",)

    # ---- scheduler ---------------------------------------------------------------------------------------
    def run(self, max_steps=200000):
        actors = {"fe0": self.front(0), "fe1": self.front(1), "be": self.back(), "mma": self.mma(), "pipe": self.tensor_pipe()}
        blocked = {}
        alive = set(actors) - {"pipe"}
        for _ in range(max_steps):
            if not alive:
                break
            names = [a for a in actors if a in alive or a == "pipe"]
            name = self.rng.choice(names)
            if name in blocked:
                bar, par = blocked[name]
                if not self.bars[bar].done(par):
                    continue
                del blocked[name]
            try:
                ev = next(actors[name])
            except StopIteration:
                alive.discard(name)
                continue
            if ev[0] == "wait":
                blocked[name] = (ev[1], ev[2])
        else:
            return f"deadlock (step limit): blocked on {blocked}, still running {sorted(alive)}"
        # drain the pipe
        while self.pipe:
            next(actors["pipe"])
        want = [(li, j, h) for li, NT in enumerate(self.items) for j in range(self.nchunk)
                for h, nn in enumerate(self.halves(NT)) if nn]
        if self.drained != want:
            return "units drained out of order or missing"
        return "; ".join(self.errors[:3]) if self.errors else None


def main(trials=3000, seed=0):
    rng = random.Random(seed)
    sizes = [16, 32, 96, 112, 128, 144, 208, 256]
    bad = 0
    for t in range(trials):
        n = rng.choice([1, 2, 3, 5, 8])
        items = [rng.choice(sizes) for _ in range(n)] if rng.random() < 0.5 else [rng.choice(sizes)] * n
        nchunk = rng.choice([1, 2, 4, 8])
        err = Sim(items, nchunk, random.Random(rng.random())).run()
        if err:
            bad += 1
            print(f"trial {t}: items={items} nchunk={nchunk}: {err}")
            if bad > 5:
                break
    print(f"{trials} schedules, {bad} failures")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
