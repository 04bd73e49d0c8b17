"""A small executable model of the scan kernels' shared-memory protocol (dawnsearch_b200/csrc/scan_topk.cu):
one in-order producer filling a ring of stages, 16 consumer warps that own stages w, w+16, ..., an append
buffer with a high-water mark, and the prune rendezvous every consumer has to reach.

It exists because the first version of the protocol could deadlock on the GPU (a warp parked at the prune
barrier leaves its next landed stages unconsumed, the producer blocks on their slots, a warp waiting for a
later stage never reaches the barrier).  The model reproduces that deadlock under an adversarial schedule and
shows that the shipped rule -- a wait for data gives up when a prune is due (wait_stage_or_prune) -- cannot
get stuck under random and adversarial schedules.  Pure Python, no GPU.
"""
import random

import pytest

WARPS = 16
RING = 32
HIGH_WATER = 12     # scaled down with the stage size so that prunes are frequent
STAGE_ROWS = 2


class Model:
    def __init__(self, n_data, interruptible, rng, pass_rate=1.0):
        self.n_data = n_data
        self.n_total = n_data + WARPS          # one end-of-stream stage per warp
        self.interruptible = interruptible
        self.rng = rng
        self.pass_rate = pass_rate
        self.issued = 0                        # producer position: stages [0, issued) have been issued
        self.landed = set()                    # stages whose bytes are in the ring
        self.in_flight = set()
        self.slot_free = [True] * RING         # slot's previous stage has been released
        self.next = list(range(WARPS))         # next stage of each warp
        # TOP -> (WAIT ->) LOADED -> APPEND -> TOP ... ; BARRIER while joining a prune; END in the end-of-stream loop
        self.state = ["TOP"] * WARPS
        self.cnt = 0
        self.max_cnt = 0
        self.prunes = 0
        self.done = False

    # ---- enabled transitions -----------------------------------------------------------
    def enabled(self):
        ev = []
        if self.issued < self.n_total and self.slot_free[self.issued % RING]:
            ev.append(("issue", None))
        for g in self.in_flight:
            ev.append(("land", g))
        for w in range(WARPS):
            s = self.state[w]
            if s == "TOP":
                ev.append(("top", w))
            elif s == "WAIT":
                if self.next[w] in self.landed:
                    ev.append(("load", w))
                elif self.interruptible and self.cnt >= HIGH_WATER:
                    ev.append(("bail", w))
            elif s == "LOADED":
                ev.append(("append", w))
        if all(s in ("BARRIER", "END") for s in self.state) and any(s == "BARRIER" for s in self.state):
            ev.append(("prune", None))
        if all(s == "END" for s in self.state):
            ev.append(("finish", None))
        return ev

    def step(self, ev):
        kind, x = ev
        if kind == "issue":
            g = self.issued
            self.slot_free[g % RING] = False
            self.in_flight.add(g)
            self.issued += 1
        elif kind == "land":
            self.in_flight.discard(x)
            self.landed.add(x)
        elif kind == "top":                    # loop top: join a prune if the buffer is at its high-water mark
            self.state[x] = "BARRIER" if self.cnt >= HIGH_WATER else "WAIT"
        elif kind == "bail":                   # wait_stage_or_prune gave up: the stage is waited for again later
            self.state[x] = "BARRIER"
        elif kind == "load":                   # bytes -> registers, slot handed back to the producer
            g = self.next[x]
            self.landed.discard(g)
            self.slot_free[g % RING] = True
            self.state[x] = "END" if g >= self.n_data else "LOADED"
        elif kind == "append":                 # survivors -> shared buffer (possibly long after the load)
            rows = sum(1 for _ in range(STAGE_ROWS) if self.rng.random() < self.pass_rate)
            self.cnt += rows
            self.max_cnt = max(self.max_cnt, self.cnt)
            self.next[x] += WARPS
            self.state[x] = "TOP"
        elif kind == "prune":
            self.cnt = 0
            self.prunes += 1
            self.state = ["TOP" if s == "BARRIER" else s for s in self.state]
        elif kind == "finish":
            self.done = True


def run(n_data, interruptible, seed, adversary=None, pass_rate=1.0, max_steps=2_000_000):
    rng = random.Random(seed)
    m = Model(n_data, interruptible, rng, pass_rate)
    for _ in range(max_steps):
        if m.done:
            return "done", m
        ev = m.enabled()
        if not ev:
            return "deadlock", m
        if adversary is not None:
            ev = adversary(m, ev) or ev
        m.step(rng.choice(ev))
    return "timeout", m


def starve_one_append(victim):
    """Schedule everything else first: the victim's append (its label load queued behind the bulk copies on the
    GPU) only runs when nothing else can."""
    def pick(m, ev):
        others = [e for e in ev if not (e[0] == "append" and e[1] == victim)]
        return others
    return pick


def test_plain_wait_deadlocks_under_a_slow_append():
    """The first protocol: once one warp is 32 stages behind and the buffer crosses its high-water mark,
    nobody can move."""
    outcomes = {run(4000, interruptible=False, seed=s, adversary=starve_one_append(3))[0] for s in range(20)}
    assert "deadlock" in outcomes


@pytest.mark.parametrize("pass_rate", [1.0, 0.3, 0.02])
def test_interruptible_wait_never_deadlocks(pass_rate):
    for seed in range(40):
        adv = starve_one_append(seed % WARPS) if seed % 2 else None
        outcome, m = run(3000, interruptible=True, seed=seed, adversary=adv, pass_rate=pass_rate)
        assert outcome == "done", (seed, outcome, m.state, m.cnt)
        # the append buffer bound the kernel relies on: high water + one stage per warp
        assert m.max_cnt <= HIGH_WATER - 1 + WARPS * STAGE_ROWS


def test_every_stage_is_consumed_exactly_once():
    outcome, m = run(1000, interruptible=True, seed=7)
    assert outcome == "done" and m.issued == m.n_total and not m.landed and not m.in_flight
    assert all(n >= m.n_data for n in m.next)  # every warp ended on an end-of-stream stage
