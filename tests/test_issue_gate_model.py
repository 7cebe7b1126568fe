"""A small executable model of the conv kernels' barrier protocol (conv1.cu / conv3x3.cu: one TMA producer, W
MMA-issuing warps taking the tiles round-robin, two epilogue groups, rings of S operand stages and A accumulators,
mbarrier waits BY PARITY) under arbitrary interleavings.

It pins the analysis behind ``issue_gate`` (csrc/sia_ptx.cuh): without the gate a delayed issuer
lets another one pass a parity wait on a phase that is two uses old (W = 3 over rings of 4; W = 2 over a ring of 3);
with the gate -- the issuer of tile t first waits until tile t - ring has been issued -- no schedule does, and every
schedule runs to completion.  CPU only; the product is the CUDA code, this is its invariant written down.
"""
import random

import pytest


class Barrier:
    """mbarrier reduced to what the protocol uses: ``done`` = completed phases; try_wait.parity(p) succeeds when the
    phase with parity p is the most recently completed one (a fresh barrier passes a wait on parity 1)."""

    def __init__(self):
        self.done = 0

    def passes(self, parity: int) -> bool:
        return (self.done & 1) != parity


class Stale(Exception):
    pass


def run(n_tiles: int, n_warps: int, n_stage: int, n_acc: int, gate: bool, rng: random.Random, starve: int | None,
        slow_tma: bool = False):
    """One random interleaving.  Returns "ok"; raises Stale when a wait is satisfied by the wrong phase; returns
    "deadlock" when nobody can move before all tiles are done."""
    full = [Barrier() for _ in range(n_stage)]
    empty = [Barrier() for _ in range(n_stage)]
    tfull = [Barrier() for _ in range(n_acc)]
    tempty = [Barrier() for _ in range(n_acc)]
    issued = [0] * n_warps

    in_flight: list[int] = []                       # stages whose TMA load has been issued and has not landed yet
    producer_done = [False]

    def producer():
        for t in range(n_tiles):
            s, ph = t % n_stage, (t // n_stage) & 1
            while not empty[s].passes(ph ^ 1):
                yield False
            in_flight.append(s)                     # the TMA load of tile t is issued ...
            yield True
        producer_done[0] = True

    def tma():
        while not (producer_done[0] and not in_flight):
            if not in_flight:
                yield False
                continue
            s = in_flight.pop(rng.randrange(len(in_flight)))      # ... and lands later, in any order
            full[s].done += 1
            yield True

    def gate_open(t: int, ring: int) -> bool:
        g = t - ring
        if g < 0 or g % n_warps == t % n_warps:
            return True
        return issued[g % n_warps] >= g // n_warps + 1

    def issuer(w: int):
        for t in range(w, n_tiles, n_warps):
            s, a = t % n_stage, t % n_acc
            if gate:
                while not (gate_open(t, n_stage) and gate_open(t, n_acc)):
                    yield False
            while not tempty[a].passes(((t // n_acc) & 1) ^ 1):
                yield False
            if tempty[a].done != t // n_acc:        # exactly the epilogues of tiles t - A, t - 2A, ... have released it
                raise Stale(f"issuer {w}, tile {t}: accumulator {a} taken after {tempty[a].done} releases")
            yield True
            while not full[s].passes((t // n_stage) & 1):
                yield False
            if full[s].done != t // n_stage + 1:    # exactly the loads of tiles ..., t - S, t have landed
                raise Stale(f"issuer {w}, tile {t}: stage {s} read after {full[s].done} loads")
            empty[s].done += 1                      # tcgen05.commit -> empty barrier
            tfull[a].done += 1                      # tcgen05.commit -> accumulator full
            issued[w] += 1
            yield True

    def epilogue(g: int):
        for j in range(g, n_tiles, 2):
            a = j % n_acc
            while not tfull[a].passes((j // n_acc) & 1):
                yield False
            if tfull[a].done != j // n_acc + 1:
                raise Stale(f"epilogue {g}, tile {j}: accumulator {a} read after {tfull[a].done} commits")
            tempty[a].done += 1
            yield True

    actors = {"p": producer(), "tma": tma(), "e0": epilogue(0), "e1": epilogue(1)}
    actors.update({f"i{w}": issuer(w) for w in range(n_warps)})
    live = dict(actors)
    blocked: set[str] = set()
    while live:
        names = [n for n in live if n not in blocked]
        if not names:
            return "deadlock"
        if starve is not None and f"i{starve}" in names and len(names) > 1 and rng.random() < 0.97:
            names.remove(f"i{starve}")              # one issuer hardly ever gets an issue slot
        if slow_tma and "tma" in names and len(names) > 1 and rng.random() < 0.8:
            names.remove("tma")                     # loads stay in flight for a while
        name = rng.choice(names)
        try:
            progressed = next(live[name])
        except StopIteration:
            del live[name]
            blocked.clear()
            continue
        if progressed:
            blocked.clear()
        else:
            blocked.add(name)
    return "ok"


CONFIGS = [(3, 4, 4), (2, 3, 4), (2, 6, 4), (3, 3, 4), (2, 4, 4)]      # (issuing warps, operand stages, accumulators)


@pytest.mark.parametrize("n_warps,n_stage,n_acc", CONFIGS)
def test_gate_makes_every_parity_wait_exact(n_warps, n_stage, n_acc):
    rng = random.Random(1234 + 7 * n_warps + n_stage)
    for trial in range(400):
        starve = None if trial % 4 == 0 else trial % n_warps
        assert run(60, n_warps, n_stage, n_acc, True, rng, starve, slow_tma=trial % 3 == 0) == "ok"


@pytest.mark.parametrize("n_warps,n_stage,n_acc", [(3, 4, 4), (2, 3, 4)])
def test_without_the_gate_a_delayed_issuer_aliases_a_phase(n_warps, n_stage, n_acc):
    """conv1 / the pixel-pair kernel (3 issuers, rings of 4) and conv3x3_kernel<64,128,3> (2 issuers, 3 stages) as they
    were: some schedule with one starved issuer (and, for the second, loads that land out of order) passes a wait on a
    stale phase (or deadlocks)."""
    rng = random.Random(99)
    outcomes = set()
    for trial in range(400):
        try:
            outcomes.add(run(60, n_warps, n_stage, n_acc, False, rng, trial % n_warps))
        except Stale:
            outcomes.add("stale")
    assert "stale" in outcomes or "deadlock" in outcomes


def test_two_issuers_over_rings_of_four_never_alias_even_without_the_gate():
    rng = random.Random(5)
    for trial in range(300):
        assert run(60, 2, 4, 4, False, rng, trial % 2) == "ok"


def test_the_kernels_use_modelled_configurations_and_call_the_gate():
    """Ties the model to the sources: the (issuers, stages, accumulators) of every multi-issuer kernel instance are
    among CONFIGS, and each of their issue loops calls issue_gate / issue_done."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "skin_image_analysis_b200", "csrc")
    conv1 = open(os.path.join(root, "conv1.cu")).read()
    conv3 = open(os.path.join(root, "conv3x3.cu")).read()

    def define(text, name):
        return int(re.search(rf"#define {name} (\d+)", text).group(1))

    def constexpr(text, name):
        return int(re.search(rf"constexpr int {name} = (\d+);", text).group(1))

    found = {(define(conv1, "SIA_C1_MMA_WARPS"), define(conv1, "SIA_C1_NSTAGE"), define(conv1, "SIA_C1_NACC")),
             (define(conv3, "SIA_CP_MMA_WARPS"), constexpr(conv3, "CP_NSTAGE"), constexpr(conv3, "CV_NACC"))}
    for stages in re.findall(r"launch_conv3x3<\d+, \d+, (\d+)>", conv3):
        found.add((define(conv3, "SIA_CV_MMA_WARPS"), int(stages), constexpr(conv3, "CV_NACC")))
    assert found and found <= set(CONFIGS), found - set(CONFIGS)
    assert conv1.count("issue_gate(issued, lt,") >= 1 and conv1.count("issue_done(issued, lt,") == 1
    assert conv3.count("issue_gate(issued, lt,") >= 2 and conv3.count("issue_done(issued, lt,") == 2
