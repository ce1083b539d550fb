"""Executable model of the shared-memory ring protocol of csrc/syncbn.cu (producer warp / consumer warps).

The kernels stream a strip of `n` chunks through `NS` stages twice (phase 1: statistics, phase 2: normalize) while
keeping the last ring-full resident between the phases.  The index arithmetic (`load_chunk_of`, the stage of a use,
when a consumer releases a stage) is easy to get subtly wrong and a mistake means a deadlock or a silently
overwritten chunk on the GPU; this test replays exactly those formulas on the CPU for every (n, NS) and checks:
every use finds the chunk it expects, no stage is overwritten before its last reader is done, every producer wait is
matched by exactly one release (so the mbarrier parities stay in step), and nothing deadlocks.
"""
import pytest


def load_chunk_of(k, n, nres0):           # csrc/syncbn.cu: load k of the CTA's sequence → chunk index
    return k if k < n else nres0 - 1 - (k - n)


def simulate(n, NS, training=True):
    nres0 = max(n - NS, 0) if training else 0
    total_loads = n + nres0 if training else n
    nresident = n - nres0 if training else 0

    # consumer side: the ordered list of uses (phase, chunk, stage, waits_for_load, releases)
    uses = []
    if training:
        for i in range(n):                                              # phase 1
            uses.append(("stats", i, i % NS, i, i + NS < n))
    for u in range(n):                                                  # phase 2
        if training:
            streamed = u >= nresident
            c = nres0 - 1 - (u - nresident) if streamed else nres0 + u
            kload = n + (u - nresident)
            s = kload % NS if streamed else (nres0 + u) % NS
            stage_load = kload if streamed else nres0 + u
        else:
            streamed, c, kload, s, stage_load = True, u, u, u % NS, u
        uses.append(("norm", c, s, kload if streamed else None, stage_load + NS < total_loads))

    # event-driven replay: producer may issue load k once (k < NS) or the stage's previous occupant was released
    stage_content = [None] * NS          # chunk currently in the stage
    stage_released = [True] * NS         # free for the producer
    loads_done = set()
    k = 0                                # next load to issue
    ui = 0                               # next use
    releases = waits = 0
    progress = True
    while progress:
        progress = False
        # producer
        while k < total_loads:
            s = k % NS
            if k >= NS:
                if not stage_released[s]:
                    break
                waits += 1
            stage_released[s] = False
            stage_content[s] = load_chunk_of(k, n, nres0)
            loads_done.add(k)
            k += 1
            progress = True
        # consumers (all warps move in lock-step through the uses)
        while ui < len(uses):
            phase, chunk, s, need_load, rel = uses[ui]
            if need_load is not None and need_load not in loads_done:
                break
            assert stage_content[s] == chunk, (n, NS, phase, ui, chunk, stage_content[s])
            if rel:
                assert not stage_released[s]
                stage_released[s] = True
                releases += 1
            ui += 1
            progress = True
    assert ui == len(uses) and k == total_loads, f"deadlock: n={n} NS={NS} use {ui}/{len(uses)} load {k}/{total_loads}"
    assert releases == waits == max(total_loads - NS, 0)
    # every chunk is normalised exactly once, and (training) seen exactly once by the statistics pass
    assert sorted(c for p, c, *_ in uses if p == "norm") == list(range(n))
    if training:
        assert [c for p, c, *_ in uses if p == "stats"] == list(range(n))
    return total_loads


@pytest.mark.parametrize("NS", [2, 3, 4, 5, 7, 10, 16])
def test_ring_protocol_training(NS):
    for n in range(1, 60):
        loads = simulate(n, NS, training=True)
        assert loads == n + max(n - NS, 0)        # only what did not stay resident is read a second time


@pytest.mark.parametrize("NS", [2, 4, 5, 16])
def test_ring_protocol_eval(NS):
    for n in range(1, 40):
        assert simulate(n, NS, training=False) == n


@pytest.mark.parametrize("NS", [2, 3, 4, 7, 16])
def test_l2_hint_classification(NS):
    """producer_loop<…, HINT>: a load is issued evict-last iff that chunk is loaded a second time later; every other load
    (chunks that stay resident, and the second loads themselves) is a last use and goes evict-first"""
    for n in range(1, 60):
        nres0 = max(n - NS, 0)
        total = n + nres0
        seq = [load_chunk_of(k, n, nres0) for k in range(total)]
        for k, chunk in enumerate(seq):
            again_formula = k < n and k < nres0                    # csrc/syncbn.cu producer_loop
            again_truth = chunk in seq[k + 1:]
            assert again_formula == again_truth, (n, NS, k, chunk)
        # every chunk is loaded at most twice, and the second load is never marked for keeping
        assert all(seq.count(c) <= 2 for c in set(seq))
