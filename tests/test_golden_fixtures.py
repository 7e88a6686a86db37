"""Committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py): recorded
closed-loop traces with the sha256 of every step's outputs, of the final rows and the final
counters.  Replayed against the oracle and the host build of the device logic on CPU and against
the CUDA engine through the C ABI on a GPU; nothing here runs the oracle to learn what is expected.
"""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402
from ra_suite import make_backend  # noqa: E402

BACKENDS = ["oracle", "emu", pytest.param("engine", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", sorted(G.TRACES))
def test_golden_trace(name, backend):
    meta = json.load(open(os.path.join(HERE, "golden", name + ".json")))
    batches = G.unpack_batches(open(os.path.join(HERE, "golden", name + ".events.z"), "rb").read())
    assert len(batches) == meta["steps"] and sum(len(b) for b in batches) == meta["events"]
    b = make_backend(backend, meta["groups"], meta["members"], **meta["cfg"])
    per_step, rows, counters = G.replay_digests(b, batches)
    for t, (got, want) in enumerate(zip(per_step, meta["step_sha256"])):
        assert got == want, "outputs of step %d differ from the fixture" % t
    assert rows == meta["rows_sha256"]
    assert {k: counters[k] for k in meta["counters"]} == meta["counters"]      # (a fixture may predate a counter)


def test_fixtures_match_their_generator():
    """The committed files are what make_golden.py produces today (same seeds, same simulator)."""
    import trace_gen
    from oracle_lib import Oracle
    name = "t3_clean"
    g, m, steps, seed, knobs, cfg = G.TRACES[name]
    batches = trace_gen.generate(lambda gg, mm: Oracle(gg, mm, **cfg), g, m, steps, seed, **knobs)
    assert G.pack_batches(batches) == open(os.path.join(HERE, "golden", name + ".events.z"), "rb").read()


FLOODS = json.load(open(os.path.join(HERE, "golden", "floods.json")))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("case", FLOODS, ids=lambda c: "%dx%d" % (c["groups"], c["members"]))
def test_golden_flood(case, backend):
    """Device-routed floods (mailboxes + host model inside the engine) against the committed sha256 of
    every member's final ra_row_state and the counters, up to BASELINE.json's 100k x 5."""
    g, m = case["groups"], case["members"]
    b = make_backend(backend, g, m, route_on_device=True)
    kw = dict(threads=4) if backend == "oracle" else {}
    rows, counters = G.flood_digest(b, g, m, case["steps"], case["cmds"], case["permille"], case["seed"], **kw)
    assert {k: counters[k] for k in case["counters"]} == case["counters"]
    assert rows == case["rows_sha256"]
