"""A person in seat 1 (SPEC D3h), pinned by fixtures from the reference's REAL nodes: its router logs the person's
messages (process_human_action_if_needed), its PhaseNode stays — appending history — while the stub model waits for
the person's answer, its RefereeNode applies nothing on a stay (tests/golden/human/, oracle/ref_harness/gen_golden.py
--human).  Oracle B + the host adapter, and on the GPU the drop-in nodes, must reproduce the dict state of every step."""
import glob
import gzip
import json
import os

import numpy as np
import pytest

from conftest import has_reference
from helpers import first_diff, normalise, oracle_step_fn, play_with_human

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "human", "*.json.gz")))
IDS = [os.path.basename(p)[:-8] for p in GOLDEN]


def load(path):
    with gzip.open(path, "rt", encoding="utf-8") as f:
        return json.load(f)


def _same(want, got):
    got = json.loads(json.dumps(got))
    assert len(want) == len(got)
    for k, (w, g) in enumerate(zip(want, got)):
        d = first_diff(w, g)
        assert d is None, "step %d: %s" % (k, d)


def test_fixtures_exist_and_people_kept_the_table_waiting():
    assert len(GOLDEN) >= 6
    for p in GOLDEN:
        g = load(p)
        ids = [s["current_phase_id"] for s in g["trace"]]
        hist = [len(s["phase_history"]) for s in g["trace"]]
        assert hist == list(range(len(hist)))                                   # one history entry per graph run, stays included
        assert any(a == b for a, b in zip(ids[2:], ids[3:]))                    # some phase was visited twice in a row
        assert "1" in g["trace"][-1]["playerActions"]                           # the router logged the person's messages


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_b_and_adapter_match_the_reference_nodes(path, games, oracle_for):
    g = load(path)
    cg = games(g["game"], g["players"])
    recs, mine = play_with_human(cg, g["human_messages"], oracle_step_fn(oracle_for(cg), g["sid"], g["seed"]), g["human_seats"])
    _same(g["trace"], mine)
    assert cg.table.phases[recs[-1][0]].kind == 3


@pytest.mark.reference
@pytest.mark.skipif(not has_reference(), reason="needs /root/reference")
@pytest.mark.parametrize("path", GOLDEN[:2] + GOLDEN[-1:], ids=lambda p: os.path.basename(p)[:-8])
def test_fixtures_reproduce_from_the_live_reference(path):
    from oracle.ref_harness.driver import run_session
    g = load(path)
    live = json.loads(json.dumps(run_session(g["game"], g["players"], g["seed"], g["sid"], human=True)))
    assert live[0].pop("_human_messages") == g["human_messages"]
    _same(g["trace"], live)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_gpu_step_session_with_a_person(path):
    """GpuReferee.step_session: the person's seat and input come from the dict (last human message), the step runs
    on the GPU (run-time-table kernel with the human-seat path)."""
    from game_engine_b200.nodes import GpuReferee
    g = load(path)
    ref = GpuReferee(g["game"], g["players"], seed=g["seed"], session_id=g["sid"], human_seats=g["human_seats"])
    state = ref.initial_state()
    trace = [normalise(state)]
    for text in g["human_messages"]:
        state["messages"] = [{"type": "human", "content": text}]
        state["playerActions"] = ref.codec.log_human_action(state, text, now_ms=0)
        state.update(ref.step_session(state, now_ms=0, now_iso=""))
        trace.append(normalise(state))
    _same(g["trace"], trace)
