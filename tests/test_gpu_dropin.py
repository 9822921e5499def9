"""GPU path through the reference-facing host layer, checked against the fixtures produced by the reference's
real nodes (tests/golden): the batch path + adapter, the fused step_session(), and the three drop-in nodes."""
import asyncio
import glob
import gzip
import json
import os

import pytest

from helpers import first_diff, normalise, pick, replay_records

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json.gz")))
IDS = [os.path.basename(p)[:-8] for p in GOLDEN]


def load(path):
    with gzip.open(path, "rt", encoding="utf-8") as f:
        return json.load(f)


def _assert_traces_equal(want, got):
    got = json.loads(json.dumps(got))
    assert len(want) == len(got)
    for k, (w, g) in enumerate(zip(want, got)):
        d = first_diff(w, g)
        assert d is None, "step %d: %s" % (k, d)


@pytest.mark.parametrize("kernel", ["tps", "coop"])
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_gpu_records_expand_to_the_reference_trace(path, kernel, games):
    from game_engine_b200.batch import SessionBatch, Table
    g = load(path)
    cg = games(g["game"], g["players"])
    b = SessionBatch(Table(cg), 1, first_session_id=g["sid"], seed=g["seed"], kernel=kernel)
    recs = [b.export_state()[0].copy()]
    for _ in range(len(g["trace"]) - 1):
        b.step(1)
        recs.append(b.export_state()[0].copy())
    _assert_traces_equal(g["trace"], replay_records(cg, recs))


@pytest.mark.parametrize("path", pick(GOLDEN, "lie_p3_", "lie_p4_seed0_", "lie_p5_", "draft_p6_", "draft_p8_", "mafia_p4_", "mafia_p8_seed1_", "mafia_p16_",
                                             "revote_p8_seed42_", "revote_p8_seed46_", "revote_p16_"),
                         ids=lambda p: os.path.basename(p)[:-8])
def test_step_session_chain(path):
    """L1 adapter: dict in -> dict out, every step rebuilt from the dict alone (no hidden device state)."""
    from game_engine_b200.nodes import GpuReferee
    g = load(path)
    ref = GpuReferee(g["game"], g["players"], seed=g["seed"], session_id=g["sid"])
    state = ref.initial_state()
    trace = [normalise(state)]
    for _ in range(len(g["trace"]) - 1):
        state.update(ref.step_session(state, now_ms=0, now_iso=""))
        trace.append(normalise(state))
    _assert_traces_equal(g["trace"], trace)
    # one more step on the terminal session changes nothing
    again = ref.step_session(state, now_ms=0, now_iso="")
    assert first_diff(normalise({**state, **again}), trace[-1]) is None


@pytest.mark.parametrize("path", pick(GOLDEN, "lie_p4_seed0_", "draft_p12_", "mafia_p8_seed7_", "revote_p8_seed46_"), ids=lambda p: os.path.basename(p)[:-8])
def test_drop_in_nodes_follow_the_reference_graph(path):
    """BotBehaviorNode -> PhaseNode -> (RefereeNode | ActionExecutor) with the reference's goto targets and
    update keys (reference agent/game_agent_v2.py:609-617, 1043-1052, 1238-1241, 794-803)."""
    from game_engine_b200.nodes import GpuReferee, terminal
    g = load(path)
    ref = GpuReferee(g["game"], g["players"], seed=g["seed"], session_id=g["sid"])

    async def run():
        state = ref.initial_state()
        trace = [normalise(state)]
        for k in range(len(g["trace"]) - 1):
            cmd = await ref.BotBehaviorNode(state, {})
            assert cmd.goto == "PhaseNode" and set(cmd.update) == {"player_states", "playerActions", "roomSession", "dsl"}
            state.update(cmd.update)
            cmd = await ref.PhaseNode(state, {})
            state.update(cmd.update)
            if k == 0:
                assert cmd.goto == "ActionExecutor" and "current_phase_name" not in cmd.update
            else:
                assert cmd.goto == "RefereeNode"
                cmd = await ref.RefereeNode(state, {})
                assert cmd.goto == "ActionExecutor"
                assert set(cmd.update) == {"player_states", "game_notes", "roomSession", "dsl", "phase_history"}
                state.update(cmd.update)
            trace.append(normalise(state))
        assert terminal(ref.cg, state)
        return trace

    _assert_traces_equal(g["trace"], asyncio.run(run()))


@pytest.mark.parametrize("path", pick(GOLDEN, "lie_p4_seed1_", "draft_p8_", "mafia_p12_", "revote_p5_"), ids=lambda p: os.path.basename(p)[:-8])
def test_v3_merged_node_follows_the_reference(path):
    """BotBehaviorNode -> ActionExecutorV3 (the newer graph, reference agent/game_agent_v3.py:1099-1117): same
    player_states / playerActions / phase ids as the v2 fixtures; history entries carry no timestamp and are only
    appended when the phase changes; game_notes are not part of v3's update."""
    from game_engine_b200.nodes import GpuReferee, terminal
    g = load(path)
    ref = GpuReferee(g["game"], g["players"], seed=g["seed"], session_id=g["sid"])

    async def run():
        state = ref.initial_state()
        for k in range(len(g["trace"]) - 1):
            cmd = await ref.BotBehaviorNode(state, {})
            state.update(cmd.update)
            cmd = await ref.ActionExecutorV3(state, {})
            assert cmd.goto == "UIUpdateNode"
            if k == 0:
                assert set(cmd.update) == {"current_phase_id", "phase_history"}
            else:
                assert set(cmd.update) == {"player_states", "playerActions", "phase_history", "current_phase_id", "current_phase_name"}
                assert "timestamp" not in cmd.update["phase_history"][-1]
            state.update(cmd.update)
            want = g["trace"][k + 1]
            got = json.loads(json.dumps(normalise(state)))
            for key in ("current_phase_id", "player_states", "playerActions", "phase_history"):
                d = first_diff(want[key], got[key], "/" + key)
                assert d is None, "step %d: %s" % (k + 1, d)
        assert terminal(ref.cg, state) and state["game_notes"] == []
    asyncio.run(run())
