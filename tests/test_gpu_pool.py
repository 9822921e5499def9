"""RefereePool: many rooms behind the node API, one device batch (the reference serves many LangGraph threads per
process, src/app/api/copilotkit/route.ts:24-38).  Every room's dict trace must be what Oracle B + the adapter give for
its session id, however the rooms' graph runs are interleaved and coalesced."""
import asyncio
import json

import numpy as np
import pytest

from helpers import first_diff, normalise, oracle_b_records, oracle_step_fn, play_with_human, replay_records

pytestmark = pytest.mark.gpu


def _expected(cg, o, sid, seed, n_steps):
    recs = oracle_b_records(o, sid, seed, n_steps)
    return json.loads(json.dumps(replay_records(cg, recs)))


@pytest.mark.parametrize("game,P", [("werewolf-(mafia)", 8), ("two-truths-and-a-lie", 5), ("werewolf-revote", 6)])
def test_step_sessions_interleaves_rooms_freely(games, oracle_for, game, P):
    from game_engine_b200.nodes import RefereePool, terminal
    cg = games(game, P)
    o = oracle_for(cg)
    seed, first = 91, 5000
    pool = RefereePool(game, P, capacity=64, seed=seed, first_session_id=first)
    rng = np.random.default_rng(3)
    rooms = {r: pool.open_room("room-%d" % r) for r in range(40)}
    states = {r: pool.initial_state() for r in rooms}
    traces = {r: [normalise(states[r])] for r in rooms}
    for _ in range(400):
        live = [r for r in rooms if not terminal(cg, states[r])]
        if not live:
            break
        pick = [r for r in live if rng.random() < 0.5] or live[:1]
        for r, upd in zip(pick, pool.step_sessions([(rooms[r], states[r]) for r in pick], now_ms=0, now_iso="")):
            states[r].update(upd)
            traces[r].append(normalise(states[r]))
    assert all(terminal(cg, states[r]) for r in rooms)
    for r, slot in rooms.items():
        want = _expected(cg, o, pool.session_id(slot), seed, len(traces[r]) - 1)
        got = json.loads(json.dumps(traces[r]))
        for k, (w, g) in enumerate(zip(want, got)):
            d = first_diff(w, g)
            assert d is None, "room %d step %d: %s" % (r, k, d)
    with pytest.raises(ValueError):
        pool.step_sessions([(0, states[0]), (0, states[0])])


def test_concurrent_graph_runs_share_device_calls(games, oracle_for):
    """Thirty rooms play whole games through the pool's node callables at the same time (one asyncio task per room, as
    LangGraph runs threads); their graph runs are coalesced, and every room still gets its own game."""
    from game_engine_b200.nodes import RefereePool, terminal
    game, P, seed, first = "werewolf-(mafia)", 8, 17, 100
    cg = games(game, P)
    o = oracle_for(cg)
    pool = RefereePool(game, P, capacity=32, seed=seed, first_session_id=first, max_batch=64, max_delay_ms=1.0)

    async def room(thread_id):
        cfg = {"configurable": {"thread_id": thread_id}}
        state = pool.initial_state()
        trace = [normalise(state)]
        for k in range(200):
            if terminal(cg, state):
                break
            cmd = await pool.BotBehaviorNode(state, cfg)
            state.update(cmd.update)
            cmd = await pool.PhaseNode(state, cfg)
            state.update(cmd.update)
            if cmd.goto == "RefereeNode":
                cmd = await pool.RefereeNode(state, cfg)
                state.update(cmd.update)
            trace.append(normalise(state))
        return thread_id, trace

    async def main():
        return await asyncio.gather(*[room("thread-%d" % i) for i in range(30)])

    results = asyncio.run(main())
    node_calls = 0
    for thread_id, trace in results:
        slot = pool.open_room(thread_id)                 # the slot the room was given on its first graph run
        want = _expected(cg, o, pool.session_id(slot), seed, len(trace) - 1)
        got = json.loads(json.dumps(trace))
        for k, (w, g) in enumerate(zip(want, got)):
            d = first_diff(w, g)
            assert d is None, "%s step %d: %s" % (thread_id, k, d)
        node_calls += 3 * (len(trace) - 1)
    assert pool.calls * 5 < node_calls, (pool.calls, node_calls)      # rooms shared device calls


def test_a_pool_with_a_person_in_every_room(games, oracle_for):
    from game_engine_b200.nodes import RefereePool
    game, P, seed, first = "two-truths-and-a-lie", 4, 3, 9
    cg = games(game, P)
    o = oracle_for(cg)
    pool = RefereePool(game, P, capacity=8, seed=seed, first_session_id=first, human_seats=(1,))
    scripts = ["Continue", "Continue", "Continue", "Continue", "Player 1 submitted their statements", "Continue",
               "Player 1 chose statement 2", "Continue", "Continue", "Continue", "Continue", "Player 1 voted \"1\" in voting v", "Continue"] * 6
    for slot in (0, 5):
        sid = pool.session_id(slot)
        _, want = play_with_human(cg, scripts, oracle_step_fn(o, sid, seed))
        state = pool.initial_state()
        trace = [normalise(state)]
        for text in scripts:
            state["messages"] = [{"type": "human", "content": text}]
            state["playerActions"] = pool.codec.log_human_action(state, text, now_ms=0)
            state.update(pool.step_sessions([(slot, state)], now_ms=0, now_iso="")[0])
            trace.append(normalise(state))
        got = json.loads(json.dumps(trace))
        for k, (w, g) in enumerate(zip(json.loads(json.dumps(want)), got)):
            d = first_diff(w, g)
            assert d is None, "slot %d step %d: %s" % (slot, k, d)
