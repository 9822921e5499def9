"""The step as the reference's tool calls ("adapter can also emit the step as tool-call lists").

`SessionCodec.tool_calls_for(state, before, after)` turns one step of the packed record into the lists
`update_player_actions` / `set_next_phase` / `update_player_state` / `add_game_note` (reference
agent/tools/backend_tools.py:10-157).  Here those lists are handed to the reference's REAL BotBehaviorNode, PhaseNode and
RefereeNode through a pass-through chat model, so the reference's own `_execute_*` code applies them in order
(game_agent_v2.py:589-605, 1124-1215, 762-786) — and the state it arrives at must be the state `step_update` builds
directly.  Needs /root/reference (build container)."""
import asyncio
import json

import numpy as np
import pytest

from conftest import has_reference
from helpers import first_diff, normalise

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not has_reference(), reason="needs /root/reference")]


class PassThrough:
    """Chat model stand-in that answers every node with the tool calls prepared for it."""

    def __init__(self):
        self.calls = {"BotBehaviorNode": [], "PhaseNode": [], "RefereeNode": []}

    def bind_tools(self, tools, **kw):
        names = frozenset(t.name for t in tools)
        node = {frozenset({"update_player_actions"}): "BotBehaviorNode", frozenset({"set_next_phase"}): "PhaseNode",
                frozenset({"update_player_state", "add_game_note"}): "RefereeNode"}[names]
        return _Bound(self, node)


class _Bound:
    def __init__(self, model, node):
        self.model, self.node = model, node

    async def ainvoke(self, messages, config=None):
        from oracle.ref_harness.shims import AIMessage
        return AIMessage(content="", tool_calls=list(self.model.calls[self.node]))


CASES = [("werewolf-(mafia)", 8, 3, 11, False), ("werewolf-revote", 8, 46, 3010, False), ("two-truths-and-a-lie", 4, 1, 2, False),
         ("werewolf-draft", 6, 2, 3, False), ("werewolf-(mafia)", 6, 5, 77, True), ("two-truths-and-a-lie", 4, 3, 9, True)]


@pytest.mark.parametrize("game,P,seed,sid,human", CASES)
def test_reference_nodes_apply_our_tool_calls_to_the_same_state(games, oracle_for, game, P, seed, sid, human):
    from game_engine_b200.adapter import SessionCodec
    from oracle.ref_harness import shims
    from oracle.ref_harness.driver import REFERENCE_GAME_NAME, HumanScript, load_rules
    mod = shims.load_reference()
    model = PassThrough()
    shims.set_model(model)
    cg = games(game, P)
    o = oracle_for(cg)
    codec = SessionCodec(cg)
    script = HumanScript(load_rules(game)) if human else None
    seats = (1,) if human else ()

    async def run():
        players = [{"name": "Player %d" % (i + 1), "gamePlayerId": str(i + 1)} for i in range(P)]
        ref_state = {"gameName": REFERENCE_GAME_NAME.get(game, game), "roomSession": {"players": players}, "messages": [],
                     "current_phase_id": 0, "player_states": {}, "playerActions": {}, "phase_history": [], "game_notes": []}
        mine = codec.initial_state()
        checked = 0
        for step in range(400):
            if script is not None:
                text = script.plan(step, ref_state)
                ref_state["messages"] = [shims.HumanMessage(content=text)]
                mine["messages"] = [{"type": "human", "content": text}]
                mine["playerActions"] = codec.log_human_action(mine, text, now_ms=0)
            cmd = await mod.InitialRouterNode(ref_state, {})
            ref_state.update(cmd.update)
            if cg.table.phases[cg.index_of(ref_state["current_phase_id"])].kind == 3:
                break
            # our side: one step of Oracle B on the record of the dict state, expressed as tool calls
            before = codec.record_from_state(mine)
            mask, row = codec.human_inputs(mine, seats) if seats else (0, None)
            rec = np.array(before.reshape(1, -1), copy=True)
            if seats:
                o.step_humans(rec, sid, seed, np.array([mask], dtype=np.uint32), row.reshape(1, -1))
            else:
                o.step(rec, sid, seed, 1)
            model.calls = codec.tool_calls_for(mine, before, rec[0], human_mask=mask)
            mine.update(codec.step_update(mine, before, rec[0], now_ms=0, now_iso="", human_mask=mask))
            # the reference's nodes apply them
            cmd = await mod.BotBehaviorNode(ref_state, {})
            ref_state.update(cmd.update)
            cmd = await mod.PhaseNode(ref_state, {})
            ref_state.update(cmd.update)
            if cmd.goto == "RefereeNode":
                cmd = await mod.RefereeNode(ref_state, {})
                ref_state.update(cmd.update)
            d = first_diff(json.loads(json.dumps(normalise(ref_state))), json.loads(json.dumps(normalise(mine))))
            assert d is None, "step %d: %s" % (step + 1, d)
            checked += 1
        return checked

    assert asyncio.run(run()) > 10
