"""Drives the reference's real nodes (InitialRouterNode -> BotBehaviorNode -> PhaseNode -> RefereeNode,
reference agent/game_agent_v2.py:198, 468, 987, 619) with the stub LLM, one graph run per step, and records
the state after every step.  ActionExecutor (UI rendering by LLM) is out of scope and not run; its only
state effect on the hot path is none.  Test infrastructure only."""
from __future__ import annotations

import asyncio
import copy
import os
from typing import Any, Dict, List

import yaml

from . import shims
from .stub_llm import StubChatModel

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
KEEP = ("current_phase_id", "current_phase_name", "player_states", "playerActions", "phase_history", "game_notes")


# Games the reference's own loader (agent/tools/utils.py:557-581: games/<gameName>.yaml) finds under another name:
# its earlier werewolf generation lives in game_draft/, reachable through a relative gameName.
REFERENCE_GAME_NAME = {
    "werewolf-draft": "../game_draft/werewolf-(mafia)",
    # the repo's extended game (tie -> re-vote, BASELINE config 4), loaded by the reference's loader from this repo
    "werewolf-revote": os.path.relpath(os.path.join(REPO, "game_engine_b200", "games", "werewolf-revote"),
                                       os.path.join(shims.REFERENCE_ROOT, "games")),
    # the repo's two-truths variant with numeric conditions (tools/make_handicap_variant.py)
    "two-truths-handicap": os.path.relpath(os.path.join(REPO, "game_engine_b200", "games", "two-truths-handicap"),
                                           os.path.join(shims.REFERENCE_ROOT, "games")),
}


def load_rules(game: str) -> dict:
    with open(os.path.join(REPO, "game_engine_b200", "rules", game + ".rules.yaml"), encoding="utf-8") as f:
        return yaml.safe_load(f)


def snapshot(state: Dict[str, Any]) -> Dict[str, Any]:
    """State with the nondeterministic parts (wall-clock timestamps) removed."""
    s = copy.deepcopy({k: state.get(k) for k in KEEP})
    for e in s.get("phase_history") or []:
        e.pop("timestamp", None)
    for pa in (s.get("playerActions") or {}).values():
        for a in (pa.get("actions") or {}).values():
            a.pop("timestamp", None)
    s["current_phase_name"] = s.get("current_phase_name") or ""
    return s


class HumanScript:
    """A person in seat 1 for the fixtures: what they type before every graph run.  `plan(step, state)` returns the
    message text; the default person first clicks Continue on a phase that needs them (so the phase waits once),
    then answers with the lowest-numbered legal choice, in the UI's vote message format."""

    seats = (1,)

    def __init__(self, rules: dict):
        self.rules = rules
        self.asked = set()

    def plan(self, step: int, state: Dict[str, Any]) -> str:
        from .stub_llm import holds
        ph = state.get("dsl", {}).get("phases", {})
        X = state.get("current_phase_id", 0)
        pr = (self.rules.get("phases") or {})
        act = (pr.get(X) or pr.get(str(X)) or {}).get("action")
        phase = ph.get(X) or ph.get(str(X)) or {}
        cc = phase.get("completion_criteria") or {}
        ps = state.get("player_states") or {}
        if step == 0 or not act or cc.get("type") != "player_action" or "1" not in ps:
            return "Continue"
        me = ps["1"]
        if not holds(cc["target_players"]["condition"], me):
            return "Continue"
        key = (len(state.get("phase_history") or []), X)
        if X not in self.asked:                       # first visit of this phase: keep the table waiting once
            self.asked.add(X)
            return "Continue"
        self.asked.discard(X)
        if act["op"] == "PICK_PLAYER":
            legal = [int(q) for q in sorted(ps, key=int) if holds(act["legal"], ps[q]) and not (act.get("exclude_self") and q == "1")]
            if not legal:
                return "Continue"
            return 'Player 1 voted "Player %d" in voting v%d' % (legal[0], key[0])
        if act["op"] == "PICK_OPTION":
            return "Player 1 chose statement %d" % (1 + key[0] % int(act["options"]))
        return "Player 1 submitted their statements"


async def _run(game: str, n_players: int, seed: int, sid: int, max_steps: int, human=None) -> List[Dict[str, Any]]:
    mod = shims.load_reference()
    stub = StubChatModel(load_rules(game), seed, sid, human_seats=human.seats if human else ())
    shims.set_model(stub)
    players = [{"name": "Player %d" % (i + 1), "gamePlayerId": str(i + 1)} for i in range(n_players)]
    state: Dict[str, Any] = {"gameName": REFERENCE_GAME_NAME.get(game, game), "roomSession": {"players": players}, "messages": [], "current_phase_id": 0,
                             "player_states": {}, "playerActions": {}, "phase_history": [], "game_notes": []}
    trace: List[Dict[str, Any]] = []
    messages: List[Any] = []
    for step in range(max_steps + 1):
        if human is not None:                                 # what the person typed before this graph run
            text = human.plan(step, state)
            state["messages"] = [shims.HumanMessage(content=text)]
        stub.state = state
        cmd = await mod.InitialRouterNode(state, {})          # loads the DSL, initialises player_states (first run)
        state.update(cmd.update)
        if step == 0:
            trace.append(snapshot(state))
        phase = state["dsl"]["phases"].get(state["current_phase_id"])
        if phase.get("next_phase") is None or step == max_steps:
            break
        assert cmd.goto == "BotBehaviorNode"
        stub.state = state
        cmd = await mod.BotBehaviorNode(state, {})
        state.update(cmd.update)
        assert cmd.goto == "PhaseNode"
        stub.state = state
        cmd = await mod.PhaseNode(state, {})
        state.update(cmd.update)
        if cmd.goto == "RefereeNode":
            stub.state = state
            cmd = await mod.RefereeNode(state, {})
            state.update(cmd.update)
        trace.append(snapshot(state))
        if human is not None:
            messages.append(text)
    if human is not None:
        trace[0]["_human_messages"] = messages                # messages[k] = what the person typed before step k + 1
    return trace


def run_session(game: str, n_players: int, seed: int, sid: int, max_steps: int = 400, human: bool = False) -> List[Dict[str, Any]]:
    """trace[k] = reference dict state after k steps (trace[0] = initial state).  human=True seats a scripted person
    (HumanScript) in seat 1; trace[0]["_human_messages"] then lists what they typed before every step."""
    return asyncio.run(_run(game, n_players, seed, sid, max_steps, HumanScript(load_rules(game)) if human else None))
