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


async def _run(game: str, n_players: int, seed: int, sid: int, max_steps: int) -> List[Dict[str, Any]]:
    mod = shims.load_reference()
    stub = StubChatModel(load_rules(game), seed, sid)
    shims.set_model(stub)
    players = [{"name": "Player %d" % (i + 1), "gamePlayerId": str(i + 1)} for i in range(n_players)]
    state: Dict[str, Any] = {"gameName": REFERENCE_GAME_NAME.get(game, game), "roomSession": {"players": players}, "messages": [], "current_phase_id": 0,
                             "player_states": {}, "playerActions": {}, "phase_history": [], "game_notes": []}
    trace: List[Dict[str, Any]] = []
    for step in range(max_steps + 1):
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
    return trace


def run_session(game: str, n_players: int, seed: int, sid: int, max_steps: int = 400) -> List[Dict[str, Any]]:
    """trace[k] = reference dict state after k steps (trace[0] = initial state)."""
    return asyncio.run(_run(game, n_players, seed, sid, max_steps))
