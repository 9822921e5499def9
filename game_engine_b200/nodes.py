"""Drop-in node callables for the reference's agent graph.

The reference registers `BotBehaviorNode`, `PhaseNode`, `RefereeNode` as LangGraph nodes with the signature
`async def Node(state, config) -> Command` (reference agent/game_agent_v2.py:468, 987, 619; registration
:1574-1580).  `GpuReferee` provides callables with the same names, signature, `goto` targets and update keys
whose decisions come from the CUDA step kernel instead of three LLM calls:

    ref = GpuReferee("werewolf-(mafia)", n_players=8, seed=7, session_id=42)
    workflow.add_node("BotBehaviorNode", ref.BotBehaviorNode)
    workflow.add_node("PhaseNode", ref.PhaseNode)
    workflow.add_node("RefereeNode", ref.RefereeNode)

Each node is stateless: it rebuilds the packed record of the session from the incoming dict, runs ONE step on
the GPU and returns only the keys its reference counterpart returns, so the three nodes can sit in the
reference graph unchanged (three tiny launches per graph run instead of three LLM round-trips).
`step_session(state)` is the fused form: the union of the three updates.  `ActionExecutorV3` serves the newer
graph variant (agent/game_agent_v3.py), whose single backend node merges the phase and referee decisions.

Error behaviour follows the reference: a node never raises into the graph; on failure it degrades to
"no change / stay at phase" and logs.  The batch API (`SessionBatch`) raises instead.
"""
from __future__ import annotations

import asyncio
import logging
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np

from . import table as T
from .adapter import SessionCodec
from .batch import SessionBatch, Table
from .compiler import compile_game

logger = logging.getLogger("game_engine_b200.nodes")

try:                                    # real LangGraph when it is installed
    from langgraph.types import Command  # type: ignore
except Exception:                       # same two attributes the reference's nodes use
    @dataclass
    class Command:                      # type: ignore[no-redef]
        goto: Any = None
        update: Dict[str, Any] = field(default_factory=dict)


class GpuReferee:
    def __init__(self, game: str, n_players: int, seed: int = 0, session_id: int = 0, device: int = 0,
                 kernel: str = "auto"):
        self.cg = compile_game(game, n_players)
        self.codec = SessionCodec(self.cg)
        self.table = Table(self.cg)
        self.seed, self.session_id = int(seed), int(session_id)
        self.batch = SessionBatch(self.table, 1, first_session_id=self.session_id, seed=self.seed, device=device, kernel=kernel)

    # ---- the GPU step on one dict-described session
    def _gpu_step(self, before: np.ndarray) -> np.ndarray:
        # one C call: record in -> one step -> record out (ge_run_host), a single synchronisation
        rec_in = np.ascontiguousarray(before.reshape(1, -1), dtype=np.uint8)
        rec_out = np.empty_like(rec_in)
        self.batch.run_host(rec_in, rec_out, 1)
        return rec_out[0]

    def step_session(self, state: Dict[str, Any], now_ms: Optional[int] = None, now_iso: Optional[str] = None) -> Dict[str, Any]:
        """L1 adapter: AgentState dict -> union of the update dicts of the three hot-path nodes."""
        before = self.codec.record_from_state(state)
        after = self._gpu_step(before)
        upd = self.codec.step_update(state, before, after, now_ms=now_ms, now_iso=now_iso)
        upd["dsl"] = state.get("dsl", self.cg.dsl)
        upd["roomSession"] = state.get("roomSession", {})
        return upd

    def initial_state(self, room_players=None) -> Dict[str, Any]:
        return self.codec.initial_state(room_players)

    # ---- node-compatible callables (same goto targets and update keys as the reference)
    async def BotBehaviorNode(self, state: Dict[str, Any], config: Any = None) -> Command:
        try:
            upd = await asyncio.to_thread(self.step_session, state)
            actions = upd["playerActions"]
        except Exception as e:      # reference: blanket except, no change (game_agent_v2.py:566-567)
            logger.error("[BotBehaviorNode] GPU step failed, no actions recorded: %s", e)
            actions = dict(state.get("playerActions", {}))
        return Command(goto="PhaseNode", update={
            "player_states": dict(state.get("player_states", {})), "playerActions": actions,
            "roomSession": state.get("roomSession", {}), "dsl": state.get("dsl", {})})

    async def PhaseNode(self, state: Dict[str, Any], config: Any = None) -> Command:
        hist = list(state.get("phase_history", []))
        cur = state.get("current_phase_id", 0)
        try:
            upd = await asyncio.to_thread(self.step_session, state)
            first_visit = cur == 0 and not any(e.get("phase_id") == 0 for e in hist)
            out = {"current_phase_id": upd["current_phase_id"], "player_states": state.get("player_states", {}),
                   "roomSession": state.get("roomSession", {}), "dsl": state.get("dsl", {}), "phase_history": upd["phase_history"]}
            if first_visit:         # reference :1043-1052 skips the referee on the phase-0 first visit
                return Command(goto="ActionExecutor", update=out)
            out["current_phase_name"] = upd["current_phase_name"]
            return Command(goto="RefereeNode", update=out)
        except Exception as e:      # reference: invalid decision -> stay at phase (:1165-1170, :1201-1204)
            logger.error("[PhaseNode] GPU step failed, staying at phase %s: %s", cur, e)
            return Command(goto="RefereeNode", update={"current_phase_id": cur, "phase_history": hist})

    async def RefereeNode(self, state: Dict[str, Any], config: Any = None) -> Command:
        """Runs after PhaseNode: `current_phase_id` is already the NEW phase and `phase_history` already has it
        appended (reference :673-681 recovers the phase just left from phase_history[-2]); rebuild the pre-step
        view, step, and return the referee's keys."""
        try:
            hist = list(state.get("phase_history", []))
            if len(hist) < 2:
                raise ValueError("RefereeNode needs phase_history[-2]")
            pre = dict(state)
            pre["current_phase_id"] = hist[-2]["phase_id"]
            pre["phase_history"] = hist[:-1]
            upd = await asyncio.to_thread(self.step_session, pre)
            ps, notes = upd["player_states"], upd["game_notes"]
        except Exception as e:
            logger.error("[RefereeNode] GPU step failed, no state change: %s", e)
            ps, notes = state.get("player_states", {}), list(state.get("game_notes", []))
        return Command(goto="ActionExecutor", update={
            "player_states": ps, "game_notes": notes, "roomSession": state.get("roomSession", {}),
            "dsl": state.get("dsl", {}), "phase_history": state.get("phase_history", [])})


    async def ActionExecutorV3(self, state: Dict[str, Any], config: Any = None) -> Command:
        """Drop-in for the backend half of the NEWER graph's merged node (reference agent/game_agent_v3.py:540-874:
        one LLM call bound to `update_complete_player_states` + `set_next_phase`): same phase-0 rule (:615-633), same
        update keys and `goto` (:862-874), history appended only when the phase changes and without a timestamp
        (:838-849).  Runs after BotBehaviorNode, which already recorded the bots' actions, so `playerActions` passes
        through unchanged; v3 does not return `game_notes`."""
        hist = list(state.get("phase_history", []))
        cur = state.get("current_phase_id", 0)
        try:
            upd = await asyncio.to_thread(self.step_session, state)
            if cur == 0 and not any(e.get("phase_id") == 0 for e in hist):
                return Command(goto="UIUpdateNode", update={"current_phase_id": 0, "phase_history": upd["phase_history"]})
            new_id = upd["current_phase_id"]
            if new_id != cur:
                hist.append({"phase_id": new_id, "phase_name": upd["current_phase_name"]})
            return Command(goto="UIUpdateNode", update={
                "player_states": upd["player_states"], "playerActions": dict(state.get("playerActions", {})),
                "phase_history": hist, "current_phase_id": new_id, "current_phase_name": upd["current_phase_name"]})
        except Exception as e:      # reference: invalid phase id -> keep the current phase (:833-836)
            logger.error("[ActionExecutorV3] GPU step failed, staying at phase %s: %s", cur, e)
            return Command(goto="UIUpdateNode", update={
                "player_states": state.get("player_states", {}), "playerActions": dict(state.get("playerActions", {})),
                "phase_history": hist, "current_phase_id": cur, "current_phase_name": state.get("current_phase_name", "")})


def terminal(cg, state: Dict[str, Any]) -> bool:
    return cg.table.phases[cg.index_of(state.get("current_phase_id", 0))].kind == T.KIND_TERMINAL
