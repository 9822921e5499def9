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


class _HotPathNodes:
    """The node callables, written once against `await self._step(state, config)` -> the union of the three nodes'
    updates for one graph run.  GpuReferee steps its own one-session batch; RefereePool coalesces the concurrent
    graph runs of many rooms into one device call."""

    async def _step(self, state: Dict[str, Any], config: Any) -> Dict[str, Any]:
        raise NotImplementedError

    # ---- node-compatible callables (same goto targets and update keys as the reference)
    async def BotBehaviorNode(self, state: Dict[str, Any], config: Any = None) -> Command:
        try:
            upd = await self._step(state, config)
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
            upd = await self._step(state, config)
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
            upd = await self._step(pre, config)
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
            upd = await self._step(state, config)
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


class GpuReferee(_HotPathNodes):
    """One room.  `human_seats`: seats played by people (the reference's room: (1,)); their input for a graph run is
    read from the last human message of the state (SPEC D3h, adapter.human_inputs) and a phase that needs it waits."""

    def __init__(self, game: str, n_players: int, seed: int = 0, session_id: int = 0, device: int = 0,
                 kernel: str = "auto", human_seats=()):
        self.cg = compile_game(game, n_players)
        self.codec = SessionCodec(self.cg)
        self.table = Table(self.cg)
        self.seed, self.session_id = int(seed), int(session_id)
        self.human_seats = tuple(int(x) for x in human_seats)
        self.batch = SessionBatch(self.table, 1, first_session_id=self.session_id, seed=self.seed, device=device, kernel=kernel)
        if self.human_seats:
            mask = sum(1 << (x - 1) for x in self.human_seats)
            self.batch.set_human_seats(np.array([mask], dtype=np.uint32))

    # ---- the GPU step on one dict-described session
    def _gpu_step(self, before: np.ndarray, inputs: Optional[np.ndarray] = None) -> np.ndarray:
        # one C call: record in -> one step -> record out (ge_run_host), a single synchronisation
        rec_in = np.ascontiguousarray(before.reshape(1, -1), dtype=np.uint8)
        rec_out = np.empty_like(rec_in)
        if inputs is not None:
            self.batch.set_human_choices(inputs.reshape(1, -1))
        self.batch.run_host(rec_in, rec_out, 1)
        return rec_out[0]

    def step_session(self, state: Dict[str, Any], now_ms: Optional[int] = None, now_iso: Optional[str] = None) -> Dict[str, Any]:
        """L1 adapter: AgentState dict -> union of the update dicts of the three hot-path nodes."""
        before = self.codec.record_from_state(state)
        mask, row = self.codec.human_inputs(state, self.human_seats) if self.human_seats else (0, None)
        after = self._gpu_step(before, row)
        upd = self.codec.step_update(state, before, after, now_ms=now_ms, now_iso=now_iso, human_mask=mask)
        upd["dsl"] = state.get("dsl", self.cg.dsl)
        upd["roomSession"] = state.get("roomSession", {})
        return upd

    def tool_calls(self, state: Dict[str, Any]) -> Dict[str, list]:
        """The same step as the tool-call lists the reference's nodes would apply (adapter.tool_calls_for)."""
        before = self.codec.record_from_state(state)
        mask, row = self.codec.human_inputs(state, self.human_seats) if self.human_seats else (0, None)
        after = self._gpu_step(before, row)
        return self.codec.tool_calls_for(state, before, after, human_mask=mask)

    def initial_state(self, room_players=None) -> Dict[str, Any]:
        return self.codec.initial_state(room_players)

    async def _step(self, state: Dict[str, Any], config: Any) -> Dict[str, Any]:
        return await asyncio.to_thread(self.step_session, state)


class RefereePool(_HotPathNodes):
    """Many rooms behind the node API, one device batch (the reference serves many LangGraph threads per process,
    src/app/api/copilotkit/route.ts:24-38): every room owns a SLOT of the batch (its session id is
    first_session_id + slot) and the graph runs of different rooms that arrive within `max_delay_ms` — or as soon as
    `max_batch` are waiting — are stepped by ONE host-buffer call (records in, one step, records out).  Slots of rooms
    that are not part of a call hold a parked terminal record, so they do not move.

        pool = RefereePool("werewolf-(mafia)", 8, capacity=4096, seed=7)
        workflow.add_node("BotBehaviorNode", pool.BotBehaviorNode)      # rooms are told apart by config thread_id
        ...
        updates = pool.step_sessions([(slot_a, state_a), (slot_b, state_b)])     # or the synchronous batch form
    """

    def __init__(self, game: str, n_players: int, capacity: int = 1024, seed: int = 0, first_session_id: int = 0,
                 device: int = 0, human_seats=(), max_batch: int = 256, max_delay_ms: float = 2.0):
        self.cg = compile_game(game, n_players)
        self.codec = SessionCodec(self.cg)
        self.table = Table(self.cg)
        self.capacity, self.seed, self.first_session_id = int(capacity), int(seed), int(first_session_id)
        self.human_seats = tuple(int(x) for x in human_seats)
        self.batch = SessionBatch(self.table, self.capacity, first_session_id=self.first_session_id, seed=self.seed, device=device, kernel="tps")
        self.batch.set_compaction(0)                     # slots must stay where they are
        self._mask = sum(1 << (x - 1) for x in self.human_seats)
        if self._mask:
            self.batch.set_human_seats(np.full(self.capacity, self._mask, dtype=np.uint32))
        # parked record: a finished game (terminal phase) never steps
        term = next(i for i, ph in enumerate(self.cg.table.phases) if ph.kind == T.KIND_TERMINAL)
        parked = self.codec.initial_record().copy()
        parked[0], parked[2] = term, 1
        self._parked = np.tile(parked, (self.capacity, 1))
        self._out = np.empty_like(self._parked)
        self._slots: Dict[Any, int] = {}
        self._free = list(range(self.capacity - 1, -1, -1))
        self.max_batch, self.max_delay = int(max_batch), float(max_delay_ms) / 1e3
        self._pending: list = []
        self._flusher: Optional[asyncio.Task] = None
        self.calls = 0                                   # device calls made (for the tests / metrics)

    # ---- rooms
    def open_room(self, room_id: Any = None) -> int:
        if room_id is not None and room_id in self._slots:
            return self._slots[room_id]
        if not self._free:
            raise RuntimeError("RefereePool is full (%d rooms)" % self.capacity)
        slot = self._free.pop()
        self._slots[room_id if room_id is not None else ("slot", slot)] = slot
        return slot

    def close_room(self, room_id: Any) -> None:
        slot = self._slots.pop(room_id, None)
        if slot is not None:
            self._free.append(slot)

    def session_id(self, slot: int) -> int:
        return self.first_session_id + int(slot)

    def initial_state(self, room_players=None) -> Dict[str, Any]:
        return self.codec.initial_state(room_players)

    # ---- one device call for many rooms
    def step_sessions(self, items, now_ms: Optional[int] = None, now_iso: Optional[str] = None):
        """items: [(slot, AgentState dict)] with distinct slots -> [update dict] in the same order."""
        slots = [int(s) for s, _ in items]
        if len(set(slots)) != len(slots):
            raise ValueError("a room can take part in a call only once")
        rec_in = self._parked.copy()
        befores, masks = [], []
        choices = np.full((self.capacity, self.batch.human_stride), 0xFF, dtype=np.uint8) if self._mask else None
        for slot, state in items:
            before = self.codec.record_from_state(state)
            rec_in[slot] = before
            befores.append(before)
            if self._mask:
                m, row = self.codec.human_inputs(state, self.human_seats)
                choices[slot] = row
                masks.append(m)
            else:
                masks.append(0)
        if choices is not None:
            self.batch.set_human_choices(choices)
        self.batch.run_host(rec_in, self._out, 1)
        self.calls += 1
        out = []
        for (slot, state), before, m in zip(items, befores, masks):
            upd = self.codec.step_update(state, before, self._out[slot].copy(), now_ms=now_ms, now_iso=now_iso, human_mask=m)
            upd["dsl"] = state.get("dsl", self.cg.dsl)
            upd["roomSession"] = state.get("roomSession", {})
            out.append(upd)
        return out

    # ---- asynchronous micro-batching behind the node callables
    @staticmethod
    def _room_of(config: Any) -> Any:
        try:
            return (config or {}).get("configurable", {}).get("thread_id")
        except AttributeError:
            return None

    async def step_session(self, slot: int, state: Dict[str, Any]) -> Dict[str, Any]:
        """One room's graph run; resolved by the next device call, shared with whoever else is waiting."""
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        self._pending.append((int(slot), state, fut))
        if self._flusher is None or self._flusher.done():
            self._flusher = loop.create_task(self._flush_soon())
        return await fut

    async def _flush_soon(self) -> None:
        while self._pending:
            if len(self._pending) < self.max_batch:
                await asyncio.sleep(self.max_delay)
            take, seen, rest = [], set(), []
            for item in self._pending:
                if item[0] in seen or len(take) >= self.max_batch:
                    rest.append(item)                     # the same room twice: its second run goes to the next call
                else:
                    seen.add(item[0])
                    take.append(item)
            self._pending = rest
            try:
                res = await asyncio.to_thread(self.step_sessions, [(s, st) for s, st, _ in take])
                for (_, _, fut), upd in zip(take, res):
                    if not fut.done():
                        fut.set_result(upd)
            except Exception as e:                        # every waiting node degrades to "no change" (reference behaviour)
                for _, _, fut in take:
                    if not fut.done():
                        fut.set_exception(e)

    async def _step(self, state: Dict[str, Any], config: Any) -> Dict[str, Any]:
        room = self._room_of(config)
        if room is None:
            raise ValueError("RefereePool nodes need config['configurable']['thread_id'] to tell rooms apart")
        return await self.step_session(self.open_room(room), state)


def terminal(cg, state: Dict[str, Any]) -> bool:
    return cg.table.phases[cg.index_of(state.get("current_phase_id", 0))].kind == T.KIND_TERMINAL
