"""Trace export / replay of simulated sessions.

`SessionBatch.trace` (C ABI `ge_trace`) returns the canonical records of a window of sessions after every
session-phase-step.  This module turns one session's frames into the reference's ON-WIRE state objects — one per
graph run, i.e. what the browser's `useCoAgent` state holds after the agent answered (reference
src/lib/canvas/types.ts:338-360 `AgentState`; server-side twin agent/game_agent_v2.py:97-117) — so a simulated
game can be stepped through in the reference UI, written as JSON lines, read back and re-verified on the GPU.

All rule evaluation stays in the CUDA step kernel: the frames are its outputs, the host only formats them
(`SessionCodec.step_update`, the same code the drop-in nodes use).
"""
from __future__ import annotations

import json
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

from . import table as T
from .adapter import Record, SessionCodec
from .compiler import CompiledGame

# keys of the TS AgentState the simulator has content for; the UI-only ones are emitted empty
WIRE_KEYS = ("items", "itemsCreated", "vote", "deadPlayers", "gameName", "current_phase_id", "current_phase_name",
             "player_states", "roomSession", "playerActions", "phase_history", "game_notes", "stateVersion",
             "stateTimestamp", "updatedBy")


def _wire_actions(actions: Dict[str, Any]) -> Dict[str, Any]:
    """The server keeps {pid: {name, actions: {id: {action, timestamp, phase, id}}}} (backend_tools.py:316-341);
    it travels unchanged — the TS type's flat form is the legacy shape (test_player_actions.py:53-58)."""
    return {pid: {"name": v.get("name"), "actions": {k: dict(a) for k, a in v.get("actions", {}).items()}} for pid, v in actions.items()}


def materialise(cg: CompiledGame, frames: np.ndarray, room_players: Optional[List[dict]] = None,
                t0_ms: int = 0, step_ms: int = 1000) -> List[Dict[str, Any]]:
    """frames: uint8[K + 1, S] = one session's canonical records after 0..K steps (frame 0 must be the initial
    record).  Returns K + 1 on-wire state objects; object k is the state after k graph runs.  Timestamps are
    synthetic and deterministic: t0_ms + k * step_ms."""
    codec = SessionCodec(cg)
    frames = np.asarray(frames, dtype=np.uint8)
    assert frames.ndim == 2 and frames.shape[1] == cg.record_size
    if not np.array_equal(frames[0], codec.initial_record()):
        raise ValueError("frame 0 is not the initial record: a trace must start at session creation")
    state = codec.initial_state(room_players)
    out = [wire_state(cg, state, frames[0], version=0, ts_ms=t0_ms)]
    votes: List[Dict[str, str]] = []
    for k in range(1, frames.shape[0]):
        ts = t0_ms + k * step_ms
        upd = codec.step_update(state, frames[k - 1], frames[k], now_ms=ts, now_iso=_iso(ts))
        state.update(upd)
        b = Record(cg, frames[k - 1])
        a = Record(cg, frames[k])
        if a.step != b.step and b.step > 0:
            ph = cg.table.phases[b.phase]
            if ph.exit_op in (T.EX_DAY_VOTE, T.EX_T_VOTES):          # VoteRecord[] (types.ts:312-316)
                actors = b.eval_pred(int(ph.actor_pred))
                vid = "phase-%d-step-%d" % (cg.phase_ids[b.phase], b.step)
                for p in range(cg.n_players):
                    if (actors >> p) & 1:
                        opt = a.target[p] if cg.family == T.FAMILY_WEREWOLF else a.vote[p]
                        votes.append({"voteid": vid, "playerid": str(p + 1), "option": str(opt)})
        out.append(wire_state(cg, state, frames[k], version=k, ts_ms=ts, votes=votes))
    return out


def wire_state(cg: CompiledGame, state: Dict[str, Any], record: np.ndarray, version: int, ts_ms: int,
               votes: Optional[List[Dict[str, str]]] = None) -> Dict[str, Any]:
    r = Record(cg, record)
    dead = []
    if cg.family == T.FAMILY_WEREWOLF:
        dead = [str(p + 1) for p in range(cg.n_players) if not (r.alive >> p) & 1]
    return {
        "items": [], "itemsCreated": 0, "vote": [dict(v) for v in (votes or [])], "deadPlayers": dead,
        "gameName": cg.name, "current_phase_id": int(state["current_phase_id"]),
        "current_phase_name": state.get("current_phase_name") or "",
        "player_states": json.loads(json.dumps(state["player_states"])),
        "roomSession": json.loads(json.dumps(state.get("roomSession") or {})),
        "playerActions": _wire_actions(state.get("playerActions") or {}),
        "phase_history": [dict(e) for e in state.get("phase_history") or []],
        "game_notes": list(state.get("game_notes") or []),
        "stateVersion": int(version), "stateTimestamp": ts_ms / 1000.0, "updatedBy": "game_engine_b200",
        "record": bytes(np.asarray(record, dtype=np.uint8)).hex(),          # lets a trace be re-verified / resumed
    }


def _iso(ts_ms: int) -> str:
    import datetime as _dt
    return _dt.datetime.fromtimestamp(ts_ms / 1000.0, _dt.timezone.utc).replace(tzinfo=None).isoformat()


def write_jsonl(path: str, header: Dict[str, Any], states: Iterable[Dict[str, Any]]) -> None:
    """Line 1 = header {game, players, seed, session_id, ...}; then one state object per line."""
    with open(path, "w") as f:
        f.write(json.dumps({"trace": "game_engine_b200/1", **header}) + "\n")
        for s in states:
            f.write(json.dumps(s) + "\n")


def read_jsonl(path: str):
    with open(path) as f:
        lines = [json.loads(x) for x in f if x.strip()]
    return lines[0], lines[1:]


def frames_of(states: List[Dict[str, Any]]) -> np.ndarray:
    return np.stack([np.frombuffer(bytes.fromhex(s["record"]), dtype=np.uint8) for s in states])


def export_session(game: str, players: int, seed: int, session_id: int, n_steps: Optional[int] = None, device: int = 0,
                   kernel: str = "auto", path: Optional[str] = None) -> List[Dict[str, Any]]:
    """Simulates one session on the GPU from creation (to its end by default) and returns / writes its trace."""
    from .batch import SessionBatch, Table
    from .compiler import compile_game
    cg = compile_game(game, players)
    tab = Table(cg)
    b = SessionBatch(tab, 1, first_session_id=session_id, seed=seed, device=device, kernel=kernel)
    try:
        cap = n_steps if n_steps is not None else max_game_steps(cg)
        frames = b.trace(cap)[:, 0, :]
    finally:
        b.close()
    if n_steps is None:                                    # drop the no-op frames after the game has ended
        steps = frames[:, 2].astype(np.int64) | (frames[:, 3].astype(np.int64) << 8)
        frames = frames[: int(steps.max()) + 1]
    states = materialise(cg, frames)
    if path:
        write_jsonl(path, {"game": game, "players": players, "seed": seed, "session_id": session_id,
                           "steps": len(states) - 1}, states)
    return states


def max_game_steps(cg: CompiledGame) -> int:
    """Longest possible game in steps (SPEC.md): werewolf 9P - 16 (+ 2 per re-vote per day), TTL 2 + 8P."""
    P = cg.n_players
    if cg.family == T.FAMILY_WEREWOLF:
        return 9 * P - 16 + 2 * cg.table.max_revotes * (P - 2)
    return 2 + 8 * P * max(1, cg.table.rounds)


def replay_check(cg: CompiledGame, seed: int, session_id: int, states: List[Dict[str, Any]], device: int = 0,
                 kernel: str = "auto", start: int = 0) -> int:
    """Re-verifies a trace on the GPU: imports the record of state `start`, steps, and compares every later
    record bit for bit.  Returns the number of steps verified; raises ValueError at the first mismatch."""
    from .batch import SessionBatch, Table
    frames = frames_of(states)
    tab = Table(cg)
    b = SessionBatch(tab, 1, first_session_id=session_id, seed=seed, device=device, kernel=kernel)
    try:
        b.import_state(frames[start:start + 1])
        got = b.trace(len(states) - 1 - start)[:, 0, :]
    finally:
        b.close()
    for k in range(got.shape[0]):
        if not np.array_equal(got[k], frames[start + k]):
            raise ValueError("trace diverges from the simulator at state %d" % (start + k))
    return got.shape[0] - 1
