"""Host adapter between the reference's AgentState dict and the packed session record.

Mirrors, field for field, what the reference's nodes read and write (reference agent/game_agent_v2.py:97-117
AgentState; update dicts at :609-617, :794-803, :1222-1241; tool applications agent/tools/backend_tools.py:204-225,
285-344; session init agent/tools/utils.py:584-653).  All rule evaluation happens on the GPU; this module only
converts representations and writes the step's results in the reference's string formats:

* player ids are strings "1".."P", phase ids are the DSL's ints, `phase_history` entries are
  `{phase_id, phase_name, timestamp}` (game_agent_v2.py:1210-1215);
* `playerActions[pid] = {"name", "actions": {id: {"action", "timestamp", "phase", "id"}}}` with a per-player
  monotone id (backend_tools.py:316-341);
* `game_notes` entries are `"<emoji> <TYPE>: <content>"` (backend_tools.py:175-200).
"""
from __future__ import annotations

import datetime as _dt
import re as _re
import time as _time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import table as T
from .compiler import CompiledGame

NOTE_EMOJI = {
    "CRITICAL": "🔴", "VOTING_STATUS": "⚠️", "DECISION": "🎯", "BOT_REMINDER": "🤖", "UI_FILTER": "🚫",
    "PHASE_STATUS": "⏳", "NEXT_PHASE": "🔮", "GAME_STATUS": "🏆", "PHASE_SUGGESTION": "💡",
    "BRANCH_RECOMMENDATION": "🔀", "EVENT": "📝",
}
NO_TARGET_TEXT = "had no legal target"
TTL_STATEMENTS = ("I once met a celebrity.", "I can speak four languages.", "I've never broken a bone.")


HUMAN_NONE = 0xFF
_CONTROL_MESSAGES = ("continue", "start game", "start game.")


def last_human_text(messages) -> Optional[str]:
    """Content of the last message when it is a HumanMessage (langchain object, or the dict form {"type": "human"} /
    {"role": "user"}), else None — the test the reference's router applies (agent/tools/utils.py:330-331)."""
    if not messages:
        return None
    m = messages[-1]
    if isinstance(m, dict):
        kind = str(m.get("type") or m.get("role") or "").lower()
        return str(m.get("content", "")) if kind in ("human", "user") else None
    if type(m).__name__ == "HumanMessage":
        return str(getattr(m, "content", ""))
    return None


def is_action_message(text: Optional[str]) -> bool:
    """The reference logs a human message as a game action unless it is chat or a generic control message
    (agent/tools/utils.py:334-337)."""
    if text is None:
        return False
    c = str(text).lower().strip()
    return "in game chat:" not in c and "to bot" not in c and c not in _CONTROL_MESSAGES


def parse_choice(text: str, names: List[str]) -> Optional[int]:
    """The number a person's action message chooses: the UI's `Player 1 voted "<option>" in voting <id>`
    (reference src/app/page.tsx:302-305; the option is a player label, a player name or an option number), or free
    text naming `Player N` / `statement N` / `option N` / a bare number.  None when nothing can be read."""
    t = str(text)
    m = _re.search(r'voted\s+"([^"]*)"', t)
    if m:
        t = m.group(1)
    else:                                       # "Player 1: protect Player 4" — the leading label is the speaker, not the choice
        m2 = _re.match(r"^\s*Player\s+\d+\b\s*[:,\-]?\s*(\S.*)$", t, _re.IGNORECASE | _re.DOTALL)
        if m2:
            t = m2.group(1)
    for i, nm in enumerate(names):
        if nm and not _re.fullmatch(r"Player \d+", nm) and nm.lower() in t.lower():
            return i + 1
    m = _re.search(r"(?:Player|statement|option)\s*#?\s*(\d+)", t, _re.IGNORECASE) or _re.search(r"\b(\d+)\b", t)
    return int(m.group(1)) if m else None


def format_note(note_type: str, content: str) -> str:
    return "%s %s: %s" % (NOTE_EMOJI.get(note_type, "📝"), note_type, content)


def _bits(mask: int, P: int) -> List[int]:
    return [p for p in range(P) if (mask >> p) & 1]


class Record:
    """Decoded view of one canonical record (SPEC.md section 5)."""

    def __init__(self, cg: CompiledGame, raw: np.ndarray):
        self.cg = cg
        r = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
        assert r.size == cg.record_size, (r.size, cg.record_size)
        self.raw = r
        self.phase, self.prev = int(r[0]), int(r[1])
        self.step = int(r[2]) | (int(r[3]) << 8)
        P = cg.n_players
        u32 = lambda off: int.from_bytes(r[off:off + 4].tobytes(), "little")
        if cg.family == T.FAMILY_WEREWOLF:
            self.winner, self.kill, self.protect, self.revote = int(r[4]), int(r[5]), int(r[6]), int(r[7])
            names = ("alive", "can_vote", "eligible", "submitted", "revealed", "investigated", "wolf", "secret", "role_lo", "role_hi")
            for i, n in enumerate(names):
                setattr(self, n, u32(8 + 4 * i))
            self.target = [int(r[48 + p]) for p in range(P)]
            self.role = [((self.role_lo >> p) & 1) | (((self.role_hi >> p) & 1) << 1) for p in range(P)]
        else:
            self.speaker, self.lie_index, self.winner = int(r[4]), int(r[5]), int(r[6])
            self.score = [int(r[8 + 4 * p]) for p in range(P)]
            self.rounds = [int(r[9 + 4 * p]) for p in range(P)]
            self.vote = [int(r[10 + 4 * p]) for p in range(P)]
            self.flags = [int(r[11 + 4 * p]) for p in range(P)]

    # ---- mask fields (SPEC.md section 2)
    def field(self, f: int) -> int:
        P = self.cg.n_players
        ALL = (1 << P) - 1
        if f == T.F_ALL:
            return ALL
        k = f - T.cmp_field_id(self.cg.family, 0)             # comparison fields (numeric conditions)
        if 0 <= k < len(self.cg.table.cmps) and f < T.F_ALL:
            vf, op, const = self.cg.table.cmps[k]
            vals = self.target if self.cg.family == T.FAMILY_WEREWOLF else (self.score, self.rounds, self.vote)[vf]
            return sum(1 << p for p in range(P) if T.cmp_holds(op, vals[p], const))
        if self.cg.family == T.FAMILY_WEREWOLF:
            base = [self.alive, self.can_vote, self.eligible, self.submitted, self.revealed, self.investigated, self.wolf, self.secret]
            if f < 8:
                return base[f]
            if 8 <= f < 12:
                return sum(1 << p for p in range(P) if self.role[p] == f - 8) if self.secret else 0
            if f == T.W_ROLES_ASSIGNED:
                return ALL if self.secret else 0
            return 0
        if f < 5:
            return sum(1 << p for p in range(P) if (self.flags[p] >> f) & 1)
        return 0

    def eval_pred(self, pred) -> int:
        """Lane mask of a predicate record (4-tuple) — or of a table predicate INDEX, following continued records."""
        if isinstance(pred, int):
            out, i = 0, pred
            while True:
                rec = self.cg.table.preds[i]
                out |= self.eval_pred(tuple(rec))
                if not (rec[0] & T.PRED_CONTINUED):
                    return out
                i += 1
        ALL = (1 << self.cg.n_players) - 1
        out = 0
        for pos, neg in ((pred[0] & ~T.PRED_CONTINUED, pred[1]), (pred[2], pred[3])):
            m = ALL
            for f in range(16):
                if (pos >> f) & 1:
                    m &= self.field(f)
                if (neg >> f) & 1:
                    m &= ~self.field(f) & ALL
            out |= m
        return out


class SessionCodec:
    """dict <-> record conversion for one compiled game."""

    def __init__(self, cg: CompiledGame):
        self.cg = cg
        self.P = cg.n_players
        # canonical field name -> the DSL's name for it (rules `fields:` aliases); identity when not aliased
        self.dsl_name = {canon: dsl for dsl, canon in (cg.field_alias or {}).items()}

    def _k(self, canon: str) -> str:
        return self.dsl_name.get(canon, canon)

    # ------------------------------------------------------------------ init (utils.py:584-653 + AgentState defaults)
    def initial_state(self, room_players: Optional[List[dict]] = None, game_name: Optional[str] = None) -> Dict[str, Any]:
        players = room_players or [{"name": "Player %d" % (i + 1), "gamePlayerId": str(i + 1)} for i in range(self.P)]
        assert len(players) == self.P
        ps = {}
        for i, pl in enumerate(players):
            entry = {k: (dict(v) if isinstance(v, dict) else v) for k, v in self.cg.template.items()}
            entry["name"] = pl.get("name", "Player %d" % (i + 1))
            ps[str(i + 1)] = entry
        return {
            "current_phase_id": 0, "current_phase_name": "", "player_states": ps, "gameName": game_name or self.cg.name,
            "dsl": self.cg.dsl, "roomSession": {"players": players}, "playerActions": {}, "phase_history": [],
            "game_notes": [],
        }

    def initial_record(self) -> np.ndarray:
        """Canonical record of a freshly created session (what ge_batch_create writes for every slot)."""
        return self.record_from_state(self.initial_state())

    # ------------------------------------------------------------------ record -> player_states
    def player_states_from_record(self, rec: Record, names: List[str], prev_ps: Optional[Dict[str, Any]] = None,
                                  written: Optional[Dict[int, set]] = None) -> Dict[str, Dict[str, Any]]:
        """Per-player dicts of a record.  Canonical fields the DSL's template does not declare exist in the
        reference's dict only once a referee tool call has written them (backend_tools.py:204-225 creates the key);
        with `prev_ps` / `written` (step_update) such a field is emitted only if the previous state had it or
        this step wrote it, otherwise every canonical field is emitted."""
        cg, P = self.cg, self.P
        out: Dict[str, Dict[str, Any]] = {}
        tpl_keys = set(cg.template.keys())
        if cg.family == T.FAMILY_WEREWOLF:
            assigned = rec.secret != 0
            village, wolf_team = cg.teams
            for p in range(P):
                bit = lambda m: bool((m >> p) & 1)
                role = cg.role_names[rec.role[p]] if assigned else ""
                team = (wolf_team if bit(rec.wolf) else village) if assigned else ""
                inv: Dict[str, str] = {}
                if assigned and rec.role[p] == 3:
                    for t in _bits(rec.investigated, P):
                        inv[str(t + 1)] = wolf_team if (rec.wolf >> t) & 1 else village
                entry = {
                    "name": names[p], "role": role, "team": team, "is_alive": bit(rec.alive), "role_revealed": bit(rec.revealed),
                    "can_vote": bit(rec.can_vote), "has_secret_role": bit(rec.secret), "night_action_eligible": bit(rec.eligible),
                    "night_action_submitted": bit(rec.submitted), "selected_target_id": rec.target[p],
                    "investigated_alignments": inv,
                }
                if "team_is_wolf" in self.dsl_name:
                    entry["team_is_wolf"] = assigned and bit(rec.wolf)
                if cg.session_fields:       # session-level re-vote state, carried by every player (rules session_fields)
                    entry[cg.session_fields["revote_count"]] = rec.revote & 0x7F
                    entry[cg.session_fields["tie_pending"]] = bool(rec.revote & 0x80)
                if self.dsl_name:
                    entry = {self._k(k): v for k, v in entry.items()}
                if prev_ps is not None:
                    old = prev_ps.get(str(p + 1), {})
                    wr = {self._k(k) for k in (written or {}).get(p, ())}
                    entry = {k: v for k, v in entry.items() if k in tpl_keys or k in old or k in wr}
                out[str(p + 1)] = entry
        else:
            for p in range(P):
                fl = rec.flags[p]
                is_sp = bool(fl & 1)
                stm = {str(i + 1): s for i, s in enumerate(TTL_STATEMENTS)} if (fl & 2) else {}
                out[str(p + 1)] = {
                    "name": names[p], "is_speaker": is_sp, "statements": stm, "statements_submitted": bool(fl & 2),
                    "lie_index": rec.lie_index if is_sp else 0, "lie_revealed": bool(fl & 4), "can_vote": bool(fl & 8),
                    "vote_choice": rec.vote[p], "has_voted": bool(fl & 16), "total_score": rec.score[p],
                    "rounds_as_speaker": rec.rounds[p],
                }
        return out

    # ------------------------------------------------------------------ dict -> record
    def record_from_state(self, state: Dict[str, Any], before_phase_node: bool = True) -> np.ndarray:
        """Packs the session described by an AgentState dict (SPEC.md section 5).

        `phase_history` gives step (= its length) and prev (= the entry before the current phase)."""
        cg, P = self.cg, self.P
        r = np.zeros(cg.record_size, dtype=np.uint8)
        hist = state.get("phase_history") or []
        phase = cg.index_of(state.get("current_phase_id", 0))
        prev = cg.index_of(hist[-2]["phase_id"]) if len(hist) >= 2 else 0
        step = len(hist)
        r[0], r[1], r[2], r[3] = phase, prev, step & 0xFF, step >> 8
        ps = state.get("player_states") or {}
        if self.dsl_name:       # read through the aliases: present every entry under its canonical field names
            canon_of = dict(cg.field_alias)
            ps = {pid: {canon_of.get(k, k): v for k, v in e.items()} for pid, e in ps.items()}
        get = lambda p: ps.get(str(p + 1), {})
        put32 = lambda off, v: r.__setitem__(slice(off, off + 4), np.frombuffer(int(v).to_bytes(4, "little"), dtype=np.uint8))
        if cg.family == T.FAMILY_WEREWOLF:
            village, wolf_team = cg.teams
            mask = lambda key: sum(1 << p for p in range(P) if get(p).get(key))
            alive, can_vote, elig, sub = mask("is_alive"), mask("can_vote"), mask("night_action_eligible"), mask("night_action_submitted")
            rev, secret = mask("role_revealed"), mask("has_secret_role")
            wolf = sum(1 << p for p in range(P) if get(p).get("team") == wolf_team)
            roles = [cg.role_names.index(get(p)["role"]) if get(p).get("role") in cg.role_names else 0 for p in range(P)]
            lo = sum(1 << p for p in range(P) if roles[p] & 1)
            hi = sum(1 << p for p in range(P) if roles[p] & 2)
            inv = 0
            for p in range(P):
                for k in (get(p).get("investigated_alignments") or {}):
                    inv |= 1 << (int(k) - 1)
            tgt = [int(get(p).get("selected_target_id") or 0) for p in range(P)]
            # kill/protect are scratch that lives between the wolves' vote and the night resolution (SPEC section 4)
            exit_op = cg.table.phases[phase].exit_op
            kill = protect = 0
            if exit_op in (T.EX_PROTECT, T.EX_INVESTIGATE_RESOLVE):
                votes = [tgt[p] for p in range(P) if roles[p] == 1 and (sub >> p) & 1 and tgt[p]]
                if votes:
                    best = max(set(votes), key=lambda c: (votes.count(c), -c))
                    kill = best
            if exit_op == T.EX_INVESTIGATE_RESOLVE:
                docs = [p for p in range(P) if roles[p] == 2 and (sub >> p) & 1]
                protect = tgt[docs[0]] if docs else 0
            winner = 0
            if cg.table.phases[phase].kind == T.KIND_TERMINAL:
                winner = 1 if (wolf & alive) == 0 else 2
            revote = 0
            if cg.session_fields:           # every player carries the same value; the lowest id is read
                g0 = get(0)
                revote = (int(g0.get(cg.session_fields["revote_count"]) or 0) & 0x7F) | (0x80 if g0.get(cg.session_fields["tie_pending"]) else 0)
            r[4], r[5], r[6], r[7] = winner, kill, protect, revote
            for i, v in enumerate((alive, can_vote, elig, sub, rev, inv, wolf, secret, lo, hi)):
                put32(8 + 4 * i, v)
            for p in range(P):
                r[48 + p] = tgt[p]
        else:
            speaker = next((p + 1 for p in range(P) if get(p).get("is_speaker")), 0)
            lie = int(get(speaker - 1).get("lie_index") or 0) if speaker else 0
            winner = 0
            if cg.table.phases[phase].kind == T.KIND_TERMINAL:
                scores = [int(get(p).get("total_score") or 0) for p in range(P)]
                winner = scores.index(max(scores)) + 1
            # between Round Start and the speaker's lie selection the reference dict keeps the previous speaker out
            r[4], r[5], r[6] = speaker, lie, winner
            for p in range(P):
                g = get(p)
                fl = (1 if g.get("is_speaker") else 0) | (2 if g.get("statements_submitted") else 0) | (4 if g.get("lie_revealed") else 0) \
                    | (8 if g.get("can_vote") else 0) | (16 if g.get("has_voted") else 0)
                r[8 + 4 * p: 12 + 4 * p] = [int(g.get("total_score") or 0), int(g.get("rounds_as_speaker") or 0), int(g.get("vote_choice") or 0), fl]
        return r

    # ------------------------------------------------------------------ one step's worth of reference-format updates
    def step_update(self, state: Dict[str, Any], before: np.ndarray, after: np.ndarray,
                    now_ms: Optional[int] = None, now_iso: Optional[str] = None, human_mask: int = 0) -> Dict[str, Any]:
        """The union of the update dicts of BotBehaviorNode + PhaseNode + RefereeNode for the step that took the
        session from record `before` to record `after` (both canonical).  `human_mask`: seats played by people — their
        actions are logged by the router, not here, and a step that waited for one of them only grows the history."""
        cg, P = self.cg, self.P
        b, a = Record(cg, before), Record(cg, after)
        now_ms = int(_time.time() * 1000) if now_ms is None else now_ms
        now_iso = _dt.datetime.now().isoformat() if now_iso is None else now_iso
        ps_old = state.get("player_states") or {}
        names = [ps_old.get(str(p + 1), {}).get("name", "Player %d" % (p + 1)) for p in range(P)]
        actions = {pid: {"name": v.get("name"), "actions": dict(v.get("actions", {}))} for pid, v in (state.get("playerActions") or {}).items()}
        history = list(state.get("phase_history") or [])
        notes = list(state.get("game_notes") or [])
        if a.step == b.step:                       # terminal: nothing happened
            return {"player_states": ps_old, "playerActions": actions, "current_phase_id": cg.phase_ids[a.phase],
                    "current_phase_name": cg.phase_names[a.phase], "phase_history": history, "game_notes": notes}
        X, Y = b.phase, a.phase
        phX = cg.table.phases[X]
        if self.stayed(before, after, human_mask):         # PhaseNode appends to the history even when it stays (:1206-1215)
            history.append({"phase_id": cg.phase_ids[Y], "phase_name": cg.phase_names[Y], "timestamp": now_iso})
            return {"player_states": ps_old, "playerActions": actions, "current_phase_id": cg.phase_ids[Y],
                    "current_phase_name": cg.phase_names[Y], "phase_history": history, "game_notes": notes}
        # ---- BotBehaviorNode part: the actions the bots took in phase X
        if b.step > 0 and phX.kind == T.KIND_ACTION:
            actors = b.eval_pred(int(phX.actor_pred)) & ~human_mask
            for p in _bits(actors, P):
                if cg.family == T.FAMILY_WEREWOLF:
                    choice = a.target[p]
                elif phX.exit_op == T.EX_T_LIE:
                    choice = a.lie_index
                elif phX.exit_op == T.EX_T_VOTES:
                    choice = a.vote[p]
                else:
                    choice = 1
                text = self.action_text(X, choice)
                pid = str(p + 1)
                slot = actions.setdefault(pid, {"name": names[p], "actions": {}})
                slot["name"] = names[p]
                ids = [int(v["id"]) for v in slot["actions"].values() if isinstance(v, dict) and str(v.get("id", "")).isdigit()]
                aid = str(max(ids, default=0) + 1)
                slot["actions"][aid] = {"action": text, "timestamp": now_ms, "phase": cg.phase_names[X], "id": aid}
        # ---- PhaseNode part
        history.append({"phase_id": cg.phase_ids[Y], "phase_name": cg.phase_names[Y], "timestamp": now_iso})
        # ---- RefereeNode part
        new_ps = self.player_states_from_record(a, names, prev_ps=ps_old, written=self.written_fields(b, a))
        if b.step > 0:
            notes += [format_note(t, c) for t, c in self.notes_for(b, a, names)]
        # the reference's phase-0 first visit returns no current_phase_name (game_agent_v2.py:1043-1052)
        name = cg.phase_names[Y] if b.step > 0 else state.get("current_phase_name", "")
        return {"player_states": new_ps, "playerActions": actions, "current_phase_id": cg.phase_ids[Y],
                "current_phase_name": name, "phase_history": history, "game_notes": notes}

    # ------------------------------------------------------------------ people at the table (SPEC.md D3h)
    def log_human_action(self, state: Dict[str, Any], text: str, player_id: str = "1", now_ms: Optional[int] = None) -> Dict[str, Any]:
        """What the reference's router does with a person's action message before the bots run
        (process_human_action_if_needed, agent/tools/utils.py:310-358, called game_agent_v2.py:324-332): the raw text,
        truncated to 200 characters, is appended to playerActions[player_id] — under the name of PHASE 0, because the
        router passes camelCase keys the state does not have (`currentPhaseId`, game_agent_v2.py:327-328).  Returns the
        new playerActions; for graphs that replace InitialRouterNode as well."""
        actions = {pid: {"name": v.get("name"), "actions": dict(v.get("actions", {}))} for pid, v in (state.get("playerActions") or {}).items()}
        if not is_action_message(text):
            return actions
        name = "Player %s" % player_id          # the router also looks under `playerStates`, which the state does not have
        for p in (state.get("roomSession") or {}).get("players", []):
            if str(p.get("gamePlayerId", "")) == str(player_id):
                name = p.get("name", name)
        slot = actions.setdefault(str(player_id), {"name": name, "actions": {}})
        ids = [int(v["id"]) for v in slot["actions"].values() if isinstance(v, dict) and str(v.get("id", "")).isdigit()]
        aid = str(max(ids, default=0) + 1)
        slot["name"] = name
        slot["actions"][aid] = {"action": str(text)[:200], "timestamp": int(_time.time() * 1000) if now_ms is None else now_ms,
                                "phase": self.cg.phase_names[0], "id": aid}
        return actions

    def human_inputs(self, state: Dict[str, Any], human_seats=(1,)) -> Tuple[int, np.ndarray]:
        """(human_mask, inputs row) of this graph run: every human seat's input is what the last human message
        chooses, when that message is an action (the reference has one person, player 1, whose message arrives in
        the run that logs it); 0xFF otherwise."""
        P = self.P
        mask = 0
        row = np.full(((P + 7) // 8) * 8, HUMAN_NONE, dtype=np.uint8)
        ps = state.get("player_states") or {}
        names = [ps.get(str(p + 1), {}).get("name", "") for p in range(P)]
        text = last_human_text(state.get("messages"))
        phase = self.cg.table.phases[self.cg.index_of(state.get("current_phase_id", 0))]
        for seat in human_seats:
            if not (1 <= int(seat) <= P):
                continue
            mask |= 1 << (int(seat) - 1)
            if is_action_message(text):
                # a MARK action ("submit your statements") is answered by any action message; the others by the number it names
                c = 1 if phase.action_op == T.ACT_MARK else parse_choice(text, names)
                if c is not None and 0 <= c < HUMAN_NONE:
                    row[int(seat) - 1] = c
        return mask, row

    def stayed(self, before: np.ndarray, after: np.ndarray, human_mask: int) -> bool:
        """True when the step b -> a was a wait for a person (SPEC D3h): same phase, history one longer, nothing else."""
        if not human_mask:
            return False
        b, a = Record(self.cg, before), Record(self.cg, after)
        phX = self.cg.table.phases[b.phase]
        if b.step == 0 or a.step != b.step + 1 or a.phase != b.phase or a.prev != b.phase or phX.kind != T.KIND_ACTION:
            return False
        if not (b.eval_pred(int(phX.actor_pred)) & human_mask):
            return False
        return bool(np.array_equal(before[4:], after[4:]))

    # ------------------------------------------------------------------ the step as the reference's tool calls
    def tool_calls_for(self, state: Dict[str, Any], before: np.ndarray, after: np.ndarray, human_mask: int = 0) -> Dict[str, List[dict]]:
        """The step b -> a as the tool-call lists the reference's three hot-path nodes apply, in the order they apply
        them (agent/tools/backend_tools.py:10-157; applied in list order, game_agent_v2.py:589-605, 1124-1143, 762-786):

            {"BotBehaviorNode": [update_player_actions...], "PhaseNode": [set_next_phase], "RefereeNode": [update_player_state..., add_game_note...]}

        Feeding these lists to the reference's own nodes through a pass-through chat model reproduces, through the
        reference's `_execute_*` code, the state `step_update` builds directly (tests/test_tool_calls.py)."""
        cg, P = self.cg, self.P
        b, a = Record(cg, before), Record(cg, after)
        out: Dict[str, List[dict]] = {"BotBehaviorNode": [], "PhaseNode": [], "RefereeNode": []}
        if a.step == b.step:                                        # terminal session: PhaseNode would not transition
            out["PhaseNode"].append({"name": "set_next_phase", "id": "ph", "args": {
                "transition": False, "next_phase_id": cg.phase_ids[a.phase], "transition_reason": "terminal"}})
            return out
        if b.step == 0:                                             # phase-0 first visit: no LLM is consulted at all
            return out
        X, Y = b.phase, a.phase
        phX = cg.table.phases[X]
        wait = self.stayed(before, after, human_mask)
        ps_old = state.get("player_states") or {}
        names = [ps_old.get(str(p + 1), {}).get("name", "Player %d" % (p + 1)) for p in range(P)]
        actors = _bits(b.eval_pred(int(phX.actor_pred)), P) if phX.kind == T.KIND_ACTION else []
        if not wait:
            for p in actors:
                if (human_mask >> p) & 1:
                    continue                                        # the router logged the person's own message
                out["BotBehaviorNode"].append({"name": "update_player_actions", "id": "bot-%d-%d" % (b.step, p + 1), "args": {
                    "player_id": str(p + 1), "actions": self.action_text(X, self._choice_of(phX, a, p)), "phase": cg.phase_names[X]}})
        out["PhaseNode"].append({"name": "set_next_phase", "id": "ph", "args": {
            "transition": not wait, "next_phase_id": cg.phase_ids[Y],
            "transition_reason": "waiting for the human player's action" if wait else "phase complete"}})
        if wait:
            return out
        # RefereeNode: every field of every player that differs, plus the fields the effects write even when the value
        # stays the same (they create keys the template does not declare), then the notes
        new_ps = self.player_states_from_record(a, names, prev_ps=ps_old, written=self.written_fields(b, a))
        n = 0
        for p in range(P):
            pid = str(p + 1)
            old, new = ps_old.get(pid, {}), new_ps[pid]
            for k, v in new.items():
                if k == "name":
                    continue
                if k not in old or old[k] != v:
                    out["RefereeNode"].append({"name": "update_player_state", "id": "r%d" % n, "args": {"player_id": pid, "state_name": k, "state_value": v}})
                    n += 1
        for t_, c_ in self.notes_for(b, a, names):
            out["RefereeNode"].append({"name": "add_game_note", "id": "n%d" % n, "args": {"note_type": t_, "content": c_}})
            n += 1
        return out

    def _choice_of(self, phX, a: Record, p: int) -> int:
        if self.cg.family == T.FAMILY_WEREWOLF:
            return a.target[p]
        if phX.exit_op == T.EX_T_LIE:
            return a.lie_index
        if phX.exit_op == T.EX_T_VOTES:
            return a.vote[p]
        return 1

    def written_fields(self, b: Record, a: Record) -> Dict[int, set]:
        """{player index: canonical fields the referee writes in the step b -> a} (SPEC.md section 4 effects)."""
        cg, P = self.cg, self.P
        w: Dict[int, set] = {p: set() for p in range(P)}
        if cg.family != T.FAMILY_WEREWOLF or a.step == b.step or b.step == 0:
            return w
        phX = cg.table.phases[b.phase]
        ex, en = phX.exit_op, cg.table.phases[a.phase].entry_op
        actors = _bits(b.eval_pred(int(phX.actor_pred)), P) if phX.kind == T.KIND_ACTION else []
        died = _bits(b.alive & ~a.alive, P)
        if ex in (T.EX_VOTE_KILL, T.EX_PROTECT, T.EX_INVESTIGATE_RESOLVE):
            for p in actors:
                w[p] |= {"selected_target_id", "night_action_submitted"}
        if ex == T.EX_INVESTIGATE_RESOLVE:
            for p in actors:
                if a.target[p]:
                    w[p].add("investigated_alignments")
        if ex == T.EX_DAY_VOTE:
            for p in actors:
                w[p].add("selected_target_id")
            for p in died:
                w[p].add("role_revealed")
        for p in died:
            w[p] |= {"is_alive", "can_vote", "night_action_eligible"}
        if en == T.EN_ASSIGN_ROLES:
            for p in range(P):
                w[p] |= {"role", "team", "has_secret_role", "night_action_eligible", "team_is_wolf"}
        if en == T.EN_NIGHT_RESET:
            for p in range(P):
                w[p] |= {"night_action_submitted", "selected_target_id"}
        return w

    def action_text(self, phase_index: int, choice: int) -> str:
        tpl = self.cg.action_text.get(phase_index, "acted")
        if "{t}" in tpl and choice == 0:
            return NO_TARGET_TEXT
        return tpl.format(t=choice, s1=TTL_STATEMENTS[0], s2=TTL_STATEMENTS[1], s3=TTL_STATEMENTS[2])

    def notes_for(self, b: Record, a: Record, names: List[str]) -> List[Tuple[str, str]]:
        """Referee notes of one step (formats fixed here and mirrored by the Oracle A stub)."""
        cg, P = self.cg, self.P
        out: List[Tuple[str, str]] = []
        ex, en = cg.table.phases[b.phase].exit_op, cg.table.phases[a.phase].entry_op
        if cg.family == T.FAMILY_WEREWOLF:
            died = _bits(b.alive & ~a.alive, P)
            role_of = lambda p: cg.role_names[a.role[p]]
            if ex == T.EX_INVESTIGATE_RESOLVE:
                if died:
                    out.append(("CRITICAL", "Player %d (%s) was eliminated during the night - marked is_alive=false" % (died[0] + 1, role_of(died[0]))))
                elif b.kill:
                    out.append(("DECISION", "Werewolves targeted Player %d, Doctor protected Player %d - no elimination" % (b.kill, b.protect)))
                else:
                    out.append(("DECISION", "No werewolf target - no elimination"))
            elif ex == T.EX_DAY_VOTE:
                if died:
                    out.append(("CRITICAL", "Player %d (%s) was eliminated by day vote - marked is_alive=false" % (died[0] + 1, role_of(died[0]))))
                elif a.revote & 0x80:
                    out.append(("DECISION", "Day vote tied - re-vote %d of %d" % (a.revote & 0x7F, cg.table.max_revotes)))
                else:
                    out.append(("DECISION", "Day vote produced no elimination"))
            if en == T.EN_ASSIGN_ROLES:
                out.append(("NEXT_PHASE", "Roles assigned: " + ", ".join("Player %d=%s" % (p + 1, role_of(p)) for p in range(P))))
            if a.winner and not b.winner:
                out.append(("GAME_STATUS", "Game over - %s win" % (cg.teams[0] if a.winner == 1 else cg.teams[1])))
        else:
            if en == T.EN_T_ROUND_START:
                out.append(("DECISION", "Selected Player %d as next speaker" % a.speaker))
            elif en == T.EN_T_SCORE:
                out.append(("SCORE_UPDATE", "Round totals: " + ", ".join("Player %d: %d points" % (p + 1, a.score[p]) for p in range(P))))
            elif en == T.EN_T_FINAL:
                out.append(("GAME_STATUS", "Game over - Player %d wins with %d points" % (a.winner, a.score[a.winner - 1])))
        return out
