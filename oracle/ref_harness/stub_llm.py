"""Rule-following stand-in for the reference's chat model (gpt-4.1-mini, game_agent_v2.py:523, 684, 1075).

The reference's nodes call `init_chat_model(...).bind_tools(tools).ainvoke([system_message], config)` and apply
the returned `tool_calls`.  This stub answers with exactly the tool calls SPEC.md prescribes.  It is a THIRD,
independent statement of the rules: it works on the reference's dict state, evaluates the DSL's condition
strings with Python (not with the DNF compiler of game_engine_b200), and draws from the same Philox stream.

Which personality answers is decided by the tools the node bound:
    [update_player_actions]              -> bots       (BotBehaviorNode)
    [set_next_phase]                     -> phase      (PhaseNode)
    [update_player_state, add_game_note] -> referee    (RefereeNode)

The node's prompt is not parsed; the harness hands the stub the same state dict the node received
(`stub.state = state`), which is what the prompt is rendered from.  Test infrastructure only.
"""
from __future__ import annotations

import re
from typing import Any, Dict, List

from .shims import AIMessage

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
NO_TARGET_TEXT = "had no legal target"
TTL_STATEMENTS = ("I once met a celebrity.", "I can speak four languages.", "I've never broken a bone.")


def philox4x32_10(key, ctr):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def draw(seed: int, sid: int, step: int, purpose: int, p0: int) -> int:
    out = philox4x32_10((seed & MASK, (seed >> 32) & MASK), (sid & MASK, (sid >> 32) & MASK, step, (purpose << 16) | (p0 >> 2)))
    return out[p0 & 3]


def mulhi(r: int, n: int) -> int:
    return (r * n) >> 32


class _P:
    """`player.<field>` access for eval of DSL conditions."""

    def __init__(self, d):
        self._d = d

    def __getattr__(self, k):
        return self._d.get(k)


def holds(cond: str, pstate: dict) -> bool:
    src = re.sub(r"\btrue\b", "True", re.sub(r"\bfalse\b", "False", cond))
    return bool(eval(src, {"__builtins__": {}}, {"player": _P(pstate)}))


def plurality(choices: List[int]) -> int:
    votes = [c for c in choices if c]
    if not votes:
        return 0
    return max(set(votes), key=lambda c: (votes.count(c), -c))


def human_number(text: str):
    """The number a person's message chooses (the UI's `Player 1 voted "<option>" in voting <id>`, reference
    src/app/page.tsx:302-305, or `... Player N` / `... statement N`), or None.  Written independently of
    game_engine_b200.adapter.parse_choice: the fixtures only use messages both read the same way."""
    m = re.search(r'voted\s+"([^"]*)"', text)
    body = m.group(1) if m else re.sub(r"^\s*Player\s+\d+\s*[:,-]?\s*", "", text, count=1)
    m = re.search(r"(\d+)", body)
    return int(m.group(1)) if m else None


class StubChatModel:
    def __init__(self, rules: dict, seed: int, sid: int, human_seats=()):
        self.rules, self.seed, self.sid = rules, seed, sid
        self.human_seats = tuple(int(x) for x in human_seats)       # SPEC D3h; the reference's room: (1,)
        self.state: Dict[str, Any] = {}
        self.calls = 0
        # rules `fields:` maps a DSL field name to the canonical one; the referee writes / reads the DSL's names
        self.dsl_name = {canon: dsl for dsl, canon in (rules.get("fields") or {}).items()}

    def F(self, canon: str) -> str:
        return self.dsl_name.get(canon, canon)

    def _sf(self, item: str) -> str:
        """DSL per-player field that carries a session-level item (rules `session_fields:`; extended re-vote game)."""
        return (self.rules.get("session_fields") or {}).get(item, item)

    def bind_tools(self, tools, **kw):
        return _Bound(self, frozenset(t.name for t in tools))

    # ------------------------------------------------------------------ helpers on the dict state
    def _phase(self, pid):
        ph = self.state["dsl"]["phases"]
        return ph.get(pid) if pid in ph else ph.get(str(pid))

    def _prules(self, pid) -> dict:
        pr = self.rules.get("phases") or {}
        return pr.get(pid) or pr.get(str(pid)) or {}

    def _ids(self) -> List[int]:
        return sorted(int(k) for k in self.state["player_states"].keys())

    def _ps(self, i: int) -> dict:
        return self.state["player_states"][str(i)]

    def _actors(self, phase: dict) -> List[int]:
        cc = phase.get("completion_criteria") or {}
        if cc.get("type") != "player_action":
            return []
        cond = cc["target_players"]["condition"]
        return [i for i in self._ids() if holds(cond, self._ps(i))]

    # ------------------------------------------------------------------ people at the table (SPEC D3h)
    def _human_text(self):
        msgs = self.state.get("messages") or []
        if not msgs or type(msgs[-1]).__name__ != "HumanMessage":
            return None
        c = str(msgs[-1].content)
        low = c.lower().strip()
        if "in game chat:" in low or "to bot" in low or low in ("continue", "start game", "start game."):
            return None
        return c

    def _legal(self, act: dict, i: int) -> List[int]:
        return [q for q in self._ids() if holds(act["legal"], self._ps(q)) and not (act.get("exclude_self") and q == i)]

    def _human_answer(self, i: int, act: dict):
        """What human seat i chooses in this graph run, or None when the phase has to wait for it."""
        if act["op"] == "PICK_PLAYER" and not self._legal(act, i):
            return 0
        text = self._human_text()
        if text is None:
            return None
        if act["op"] == "MARK":
            return 1
        c = human_number(text)
        if c is None:
            return None
        if act["op"] == "PICK_PLAYER":
            return c if c in self._legal(act, i) else None
        return c if 1 <= c <= int(act["options"]) else None

    def _waiting(self, X) -> bool:
        """Phase X (a player_action phase) still lacks the answer of a human seat among its actors."""
        act = self._prules(X).get("action")
        if not act or not self.human_seats:
            return False
        return any(self._human_answer(i, act) is None for i in self._actors(self._phase(X)) if i in self.human_seats)

    # ------------------------------------------------------------------ BotBehaviorNode
    def bots(self) -> List[dict]:
        st = self.state
        X = st.get("current_phase_id", 0)
        phase = self._phase(X)
        step = len(st.get("phase_history") or [])
        act = self._prules(X).get("action")
        if step == 0 or not act:
            return []
        if self._waiting(X):                      # bots act on the run that completes the phase
            return []
        calls = []
        for i in self._actors(phase):
            if i in self.human_seats:             # "Player ID 1 (human) NEVER generates actions" (bot prompt :3)
                continue
            r = draw(self.seed, self.sid, step, 0, i - 1)
            if act["op"] == "PICK_PLAYER":
                legal = [q for q in self._ids() if holds(act["legal"], self._ps(q)) and not (act.get("exclude_self") and q == i)]
                choice = legal[mulhi(r, len(legal))] if legal else 0
                text = act["text"].format(t=choice) if choice else NO_TARGET_TEXT
            elif act["op"] == "PICK_OPTION":
                choice = 1 + mulhi(r, int(act["options"]))
                text = act["text"].format(t=choice)
            else:
                text = act["text"].format(s1=TTL_STATEMENTS[0], s2=TTL_STATEMENTS[1], s3=TTL_STATEMENTS[2])
            calls.append({"name": "update_player_actions", "args": {"player_id": str(i), "actions": text, "phase": phase["name"]},
                          "id": "bot-%d-%d" % (step, i)})
        return calls

    # ------------------------------------------------------------------ PhaseNode
    def phase(self) -> List[dict]:
        st = self.state
        X = st.get("current_phase_id", 0)
        nxt = self._phase(X).get("next_phase")
        if nxt is None:
            return [{"name": "set_next_phase", "args": {"transition": False, "next_phase_id": X, "transition_reason": "terminal"}, "id": "ph"}]
        if (st.get("phase_history") or []) and self._waiting(X):
            return [{"name": "set_next_phase", "args": {"transition": False, "next_phase_id": X, "transition_reason": "waiting for the human player's action"}, "id": "ph"}]
        if "id" in nxt and not isinstance(nxt["id"], dict):
            target, why = nxt["id"], "phase complete"
        else:
            hist = st.get("phase_history") or []
            prev = hist[-2]["phase_id"] if len(hist) >= 2 else None
            ann = self._prules(X)["branches"]
            target, why = None, ""
            for a in ann:
                ids = self._ids()
                cnt = lambda cond: sum(1 for i in ids if holds(cond, self._ps(i)))
                op = a["op"]
                ok = (op == "ALWAYS" or (op == "COUNT_EQ0" and cnt(a["a"]) == 0) or (op == "COUNT_GE" and cnt(a["a"]) >= cnt(a["b"]))
                      or (op == "PREV_IN" and prev in a["phases"])
                      or (op == "TIE_PENDING" and bool(self._ps(ids[0]).get(self._sf("tie_pending"))))
                      or (op == "ALL_VAL_GE" and all((self._ps(i).get(a["field"]) or 0) >= (self.rules["rounds"] if a["value"] == "rounds" else int(a["value"])) for i in ids)))
                if ok:
                    target, why = nxt[a["key"]]["id"], a["key"]
                    break
            if target is None:
                target, why = nxt[ann[-1]["key"]]["id"], "fallback: last branch"
        return [{"name": "set_next_phase", "args": {"transition": True, "next_phase_id": target, "transition_reason": why}, "id": "ph"}]

    # ------------------------------------------------------------------ RefereeNode
    def _latest_choice(self, i: int, phase_name: str) -> int:
        acts = ((self.state.get("playerActions") or {}).get(str(i)) or {}).get("actions") or {}
        mine = [a for a in acts.values() if a.get("phase") == phase_name]
        if not mine:
            return 0
        text = max(mine, key=lambda a: int(a["id"]))["action"]        # latest action of this player in that phase
        if text == NO_TARGET_TEXT:
            return 0
        m = re.search(r"(?:Player|statement) (\d+)", text)
        return int(m.group(1)) if m else 1

    def referee(self) -> List[dict]:
        st = self.state
        hist = st.get("phase_history") or []
        if len(hist) < 2:
            return []
        X, Y = hist[-2]["phase_id"], st.get("current_phase_id")
        phX = self._phase(X)
        if X == Y and self._waiting(X):           # PhaseNode stayed: nothing to apply
            return []
        step = len(hist) - 1                      # step count at the start of this step
        ids = self._ids()
        calls: List[dict] = []
        upd = lambda i, k, v: calls.append({"name": "update_player_state", "args": {"player_id": str(i), "state_name": self.F(k), "state_value": v}, "id": "r%d" % len(calls)})
        note = lambda t, c: calls.append({"name": "add_game_note", "args": {"note_type": t, "content": c}, "id": "n%d" % len(calls)})
        ex, en = self._prules(X).get("exit"), self._prules(Y).get("entry")
        actors = self._actors(phX)
        act_x = self._prules(X).get("action")
        choice = {i: (self._human_answer(i, act_x) or 0) if i in self.human_seats else self._latest_choice(i, phX["name"]) for i in actors}
        roles = self.rules.get("roles") or {}
        wolf_team, village = self.rules.get("wolf_team"), self.rules.get("village_team")
        dead: List[int] = []

        def die(x):
            upd(x, "is_alive", False); upd(x, "can_vote", False); upd(x, "night_action_eligible", False)
            dead.append(x)

        if ex in ("VOTE_KILL", "PROTECT", "INVESTIGATE_RESOLVE"):
            for i in actors:
                upd(i, "selected_target_id", choice[i]); upd(i, "night_action_submitted", True)
        if ex == "INVESTIGATE_RESOLVE":
            for i in actors:
                if choice[i]:
                    memo = dict(self._ps(i).get(self.F("investigated_alignments")) or {})
                    memo[str(choice[i])] = self._ps(choice[i])["team"]
                    upd(i, "investigated_alignments", memo)
            wolves = [i for i in ids if self._ps(i)["role"] == roles["werewolf"] and self._ps(i)[self.F("night_action_submitted")]]
            kill = plurality([self._ps(i)[self.F("selected_target_id")] for i in wolves])
            docs = [i for i in ids if self._ps(i)["role"] == roles["doctor"] and self._ps(i)[self.F("night_action_submitted")]]
            protect = self._ps(docs[0])[self.F("selected_target_id")] if docs else 0
            if kill and kill != protect:
                role = self._ps(kill)["role"]
                die(kill)
                note("CRITICAL", "Player %d (%s) was eliminated during the night - marked is_alive=false" % (kill, role))
            elif kill:
                note("DECISION", "Werewolves targeted Player %d, Doctor protected Player %d - no elimination" % (kill, protect))
            else:
                note("DECISION", "No werewolf target - no elimination")
        if ex == "DAY_VOTE":
            for i in actors:
                upd(i, "selected_target_id", choice[i])
            votes = [choice[i] for i in actors if choice[i]]
            x = plurality(votes)
            R = int(self.rules.get("max_revotes") or 0)
            used = int(self._ps(ids[0]).get(self._sf("revote_count")) or 0) if R else 0
            tied = x != 0 and sum(1 for c in set(votes) if votes.count(c) == votes.count(x)) > 1
            if R and tied and used < R:             # SPEC section 4, EX_DAY_VOTE with max_revotes > 0: nobody dies
                for i in ids:
                    upd(i, self._sf("revote_count"), used + 1); upd(i, self._sf("tie_pending"), True)
                note("DECISION", "Day vote tied - re-vote %d of %d" % (used + 1, R))
            else:
                if R:
                    for i in ids:
                        upd(i, self._sf("tie_pending"), False)
                if x:
                    role = self._ps(x)["role"]
                    die(x); upd(x, "role_revealed", True)
                    note("CRITICAL", "Player %d (%s) was eliminated by day vote - marked is_alive=false" % (x, role))
                else:
                    note("DECISION", "Day vote produced no elimination")
        if ex == "T_STATEMENTS":
            for i in actors:
                upd(i, "statements", {str(k + 1): s for k, s in enumerate(TTL_STATEMENTS)}); upd(i, "statements_submitted", True)
        if ex == "T_LIE" and actors:
            upd(actors[0], "lie_index", choice[actors[0]])
        if ex == "T_VOTES":
            for i in actors:
                upd(i, "vote_choice", choice[i]); upd(i, "has_voted", True)

        if en == "ASSIGN_ROLES":
            P = len(ids)
            W = max(1, P // 4) if self.rules.get("wolves") == "quarter" else int(self.rules["wolves"])
            key = {i: draw(self.seed, self.sid, step, 1, i - 1) for i in ids}
            assigned = {}
            for i in ids:
                rank = sum(1 for q in ids if (key[q], q) < (key[i], i))
                role = roles["werewolf"] if rank < W else roles["doctor"] if rank == W else roles["detective"] if rank == W + 1 else roles["villager"]
                assigned[i] = role
                special = role != roles["villager"]
                upd(i, "role", role); upd(i, "team", wolf_team if role == roles["werewolf"] else village)
                upd(i, "has_secret_role", special); upd(i, "night_action_eligible", special)
                if "team_is_wolf" in self.dsl_name:
                    upd(i, "team_is_wolf", role == roles["werewolf"])
            note("NEXT_PHASE", "Roles assigned: " + ", ".join("Player %d=%s" % (i, assigned[i]) for i in ids))
        if en == "NIGHT_RESET":
            for i in ids:
                upd(i, "night_action_submitted", False); upd(i, "selected_target_id", 0)
                if self.rules.get("max_revotes"):
                    upd(i, self._sf("revote_count"), 0); upd(i, self._sf("tie_pending"), False)
        if en == "T_ROUND_START":
            R = int(self.rules["rounds"])
            pending = [i for i in ids if (self._ps(i).get("rounds_as_speaker") or 0) < R]
            sp = pending[0] if pending else 0
            for i in ids:
                upd(i, "is_speaker", i == sp); upd(i, "can_vote", i != sp); upd(i, "has_voted", False)
                upd(i, "statements_submitted", False); upd(i, "lie_revealed", False); upd(i, "vote_choice", 0)
                upd(i, "lie_index", 0); upd(i, "statements", {})
            note("DECISION", "Selected Player %d as next speaker" % sp)
        if en == "T_REVEAL":
            for i in ids:
                if self._ps(i)["is_speaker"]:
                    upd(i, "lie_revealed", True)
        if en == "T_SCORE":
            sp = next((i for i in ids if self._ps(i)["is_speaker"]), 0)
            lie = self._ps(sp)["lie_index"] if sp else 0
            score = {i: self._ps(i)["total_score"] for i in ids}
            fooled = 0
            for i in ids:
                p = self._ps(i)
                if p["has_voted"] and p["can_vote"] and not p["is_speaker"]:          # referee_system_prompt_1.txt:45-51
                    if p["vote_choice"] == lie:
                        score[i] += 1
                    else:
                        fooled += 1
            if sp:
                score[sp] += fooled
                upd(sp, "rounds_as_speaker", self._ps(sp)["rounds_as_speaker"] + 1)
            for i in ids:
                if score[i] != self._ps(i)["total_score"]:
                    upd(i, "total_score", score[i])
            note("SCORE_UPDATE", "Round totals: " + ", ".join("Player %d: %d points" % (i, score[i]) for i in ids))
        if en == "T_FINAL":
            best = max(ids, key=lambda i: (self._ps(i)["total_score"], -i))
            note("GAME_STATUS", "Game over - Player %d wins with %d points" % (best, self._ps(best)["total_score"]))

        # winner tag of the branch PhaseNode took (werewolf family)
        if self._phase(Y).get("next_phase") is None and self.rules.get("family") == "werewolf":
            alive_after = lambda i: self._ps(i)["is_alive"] and i not in dead
            wolves_alive = sum(1 for i in ids if self._ps(i)["team"] == wolf_team and alive_after(i))
            note("GAME_STATUS", "Game over - %s win" % (village if wolves_alive == 0 else wolf_team))
        return calls


class _Bound:
    def __init__(self, model: StubChatModel, names: frozenset):
        self.model, self.names = model, names

    async def ainvoke(self, messages, config=None):
        m = self.model
        m.calls += 1
        if self.names == {"update_player_actions"}:
            calls = m.bots()
        elif self.names == {"set_next_phase"}:
            calls = m.phase()
        elif self.names == {"update_player_state", "add_game_note"}:
            calls = m.referee()
        else:
            calls = []
        return AIMessage(content="", tool_calls=calls)
