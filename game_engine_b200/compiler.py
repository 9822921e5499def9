"""DSL -> transition-table compiler (host side, runs once per game).

Input: a game file in the reference's DSL grammar (reference prompt/dsl_declaration_generation_prompt.txt:13-60,
prompt/dsl_phases_generation_prompt.txt:84-149; shipped instances games/werewolf-(mafia).yaml,
games/two-truths-and-a-lie.yaml) plus a rules annotation (`rules/<game>.rules.yaml`) that binds what the
DSL states only in natural language (bot legality, referee effects, branch conditions) to SPEC.md opcodes.
Output: a `CompiledGame` holding the binary table (`table.Table`) and the host-side metadata the adapter needs
(phase ids/names, action text templates, role names).

The reference evaluates all of this with an LLM per step (agent/game_agent_v2.py:1075-1112); here it is
evaluated once, ahead of time.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import yaml

from . import table as T

_HERE = os.path.dirname(os.path.abspath(__file__))
GAMES_DIR = os.path.join(_HERE, "games")
RULES_DIR = os.path.join(_HERE, "rules")


class DSLCompileError(ValueError):
    pass


# --------------------------------------------------------------------------- condition grammar
_TOKEN_RE = re.compile(
    r"\s*(?:(?P<field>player\.[A-Za-z_][A-Za-z0-9_]*)|(?P<op>==|!=|<=|>=|<|>)|(?P<lp>\()|(?P<rp>\))|(?P<lb>\[)|(?P<rb>\])|"
    r"(?P<comma>,)|(?P<str>'[^']*'|\"[^\"]*\")|(?P<num>-?\d+)|(?P<word>[A-Za-z_]+))"
)


def _tokenize(s: str) -> List[Tuple[str, str]]:
    out, pos = [], 0
    s = s.strip()
    while pos < len(s):
        m = _TOKEN_RE.match(s, pos)
        if not m or m.end() == pos:
            raise DSLCompileError("cannot tokenize condition %r at %d" % (s, pos))
        kind = m.lastgroup
        out.append((kind, m.group(kind)))
        pos = m.end()
    return out


class _Parser:
    """expr := term ('or' term)* ; term := factor ('and' factor)* ; factor := 'not' factor | '(' expr ')' | cmp"""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, kind=None, val=None):
        k, v = self.peek()
        if k is None or (kind and k != kind) or (val and v.lower() != val):
            raise DSLCompileError("unexpected token %r" % (v,))
        self.i += 1
        return v

    def expr(self):
        node = self.term()
        while self.peek() == ("word", "or") or (self.peek()[0] == "word" and self.peek()[1].lower() == "or"):
            self.take()
            node = ("or", node, self.term())
        return node

    def term(self):
        node = self.factor()
        while self.peek()[0] == "word" and self.peek()[1].lower() == "and":
            self.take()
            node = ("and", node, self.factor())
        return node

    def literal(self):
        k, v = self.peek()
        if k == "str":
            self.take()
            return v[1:-1]
        if k == "num":
            self.take()
            return int(v)
        if k == "word" and v.lower() in ("true", "false"):
            self.take()
            return v.lower() == "true"
        raise DSLCompileError("expected literal, got %r" % (v,))

    def factor(self):
        k, v = self.peek()
        if k == "word" and v.lower() == "not":
            self.take()
            return ("not", self.factor())
        if k == "lp":
            self.take()
            node = self.expr()
            self.take("rp")
            return node
        if k == "field":
            name = self.take()[len("player."):]
            k2, v2 = self.peek()
            if k2 == "op":
                self.take()
                return ("cmp", name, v2, self.literal())
            if k2 == "word" and v2.lower() == "in":
                self.take()
                self.take("lb")
                vals = [self.literal()]
                while self.peek()[0] == "comma":
                    self.take()
                    vals.append(self.literal())
                self.take("rb")
                return ("in", name, vals)
            return ("cmp", name, "==", True)          # bare boolean field
        raise DSLCompileError("unexpected token %r in condition" % (v,))


def parse_condition(s: str):
    p = _Parser(_tokenize(s))
    node = p.expr()
    if p.i != len(p.t):
        raise DSLCompileError("trailing tokens in condition %r" % s)
    return node


# --------------------------------------------------------------------------- DNF over mask fields
Clause = Tuple[int, int]        # (pos bitset, neg bitset) over mask-field ids


class _FieldMap:
    def __init__(self, family: int, role_names: List[str], wolf_team: str, village_team: str,
                 alias: Optional[Dict[str, str]] = None):
        self.family, self.roles, self.wolf_team, self.village_team = family, role_names, wolf_team, village_team
        self.alias = dict(alias or {})           # DSL field name -> canonical field name (rules `fields:`)
        self.cmps: List[Tuple[int, int, int]] = []      # comparison fields allocated so far (table.Table.cmps)

    def _comparison(self, name: str, op: str, val: Any) -> List[Clause]:
        """`player.<numeric field> <op> <int>` -> a derived mask field (table.py, "comparison fields")."""
        vals = T.W_VAL_FIELDS if self.family == T.FAMILY_WEREWOLF else T.T_VAL_FIELDS
        if isinstance(val, bool) or not isinstance(val, int) or not (0 <= val <= 255):
            raise DSLCompileError("numeric field %r compared with %r (need an integer 0..255)" % (name, val))
        c = (vals[name], T.CMP_OPS[op], int(val))
        if c not in self.cmps:
            if len(self.cmps) >= T.MAX_CMP[self.family]:
                raise DSLCompileError("the table has room for %d numeric comparisons (%r is one too many)" % (T.MAX_CMP[self.family], name))
            self.cmps.append(c)
        return [self._lit(T.cmp_field_id(self.family, self.cmps.index(c)), True)]

    def literal(self, name: str, op: str, val: Any) -> List[Clause]:
        """DNF of `player.<name> <op> <val>`."""
        name = self.alias.get(name, name)
        if name in (T.W_VAL_FIELDS if self.family == T.FAMILY_WEREWOLF else T.T_VAL_FIELDS):
            return self._comparison(name, op, val)
        if op not in ("==", "!="):
            raise DSLCompileError("field %r is not numeric: only == and != apply" % name)
        neg = op == "!="
        if self.family == T.FAMILY_WEREWOLF:
            if name == "role":
                if val not in self.roles:
                    raise DSLCompileError("unknown role %r" % (val,))
                fid = T.W_ROLE_BASE + self.roles.index(val)
                return [self._lit(fid, not neg)]
            if name == "team":
                if val == self.wolf_team:
                    return [self._lit(T.W_FIELDS["team_is_wolf"], not neg)]
                if val == self.village_team:      # a villager is an ASSIGNED non-wolf (team is '' before role assignment)
                    if not neg:
                        return [(1 << T.W_ROLES_ASSIGNED, 1 << T.W_FIELDS["team_is_wolf"])]
                    return [(0, 1 << T.W_ROLES_ASSIGNED), (1 << T.W_FIELDS["team_is_wolf"], 0)]
                raise DSLCompileError("unknown team %r" % (val,))
            fields = T.W_FIELDS
        else:
            fields = T.T_FIELDS
        if name not in fields or not isinstance(val, bool):
            raise DSLCompileError("condition on unsupported field %r (%r)" % (name, val))
        return [self._lit(fields[name], val != neg)]

    @staticmethod
    def _lit(fid: int, positive: bool) -> Clause:
        return (1 << fid, 0) if positive else (0, 1 << fid)


def _dnf(node, fm: _FieldMap, negate=False) -> List[Clause]:
    kind = node[0]
    if kind == "not":
        return _dnf(node[1], fm, not negate)
    if kind == "cmp":
        _, name, op, val = node
        if negate:
            op = T.CMP_NEG[op]
        return fm.literal(name, op, val)
    if kind == "in":
        _, name, vals = node
        if negate:      # not in [..]  ->  AND of !=
            out = [(0, 0)]
            for v in vals:
                out = _and(out, fm.literal(name, "!=", v))
            return out
        out: List[Clause] = []
        for v in vals:
            out += fm.literal(name, "==", v)
        return out
    a, b = _dnf(node[1], fm, negate), _dnf(node[2], fm, negate)
    is_and = (kind == "and") != negate      # De Morgan
    return _and(a, b) if is_and else a + b


def _and(a: List[Clause], b: List[Clause]) -> List[Clause]:
    out = []
    for pa, na in a:
        for pb, nb in b:
            p, n = pa | pb, na | nb
            if p & n:
                continue                     # contradictory clause selects nobody
            out.append((p, n))
    return out


def compile_predicate_chain(cond: str, fm: _FieldMap) -> List[Tuple[int, int, int, int]]:
    """The condition as a run of predicate records: two DNF clauses each, every record but the last flagged
    "continued" (table.PRED_CONTINUED), so the DNF can have any number of clauses."""
    clauses = []
    for c in _dnf(parse_condition(cond), fm):
        if c not in clauses:
            clauses.append(c)
    if not clauses:                              # every clause was contradictory: selects nobody
        clauses = [T.CLAUSE_EMPTY]
    if len(clauses) % 2:
        clauses.append(T.CLAUSE_EMPTY)
    out = []
    for i in range(0, len(clauses), 2):
        cont = T.PRED_CONTINUED if i + 2 < len(clauses) else 0
        out.append((clauses[i][0] | cont, clauses[i][1], clauses[i + 1][0], clauses[i + 1][1]))
    return out


def compile_predicate(cond: str, fm: _FieldMap) -> Tuple[int, int, int, int]:
    """Single-record form (conditions of at most two DNF clauses)."""
    chain = compile_predicate_chain(cond, fm)
    if len(chain) > 1:
        raise DSLCompileError("condition %r needs %d predicate records" % (cond, len(chain)))
    return chain[0]


# canonical per-player keys that are not mask fields (targets of a rules `fields:` alias)
CANONICAL_EXTRA = ("investigated_alignments", "selected_target_id", "role", "team", "name", "statements", "lie_index",
                   "vote_choice", "total_score", "rounds_as_speaker")


# --------------------------------------------------------------------------- compiled game
@dataclass
class CompiledGame:
    name: str
    family: int
    n_players: int
    table: T.Table
    blob: bytes
    phase_ids: List[int]
    phase_names: List[str]
    role_names: List[str]
    teams: Tuple[str, str]                      # (village_team, wolf_team)
    action_text: Dict[int, str]                 # phase index -> action text template
    template: Dict[str, Any]                    # declaration.player_states_template entry
    audience_preds: Dict[str, Tuple[int, int, int, int]] = field(default_factory=dict)     # groups that fit one record
    audience_chains: Dict[str, List[Tuple[int, int, int, int]]] = field(default_factory=dict)    # every compiled group
    audience_errors: Dict[str, str] = field(default_factory=dict)    # groups whose selection_criteria did not compile
    wait_for: Dict[int, str] = field(default_factory=dict)           # phase index -> completion_criteria.wait_for
    field_alias: Dict[str, str] = field(default_factory=dict)      # DSL field name -> canonical field name
    dsl: Dict[str, Any] = field(default_factory=dict)
    # session-level record items that the DSL carries in per-player fields (rules `session_fields:`): the re-vote
    # counter and the tie flag of tables with max_revotes > 0 ("revote_count" / "tie_pending" -> DSL field name)
    session_fields: Dict[str, str] = field(default_factory=dict)

    @property
    def record_size(self) -> int:
        return T.record_size(self.family, self.n_players)

    def index_of(self, phase_id: int) -> int:
        return self.phase_ids.index(int(phase_id))


def load_dsl(game: str, games_dir: Optional[str] = None) -> dict:
    """Same contract as the reference's load_dsl_by_gamename (agent/tools/utils.py:557-581): `<dir>/<game>.yaml`."""
    path = os.path.join(games_dir or GAMES_DIR, game + ".yaml")
    with open(path, encoding="utf-8") as f:
        return yaml.safe_load(f)


def load_rules(game: str, rules_dir: Optional[str] = None) -> dict:
    with open(os.path.join(rules_dir or RULES_DIR, game + ".rules.yaml"), encoding="utf-8") as f:
        return yaml.safe_load(f)


def _template(dsl: dict) -> dict:
    """First entry of declaration.player_states_template (reference utils.py:599-609)."""
    tpl = dsl["declaration"]["player_states_template"]["player_states"]
    return dict(tpl[sorted(tpl.keys(), key=lambda k: int(k))[0]])


def n_wolves_for(rule: Any, n_players: int) -> int:
    if rule == "quarter":
        return max(1, n_players // 4)
    return int(rule)


def compile_game(game: str, n_players: int, dsl: Optional[dict] = None, rules: Optional[dict] = None,
                 max_revotes: Optional[int] = None, strict_audience: bool = False) -> CompiledGame:
    dsl = dsl if dsl is not None else load_dsl(game)
    rules = rules if rules is not None else load_rules(game)
    if max_revotes is None:
        max_revotes = int(rules.get("max_revotes", 0))
    decl = dsl["declaration"]
    fam = {"werewolf": T.FAMILY_WEREWOLF, "ttl": T.FAMILY_TTL}[rules["family"]]
    min_players = int(decl.get("min_players") or 2)
    if not (min_players <= n_players <= 32):
        raise DSLCompileError("%s needs %d..32 players, got %d" % (game, min_players, n_players))

    role_names = [r["name"] for r in decl.get("roles", [])]
    village_team, wolf_team = rules.get("village_team", ""), rules.get("wolf_team", "")
    if fam == T.FAMILY_WEREWOLF:
        want = [rules["roles"][k] for k in ("villager", "werewolf", "doctor", "detective")]
        if role_names != want:
            raise DSLCompileError("werewolf family expects declaration.roles == %r, got %r" % (want, role_names))
    alias = {str(k): str(v) for k, v in (rules.get("fields") or {}).items()}
    fm = _FieldMap(fam, role_names, wolf_team, village_team, alias)

    tpl = _template(dsl)
    fields = T.W_FIELDS if fam == T.FAMILY_WEREWOLF else T.T_FIELDS
    for dsl_name, canon in alias.items():
        if dsl_name not in tpl:
            raise DSLCompileError("rules alias %r is not a field of the DSL's player_states_template" % dsl_name)
        if canon not in fields and canon not in CANONICAL_EXTRA:
            raise DSLCompileError("rules alias %r -> %r: unknown canonical field" % (dsl_name, canon))
    session_fields = {str(k): str(v) for k, v in (rules.get("session_fields") or {}).items()}
    for item, dsl_name in session_fields.items():
        if item not in ("revote_count", "tie_pending"):
            raise DSLCompileError("rules session_fields: unknown item %r" % item)
        if dsl_name not in tpl:
            raise DSLCompileError("rules session_fields %r -> %r: not a field of the DSL's player_states_template" % (item, dsl_name))
    if fam == T.FAMILY_WEREWOLF and max_revotes > 0 and set(session_fields) != {"revote_count", "tie_pending"}:
        # without them the dict state cannot say "a tied vote is pending": the drop-in nodes would silently take the
        # no-tie branch (the record is rebuilt from the dict on every step)
        raise DSLCompileError("a table with max_revotes > 0 needs rules session_fields for revote_count and tie_pending")
    init_masks = 0
    for name, val in tpl.items():
        fid = fields.get(alias.get(name, name))
        if fid is not None and val is True:
            init_masks |= 1 << fid

    phases_dsl = dsl["phases"]
    ids = sorted(int(k) for k in phases_dsl.keys())
    if ids[0] != 0:
        raise DSLCompileError("phase 0 must exist (reference starts at current_phase_id 0, game_agent_v2.py:103)")
    index = {pid: i for i, pid in enumerate(ids)}
    get = lambda pid: phases_dsl.get(pid) if pid in phases_dsl else phases_dsl.get(str(pid))

    preds: List[tuple] = []

    def pred_index(cond: str) -> int:
        chain = compile_predicate_chain(cond, fm)
        for i in range(len(preds) - len(chain) + 1):              # an identical run is shared
            if preds[i:i + len(chain)] == chain and (i == 0 or not (preds[i - 1][0] & T.PRED_CONTINUED)):
                return i
        preds.extend(chain)
        return len(preds) - len(chain)

    n_wolves = n_wolves_for(rules.get("wolves", 0), n_players) if fam == T.FAMILY_WEREWOLF else 0
    if fam == T.FAMILY_WEREWOLF and n_wolves + 2 > n_players:
        raise DSLCompileError("not enough players for %d wolves + Doctor + Detective" % n_wolves)
    rounds = int(rules.get("rounds", 0))
    tab = T.Table(family=fam, n_players=n_players, n_wolves=n_wolves, rounds=rounds,
                  max_revotes=max_revotes, init_masks=init_masks)
    action_text: Dict[int, str] = {}
    wait_for: Dict[int, str] = {}
    prules = {int(k): v for k, v in (rules.get("phases") or {}).items()}
    kind_map = {"UI_displayed": T.KIND_UI, "timer": T.KIND_TIMER, "player_action": T.KIND_ACTION}

    for pid in ids:
        ph = get(pid)
        pr = prules.get(pid, {})
        cc = ph.get("completion_criteria") or {}
        nxt = ph.get("next_phase")
        kind = T.KIND_TERMINAL if nxt is None else kind_map[cc.get("type")]
        out = T.Phase(id=pid, kind=kind)
        out.entry_op = T.ENTRY_OPS[pr.get("entry", "NONE")]
        out.exit_op = T.EXIT_OPS[pr.get("exit", "NONE")]
        if kind == T.KIND_ACTION:
            cond = (cc.get("target_players") or {}).get("condition")
            if not cond:
                raise DSLCompileError("player_action phase %d has no target_players.condition" % pid)
            out.actor_pred = pred_index(cond)
            # wait_for (grammar prompt/dsl_phases_generation_prompt.txt:110-113) names three kinds of waiting; the
            # completion logic is the same for all of them (:128 "whether feedback has been collected from all
            # target_players"), which is the rule of SPEC D3h.  An unknown value is a malformed game file.
            wf = cc.get("wait_for")
            if wf is not None:
                if wf not in ("single_player_choice", "all_players_action", "multiple_players_action"):
                    raise DSLCompileError("phase %d: unknown completion_criteria.wait_for %r" % (pid, wf))
                wait_for[index[pid]] = wf
            act = pr.get("action")
            if not act:
                raise DSLCompileError("rules annotation lacks an action for player_action phase %d" % pid)
            out.action_op = T.ACTION_OPS[act["op"]]
            if out.action_op == T.ACT_PICK_PLAYER:
                out.action_arg = pred_index(act["legal"])
                out.action_flags = T.ACTF_EXCLUDE_SELF if act.get("exclude_self") else 0
            elif out.action_op == T.ACT_PICK_OPTION:
                out.action_arg = int(act["options"])
            action_text[index[pid]] = act.get("text", "acted")
        elif "action" in pr or "exit" in pr:
            raise DSLCompileError("phase %d is not a player_action phase but the rules give it an action" % pid)

        # next_phase: simple {id,name} or ordered natural-language branches (dsl_phases_generation_prompt.txt:129-149)
        if nxt is None:
            pass
        elif isinstance(nxt, dict) and "id" in nxt and not isinstance(nxt.get("id"), dict):
            out.branches.append(T.Branch(T.BR_ALWAYS, index[int(nxt["id"])]))
        else:
            ann = pr.get("branches")
            keys = list(nxt.keys())
            if not ann or [a.get("key") for a in ann] != keys:
                raise DSLCompileError("phase %d: rules annotation must list the branch keys %r in order" % (pid, keys))
            for a in ann:
                tgt = nxt[a["key"]]
                b = T.Branch(T.BRANCH_OPS[a["op"]], index[int(tgt["id"])], int(a.get("tag", 0)))
                if b.op == T.BR_COUNT_EQ0:
                    b.a = pred_index(a["a"])
                elif b.op == T.BR_COUNT_GE:
                    b.a, b.arg = pred_index(a["a"]), pred_index(a["b"])
                elif b.op == T.BR_PREV_IN:
                    for p in a["phases"]:
                        b.arg |= 1 << index[int(p)]
                elif b.op == T.BR_ALL_VAL_GE:
                    b.a = T.T_VAL_FIELDS[a["field"]]
                    b.arg = rounds if a["value"] == "rounds" else int(a["value"])
                out.branches.append(b)
        # every id referenced must exist: the reference validates ids the same way (game_agent_v2.py:1173-1204)
        tab.phases.append(out)

    # audience groups (declaration.audience_groups[*].selection_criteria): compiled after the phases, so the comparison
    # fields they allocate never displace one a phase needs.  A group that does not compile is REPORTED
    # (audience_errors; strict_audience=True raises) instead of silently missing from the masks.
    aud, aud_chains, aud_err = {}, {}, {}
    for gname, g in (decl.get("audience_groups") or {}).items():
        saved = list(fm.cmps)
        try:
            chain = compile_predicate_chain(g["selection_criteria"], fm)
            aud_chains[gname] = chain
            if len(chain) == 1:
                aud[gname] = chain[0]
        except (DSLCompileError, KeyError, TypeError) as e:
            fm.cmps[:] = saved
            if strict_audience:
                raise DSLCompileError("audience group %r: %s" % (gname, e))
            aud_err[gname] = str(e)
    tab.preds = preds
    tab.cmps = list(fm.cmps)
    return CompiledGame(
        name=game, family=fam, n_players=n_players, table=tab, blob=tab.pack(), phase_ids=ids,
        phase_names=[get(p).get("name", "Phase %d" % p) for p in ids], role_names=role_names,
        teams=(village_team, wolf_team), action_text=action_text, template=tpl, audience_preds=aud, audience_chains=aud_chains,
        audience_errors=aud_err, wait_for=wait_for, dsl=dsl,
        field_alias=alias, session_fields=session_fields,
    )
