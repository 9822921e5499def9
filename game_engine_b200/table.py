"""Binary transition-table format shared by the CUDA library, the C oracle and the host code.

This is the compiled form of the reference's phase graph (`games/*.yaml` `phases:` section,
reference games/werewolf-(mafia).yaml:166-666, games/two-truths-and-a-lie.yaml:145-403; accessor
agent/tools/utils.py:19-31) plus the SPEC.md opcodes.  Layout (little-endian), mirrored by
`include/game_engine_b200.h` (`ge_table_header_t`, `ge_phase_t`, `ge_pred_t`):

    header  32 B : "GETB" u16 version | u8 family | u8 n_phases | u8 n_players | u8 n_preds |
                   u8 n_wolves | u8 rounds | u8 max_revotes | u8 n_cmp | u8 reserved[2] |
                   u32 init_masks (bit f: mask field f starts as "all players") |
                   4 x cmp { u8 value_field | u8 op | u8 constant }      (comparison fields, see below)
    phase   48 B : u8 id | u8 kind | u8 action_op | u8 action_arg | u8 action_flags | u8 exit_op |
                   u8 entry_op | u8 n_branches | u8 actor_pred | u8 pad[7] |
                   4 x branch { u8 op | u8 next | u8 tag | u8 a | u32 arg }
    pred     8 B : u16 pos0 | u16 neg0 | u16 pos1 | u16 neg1     (DNF, two clauses; bit 15 of pos0 = "continued":
                   the predicate is this record OR the next one, so a DNF of any length is a run of records)

Comparison fields: the DSL's numeric conditions (`player.total_score >= 3`, `player.selected_target_id != 0`;
grammar prompt/dsl_phases_generation_prompt.txt:106-128) compile to up to n_cmp derived MASK fields — bit p = "player
p's value field <op> constant" — that predicates use like any other mask field.  Comparison k has mask-field id
13 + k in the werewolf family (k < 2; value field 0 = selected_target_id) and 11 + k in the TTL family (k < 4; value
fields 0 total_score, 1 rounds_as_speaker, 2 vote_choice).  ops: 0 ==, 1 !=, 2 <, 3 <=, 4 >, 5 >=.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List

MAGIC = b"GETB"
VERSION = 1
MAX_PHASES = 32
MAX_PREDS = 32
MAX_BRANCHES = 4
STATS_LEN = 560

FAMILY_WEREWOLF = 1
FAMILY_TTL = 2

KIND_UI, KIND_TIMER, KIND_ACTION, KIND_TERMINAL = 0, 1, 2, 3

ACT_NONE, ACT_PICK_PLAYER, ACT_PICK_OPTION, ACT_MARK = 0, 1, 2, 3
ACTF_EXCLUDE_SELF = 1

EX_NONE = 0
EX_VOTE_KILL, EX_PROTECT, EX_INVESTIGATE_RESOLVE, EX_DAY_VOTE = 1, 2, 3, 4
EX_T_STATEMENTS, EX_T_LIE, EX_T_VOTES = 16, 17, 18

EN_NONE = 0
EN_ASSIGN_ROLES, EN_NIGHT_RESET = 1, 2
EN_T_ROUND_START, EN_T_REVEAL, EN_T_SCORE, EN_T_FINAL = 16, 17, 18, 19

BR_ALWAYS, BR_COUNT_EQ0, BR_COUNT_GE, BR_PREV_IN, BR_ALL_VAL_GE, BR_TIE_PENDING = 0, 1, 2, 3, 4, 5

EXIT_OPS = {
    "NONE": EX_NONE, "VOTE_KILL": EX_VOTE_KILL, "PROTECT": EX_PROTECT,
    "INVESTIGATE_RESOLVE": EX_INVESTIGATE_RESOLVE, "DAY_VOTE": EX_DAY_VOTE,
    "T_STATEMENTS": EX_T_STATEMENTS, "T_LIE": EX_T_LIE, "T_VOTES": EX_T_VOTES,
}
ENTRY_OPS = {
    "NONE": EN_NONE, "ASSIGN_ROLES": EN_ASSIGN_ROLES, "NIGHT_RESET": EN_NIGHT_RESET,
    "T_ROUND_START": EN_T_ROUND_START, "T_REVEAL": EN_T_REVEAL, "T_SCORE": EN_T_SCORE,
    "T_FINAL": EN_T_FINAL,
}
ACTION_OPS = {"NONE": ACT_NONE, "PICK_PLAYER": ACT_PICK_PLAYER, "PICK_OPTION": ACT_PICK_OPTION, "MARK": ACT_MARK}
BRANCH_OPS = {
    "ALWAYS": BR_ALWAYS, "COUNT_EQ0": BR_COUNT_EQ0, "COUNT_GE": BR_COUNT_GE, "PREV_IN": BR_PREV_IN,
    "ALL_VAL_GE": BR_ALL_VAL_GE, "TIE_PENDING": BR_TIE_PENDING,
}

# mask-field ids (SPEC.md section 2)
F_ALL = 15
W_FIELDS = {
    "is_alive": 0, "can_vote": 1, "night_action_eligible": 2, "night_action_submitted": 3,
    "role_revealed": 4, "investigated": 5, "team_is_wolf": 6, "has_secret_role": 7,
}
W_ROLES_ASSIGNED = 12      # derived: all players once the session has any secret role (roles are assigned), else nobody
W_ROLE_BASE = 8
T_FIELDS = {"is_speaker": 0, "statements_submitted": 1, "lie_revealed": 2, "can_vote": 3, "has_voted": 4}
# per-player value fields usable by ALL_VAL_GE
T_VAL_FIELDS = {"total_score": 0, "rounds_as_speaker": 1, "vote_choice": 2}

# comparison fields
CMP_OPS = {"==": 0, "!=": 1, "<": 2, "<=": 3, ">": 4, ">=": 5}
CMP_NEG = {"==": "!=", "!=": "==", "<": ">=", ">=": "<", "<=": ">", ">": "<="}
W_VAL_FIELDS = {"selected_target_id": 0}
MAX_CMP = {FAMILY_WEREWOLF: 2, FAMILY_TTL: 4}
PRED_CONTINUED = 1 << 15       # in pos0: OR with the next predicate record


def cmp_field_id(family: int, k: int) -> int:
    return (13 if family == FAMILY_WEREWOLF else 11) + k


def cmp_holds(op: int, value: int, const: int) -> bool:
    return [value == const, value != const, value < const, value <= const, value > const, value >= const][op]


PRED_NONE = 0xFF
CLAUSE_EMPTY = (0, 1 << F_ALL)      # "& ~ALL" selects nobody: marks an unused clause

HEADER_FMT = "<4sHBBBBBBBB2sI12s"
PHASE_HEAD_FMT = "<9B7x"
BRANCH_FMT = "<BBBBI"
PRED_FMT = "<4H"
HEADER_SIZE = struct.calcsize(HEADER_FMT)
PHASE_SIZE = struct.calcsize(PHASE_HEAD_FMT) + MAX_BRANCHES * struct.calcsize(BRANCH_FMT)
PRED_SIZE = struct.calcsize(PRED_FMT)
assert HEADER_SIZE == 32 and PHASE_SIZE == 48 and PRED_SIZE == 8


@dataclass
class Branch:
    op: int = BR_ALWAYS
    next: int = 0          # phase INDEX
    tag: int = 0
    a: int = 0
    arg: int = 0


@dataclass
class Phase:
    id: int
    kind: int
    action_op: int = ACT_NONE
    action_arg: int = 0
    action_flags: int = 0
    exit_op: int = EX_NONE
    entry_op: int = EN_NONE
    actor_pred: int = PRED_NONE
    branches: List[Branch] = field(default_factory=list)


@dataclass
class Table:
    family: int
    n_players: int
    n_wolves: int = 0
    rounds: int = 0
    max_revotes: int = 0
    init_masks: int = 0
    phases: List[Phase] = field(default_factory=list)
    preds: List[tuple] = field(default_factory=list)     # (pos0, neg0, pos1, neg1)
    cmps: List[tuple] = field(default_factory=list)      # (value_field, op, constant): comparison fields

    def pack(self) -> bytes:
        if not (0 < len(self.phases) <= MAX_PHASES):
            raise ValueError("phase count out of range")
        if len(self.preds) > MAX_PREDS:
            raise ValueError("too many predicates")
        if not (2 <= self.n_players <= 32):
            raise ValueError("n_players must be 2..32")
        if len(self.cmps) > MAX_CMP.get(self.family, 0):
            raise ValueError("too many comparison fields")
        cmp_bytes = b"".join(bytes(c) for c in self.cmps).ljust(12, b"\0")
        out = [struct.pack(HEADER_FMT, MAGIC, VERSION, self.family, len(self.phases), self.n_players,
                           len(self.preds), self.n_wolves, self.rounds, self.max_revotes, len(self.cmps), b"\0\0",
                           self.init_masks, cmp_bytes)]
        for ph in self.phases:
            if len(ph.branches) > MAX_BRANCHES:
                raise ValueError("too many branches in phase %d" % ph.id)
            out.append(struct.pack(PHASE_HEAD_FMT, ph.id, ph.kind, ph.action_op, ph.action_arg, ph.action_flags,
                                   ph.exit_op, ph.entry_op, len(ph.branches), ph.actor_pred))
            for i in range(MAX_BRANCHES):
                b = ph.branches[i] if i < len(ph.branches) else Branch()
                out.append(struct.pack(BRANCH_FMT, b.op, b.next, b.tag, b.a, b.arg))
        for p in self.preds:
            out.append(struct.pack(PRED_FMT, *p))
        return b"".join(out)

    @staticmethod
    def unpack(blob: bytes) -> "Table":
        magic, ver, fam, nph, npl, npr, nw, rounds, mrv, ncmp, _, init_masks, cmpb = struct.unpack_from(HEADER_FMT, blob, 0)
        if magic != MAGIC or ver != VERSION:
            raise ValueError("bad table blob")
        t = Table(family=fam, n_players=npl, n_wolves=nw, rounds=rounds, max_revotes=mrv, init_masks=init_masks)
        t.cmps = [tuple(cmpb[3 * k: 3 * k + 3]) for k in range(min(ncmp, 4))]
        off = HEADER_SIZE
        for _ in range(nph):
            pid, kind, aop, aarg, afl, exo, eno, nbr, apred = struct.unpack_from(PHASE_HEAD_FMT, blob, off)
            boff = off + struct.calcsize(PHASE_HEAD_FMT)
            brs = []
            for i in range(nbr):
                op, nxt, tag, a, arg = struct.unpack_from(BRANCH_FMT, blob, boff + 8 * i)
                brs.append(Branch(op, nxt, tag, a, arg))
            t.phases.append(Phase(pid, kind, aop, aarg, afl, exo, eno, apred, brs))
            off += PHASE_SIZE
        for _ in range(npr):
            t.preds.append(struct.unpack_from(PRED_FMT, blob, off))
            off += PRED_SIZE
        return t


def record_size(family: int, n_players: int) -> int:
    """Canonical packed record size S in bytes (SPEC.md section 5)."""
    if family == FAMILY_WEREWOLF:
        return 48 + ((n_players + 7) // 8) * 8
    if family == FAMILY_TTL:
        return ((8 + 4 * n_players + 7) // 8) * 8
    raise ValueError("unknown family")
