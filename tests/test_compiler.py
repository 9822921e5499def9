"""DSL -> transition-table compiler: grammar, DNF predicates, table layout, reference-file equivalence."""
import os
import struct

import pytest
import yaml

from conftest import REFERENCE, has_reference
from game_engine_b200 import compile_game, DSLCompileError
from game_engine_b200 import compiler as C
from game_engine_b200 import table as T

WEREWOLF, TTL = "werewolf-(mafia)", "two-truths-and-a-lie"


def test_werewolf_table_matches_hand_written_expectation(games):
    cg = games(WEREWOLF, 8)
    t = cg.table
    assert cg.phase_ids == list(range(17)) + [99]
    assert cg.record_size == 56 and t.n_wolves == 2 and t.family == T.FAMILY_WEREWOLF
    assert t.init_masks == 0b11                       # is_alive, can_vote start true (template)
    kinds = [p.kind for p in t.phases]
    U, M, A, X = T.KIND_UI, T.KIND_TIMER, T.KIND_ACTION, T.KIND_TERMINAL
    assert kinds == [U, U, A, A, A, U, M, A, U, U, A, A, A, U, M, A, U, X]
    assert [p.exit_op for p in t.phases] == [0, 0, 1, 2, 3, 0, 0, 4, 0, 0, 1, 2, 3, 0, 0, 4, 0, 0]
    assert [p.entry_op for p in t.phases] == [0, 1, 2, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0]
    # simple edges
    nxt = {p.id: [cg.phase_ids[b.next] for b in p.branches] for p in t.phases}
    assert nxt[0] == [1] and nxt[8] == [9] and nxt[13] == [9] and nxt[16] == [9] and nxt[99] == []
    # the win-check phase: ordered branches, first match wins
    br = t.phases[9].branches
    assert [b.op for b in br] == [T.BR_COUNT_EQ0, T.BR_COUNT_GE, T.BR_PREV_IN, T.BR_PREV_IN]
    assert nxt[9] == [99, 99, 10, 14] and [b.tag for b in br] == [1, 2, 0, 0]
    assert br[2].arg == (1 << 8) | (1 << 16) and br[3].arg == 1 << 13
    wolves_alive = (1 << 6 | 1 << 0, 0) + T.CLAUSE_EMPTY
    villagers_alive = (1 << 0 | 1 << 12, 1 << 6) + T.CLAUSE_EMPTY          # assigned (12), not wolf (6), alive (0)
    assert t.preds[br[0].a] == wolves_alive and t.preds[br[1].a] == wolves_alive and t.preds[br[1].arg] == villagers_alive
    # actor predicates: role == 'Werewolf' and alive; can_vote and alive
    assert t.preds[t.phases[2].actor_pred] == (1 << 9 | 1, 0) + T.CLAUSE_EMPTY
    assert t.preds[t.phases[7].actor_pred] == (0b11, 0) + T.CLAUSE_EMPTY
    assert t.phases[4].action_flags == T.ACTF_EXCLUDE_SELF and t.phases[3].action_flags == 0


def test_ttl_table_matches_hand_written_expectation(games):
    cg = games(TTL, 4)
    t = cg.table
    assert cg.phase_ids == list(range(9)) + [99] and cg.record_size == 24 and t.rounds == 1
    assert t.init_masks == 1 << 3                      # can_vote
    assert [p.kind for p in t.phases] == [0, 0, 2, 2, 1, 2, 0, 0, 0, 3]
    assert [p.action_op for p in t.phases] == [0, 0, T.ACT_MARK, T.ACT_PICK_OPTION, 0, T.ACT_PICK_OPTION, 0, 0, 0, 0]
    assert [p.entry_op for p in t.phases] == [0, 16, 0, 0, 0, 0, 17, 18, 0, 19]
    b = t.phases[8].branches
    assert [(x.op, cg.phase_ids[x.next]) for x in b] == [(T.BR_ALL_VAL_GE, 99), (T.BR_ALWAYS, 1)]
    assert b[0].a == T.T_VAL_FIELDS["rounds_as_speaker"] and b[0].arg == 1
    assert t.preds[t.phases[5].actor_pred] == (1 << 3, 1 << 0) + T.CLAUSE_EMPTY     # not speaker and can_vote


def test_blob_layout_round_trip(games):
    for cg in (games(WEREWOLF, 8), games(WEREWOLF, 32), games(TTL, 4)):
        blob = cg.blob
        assert blob[:4] == b"GETB" and struct.unpack_from("<H", blob, 4)[0] == 1
        assert len(blob) == 32 + 48 * len(cg.table.phases) + 8 * len(cg.table.preds)
        assert T.Table.unpack(blob).pack() == blob


@pytest.mark.parametrize("P,W", [(4, 1), (7, 1), (8, 2), (16, 4), (32, 8)])
def test_wolf_count_rule(P, W):
    assert compile_game(WEREWOLF, P).table.n_wolves == W


def test_record_sizes():
    assert [T.record_size(T.FAMILY_WEREWOLF, p) for p in (8, 16, 32, 5)] == [56, 64, 80, 56]
    assert [T.record_size(T.FAMILY_TTL, p) for p in (4, 3, 8, 32)] == [24, 24, 40, 136]


def test_condition_grammar():
    fm = C._FieldMap(T.FAMILY_WEREWOLF, ["Villager", "Werewolf", "Doctor", "Detective"], "werewolves", "villagers")
    cp = lambda s: C.compile_predicate(s, fm)
    assert cp("player.is_alive == true") == (1, 0) + T.CLAUSE_EMPTY
    assert cp("player.is_alive == false") == (0, 1) + T.CLAUSE_EMPTY
    assert cp("player.is_alive != true") == (0, 1) + T.CLAUSE_EMPTY
    assert cp("player.team == 'villagers' and player.is_alive == true") == (1 | 1 << 12, 1 << 6) + T.CLAUSE_EMPTY
    assert cp("player.team != 'villagers'") == (0, 1 << 12, 1 << 6, 0)        # unassigned, or a wolf
    assert cp("player.role in ['Doctor', 'Detective'] and player.is_alive == true") == (1 << 10 | 1, 0, 1 << 11 | 1, 0)
    assert cp("not (player.is_alive == true or player.can_vote == true)") == (0, 0b11) + T.CLAUSE_EMPTY
    assert cp("(player.role == 'Doctor' or player.role == 'Detective') and player.is_alive") == (1 << 10 | 1, 0, 1 << 11 | 1, 0)
    with pytest.raises(DSLCompileError):
        cp("player.role == 'Seer'")
    with pytest.raises(DSLCompileError):
        cp("player.is_alive === true")
    with pytest.raises(DSLCompileError):
        cp("player.role in ['Doctor','Detective','Villager']")      # needs 3 clauses


def test_audience_groups_compile(games):
    aud = games(WEREWOLF, 8).audience_preds
    assert set(aud) == {"werewolves", "villagers", "alive_players", "dead_players", "special_roles", "night_actors",
                        "voters", "secret_holders"}
    assert aud["dead_players"] == (0, 1) + T.CLAUSE_EMPTY


def test_branch_keys_must_be_annotated_in_order():
    dsl = C.load_dsl(WEREWOLF)
    rules = C.load_rules(WEREWOLF)
    rules["phases"][9]["branches"] = list(reversed(rules["phases"][9]["branches"]))
    with pytest.raises(DSLCompileError):
        compile_game(WEREWOLF, 8, dsl=dsl, rules=rules)


def test_player_count_limits():
    with pytest.raises(DSLCompileError):
        compile_game(WEREWOLF, 3)            # min_players 4
    with pytest.raises(DSLCompileError):
        compile_game(TTL, 33)


@pytest.mark.reference
@pytest.mark.skipif(not has_reference(), reason="needs /root/reference")
@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 32), (TTL, 4)])
def test_reference_yaml_compiles_to_identical_table(game, P):
    """The shipped condensed game files lose nothing the compiler consumes."""
    with open(os.path.join(REFERENCE, "games", game + ".yaml"), encoding="utf-8") as f:
        ref = yaml.safe_load(f)
    a, b = compile_game(game, P, dsl=ref), compile_game(game, P)
    assert a.blob == b.blob and a.phase_names == b.phase_names and a.template == b.template


def test_revote_variant_compiles_with_tie_branches():
    cg = compile_game("werewolf-revote", 32)
    t = cg.table
    assert t.max_revotes == 2 and cg.phase_ids[-3:] == [17, 18, 99]
    for announce, revote in ((8, 17), (16, 18)):
        br = t.phases[cg.index_of(announce)].branches
        assert [(b.op, cg.phase_ids[b.next]) for b in br] == [(T.BR_TIE_PENDING, revote), (T.BR_ALWAYS, 9)]
        rv = t.phases[cg.index_of(revote)]
        assert rv.kind == T.KIND_ACTION and rv.exit_op == T.EX_DAY_VOTE and cg.phase_ids[rv.branches[0].next] == announce


DRAFT = "werewolf-draft"


def test_draft_variant_is_a_third_table(games):
    """The reference's earlier 13-phase werewolf generation (game_draft/): other phase graph, win check after
    every dawn, two terminal phases, and a state schema bound through the rules' `fields:` aliases."""
    cg = games(DRAFT, 8)
    t = cg.table
    assert cg.phase_ids == list(range(11)) + [98, 99] and cg.record_size == 56 and t.n_wolves == 2
    U, M, A, X = T.KIND_UI, T.KIND_TIMER, T.KIND_ACTION, T.KIND_TERMINAL
    assert [p.kind for p in t.phases] == [U, U, A, A, A, U, U, M, A, U, U, X, X]
    assert [p.exit_op for p in t.phases] == [0, 0, 1, 2, 3, 0, 0, 0, 4, 0, 0, 0, 0]
    assert [p.entry_op for p in t.phases] == [0, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    br = t.phases[10].branches
    assert [(b.op, cg.phase_ids[b.next], b.tag) for b in br] == [
        (T.BR_COUNT_EQ0, 98, 1), (T.BR_COUNT_GE, 99, 2), (T.BR_PREV_IN, 7, 0), (T.BR_ALWAYS, 2, 0)]
    assert br[2].arg == 1 << 6                       # "follows Dawn Reveal"
    assert t.init_masks == 0b11
    # aliases: has_night_action -> has_secret_role (7), wolf_chat_enabled -> team_is_wolf (6)
    assert cg.field_alias == {"has_night_action": "has_secret_role", "known_alignments": "investigated_alignments",
                              "wolf_chat_enabled": "team_is_wolf"}
    assert cg.audience_preds["night_actors"] == (1 << 7 | 1, 0) + T.CLAUSE_EMPTY
    assert cg.audience_preds["wolf_chatters"] == (1 << 6 | 1, 0) + T.CLAUSE_EMPTY
    assert len(cg.audience_preds) == 9


def test_alias_must_name_real_fields():
    dsl, rules = C.load_dsl(DRAFT), C.load_rules(DRAFT)
    rules["fields"] = {"has_night_action": "no_such_field"}
    with pytest.raises(DSLCompileError):
        compile_game(DRAFT, 8, dsl=dsl, rules=rules)
    rules["fields"] = {"not_in_template": "has_secret_role"}
    with pytest.raises(DSLCompileError):
        compile_game(DRAFT, 8, dsl=dsl, rules=rules)


@pytest.mark.reference
@pytest.mark.skipif(not has_reference(), reason="needs /root/reference")
def test_reference_draft_yaml_compiles_to_identical_table():
    with open(os.path.join(REFERENCE, "game_draft", "werewolf-(mafia).yaml"), encoding="utf-8") as f:
        ref = yaml.safe_load(f)
    a, b = compile_game(DRAFT, 12, dsl=ref), compile_game(DRAFT, 12)
    assert a.blob == b.blob and a.phase_names == b.phase_names and a.template == b.template and a.audience_preds == b.audience_preds


# ---- property: the DNF compiler agrees with a direct evaluation of the condition on random player states
def _random_condition(rng, depth=0):
    fields = ["is_alive", "can_vote", "night_action_eligible", "night_action_submitted", "role_revealed", "has_secret_role"]
    r = rng.random()
    if depth >= 2 or r < 0.45:
        k = rng.integers(0, 4)
        if k == 0:
            return "player.%s %s %s" % (fields[rng.integers(0, len(fields))], ["==", "!="][rng.integers(0, 2)], ["true", "false"][rng.integers(0, 2)])
        if k == 1:
            return "player.role %s '%s'" % (["==", "!="][rng.integers(0, 2)], ["Villager", "Werewolf", "Doctor", "Detective"][rng.integers(0, 4)])
        if k == 2:
            return "player.team %s '%s'" % (["==", "!="][rng.integers(0, 2)], ["werewolves", "villagers"][rng.integers(0, 2)])
        roles = ["Villager", "Werewolf", "Doctor", "Detective"]
        pick = [roles[i] for i in sorted(rng.choice(4, size=2, replace=False))]
        return "player.role in ['%s', '%s']" % tuple(pick)
    if r < 0.6:
        return "not (%s)" % _random_condition(rng, depth + 1)
    op = "and" if r < 0.85 else "or"
    return "(%s) %s (%s)" % (_random_condition(rng, depth + 1), op, _random_condition(rng, depth + 1))


def test_dnf_compiler_agrees_with_direct_evaluation():
    import re
    import numpy as np
    roles = ["Villager", "Werewolf", "Doctor", "Detective"]
    fm = C._FieldMap(T.FAMILY_WEREWOLF, roles, "werewolves", "villagers")
    rng = np.random.default_rng(2026)
    checked = 0
    for _ in range(600):
        cond = _random_condition(rng)
        try:
            pred = C.compile_predicate(cond, fm)
        except DSLCompileError:
            continue                                  # more than two DNF clauses: rejected, never mis-compiled
        src = re.sub(r"\btrue\b", "True", re.sub(r"\bfalse\b", "False", cond))
        for _ in range(24):
            assigned = bool(rng.integers(0, 2))
            role = roles[rng.integers(0, 4)] if assigned else ""
            st = {"is_alive": bool(rng.integers(0, 2)), "can_vote": bool(rng.integers(0, 2)),
                  "night_action_eligible": bool(rng.integers(0, 2)), "night_action_submitted": bool(rng.integers(0, 2)),
                  "role_revealed": bool(rng.integers(0, 2)), "has_secret_role": assigned and role != "Villager",
                  "role": role, "team": ("werewolves" if role == "Werewolf" else "villagers") if assigned else ""}
            want = bool(eval(src, {"__builtins__": {}}, {"player": type("P", (), st)}))
            # mask fields of ONE player (bit 0), SPEC section 2
            f = {0: st["is_alive"], 1: st["can_vote"], 2: st["night_action_eligible"], 3: st["night_action_submitted"],
                 4: st["role_revealed"], 5: False, 6: st["team"] == "werewolves", 7: st["has_secret_role"],
                 8: assigned and role == "Villager", 9: role == "Werewolf", 10: role == "Doctor", 11: role == "Detective",
                 12: assigned, 15: True}
            got = False
            for pos, neg in ((pred[0], pred[1]), (pred[2], pred[3])):
                ok = all(f.get(b, False) for b in range(16) if (pos >> b) & 1) and not any(f.get(b, False) for b in range(16) if (neg >> b) & 1)
                got = got or ok
            assert got == want, (cond, st, pred)
            checked += 1
    assert checked > 3000


# ------------------------------------------------------------------ grammar beyond the shipped games
def _fm(family):
    from game_engine_b200.compiler import _FieldMap
    if family == T.FAMILY_WEREWOLF:
        return _FieldMap(family, ["Villager", "Werewolf", "Doctor", "Detective"], "werewolves", "villagers")
    return _FieldMap(family, [], "", "")


def test_numeric_comparisons_become_comparison_fields():
    """`< <= > >= == !=` on the per-player value fields (grammar: prompt/dsl_phases_generation_prompt.txt:106-128)."""
    from game_engine_b200.compiler import compile_predicate_chain
    fm = _fm(T.FAMILY_TTL)
    chain = compile_predicate_chain("player.total_score >= 3 and player.can_vote == true", fm)
    assert fm.cmps == [(0, T.CMP_OPS[">="], 3)] and chain == [((1 << 11) | (1 << 3), 0) + T.CLAUSE_EMPTY]
    # a negated comparison flips the operator instead of needing a negative literal
    compile_predicate_chain("not (player.rounds_as_speaker < 1)", fm)
    assert fm.cmps[-1] == (1, T.CMP_OPS[">="], 1)
    # the same comparison is allocated once
    compile_predicate_chain("player.total_score >= 3 or player.vote_choice != 0", fm)
    assert fm.cmps == [(0, 5, 3), (1, 5, 1), (2, 1, 0)]
    fw = _fm(T.FAMILY_WEREWOLF)
    assert compile_predicate_chain("player.selected_target_id > 0 and player.is_alive == true", fw) == [((1 << 13) | 1, 0) + T.CLAUSE_EMPTY]
    with pytest.raises(DSLCompileError):
        compile_predicate_chain("player.is_alive >= 1", fw)                      # ordering on a boolean field
    with pytest.raises(DSLCompileError):
        compile_predicate_chain("player.selected_target_id < 'x'", fw)
    for k in range(2, 4):                                                         # the werewolf family has two slots
        try:
            compile_predicate_chain("player.selected_target_id == %d" % k, fw)
        except DSLCompileError:
            assert k == 3
            break
    else:
        raise AssertionError("a third comparison field was accepted")


def test_conditions_of_any_length_chain_predicate_records():
    from game_engine_b200.compiler import compile_predicate_chain
    fw = _fm(T.FAMILY_WEREWOLF)
    chain = compile_predicate_chain("player.role in ['Werewolf', 'Doctor', 'Detective'] and player.is_alive == true", fw)
    assert len(chain) == 2 and chain[0][0] & T.PRED_CONTINUED and not chain[1][0] & T.PRED_CONTINUED
    chain = compile_predicate_chain("(player.is_alive == true or player.role_revealed == true) and (player.can_vote == true or player.role == 'Doctor') "
                                    "and player.team != 'werewolves'", fw)
    assert len(chain) >= 2 and all(c[0] & T.PRED_CONTINUED for c in chain[:-1])
    # the evaluation of a chained predicate == Python's evaluation of the text, on every combination of the fields
    from game_engine_b200.adapter import Record
    cg = compile_game("werewolf-(mafia)", 8)
    import itertools
    import numpy as np
    from oracle.ref_harness.stub_llm import holds
    cond = "(player.is_alive == true or player.role_revealed == true) and (player.can_vote == true or player.night_action_eligible == true) and not (player.night_action_submitted == true and player.is_alive == false)"
    chain = compile_predicate_chain(cond, _fm(T.FAMILY_WEREWOLF))
    names = ["is_alive", "role_revealed", "can_vote", "night_action_eligible", "night_action_submitted"]
    slots = {"is_alive": 0, "can_vote": 1, "night_action_eligible": 2, "night_action_submitted": 3, "role_revealed": 4}
    for bits in itertools.product([False, True], repeat=5):
        raw = np.zeros(cg.record_size, dtype=np.uint8)
        ps = dict(zip(names, bits))
        for n, v in ps.items():
            raw[8 + 4 * slots[n]] = 1 if v else 0                              # player 1 only
        rec = Record(cg, raw)
        got = 0
        for c in chain:
            got |= rec.eval_pred(tuple(c))
        assert bool(got & 1) == holds(cond, ps), (ps, chain)


def test_wait_for_is_validated_and_audience_groups_that_do_not_compile_are_reported():
    import copy
    from game_engine_b200.compiler import load_dsl, load_rules
    cg = compile_game("werewolf-(mafia)", 8)
    assert set(cg.wait_for.values()) <= {"single_player_choice", "all_players_action", "multiple_players_action"} and cg.wait_for
    assert cg.audience_errors == {} and set(cg.audience_chains) == set(cg.dsl["declaration"]["audience_groups"])
    dsl, rules = copy.deepcopy(load_dsl("werewolf-(mafia)")), load_rules("werewolf-(mafia)")
    dsl["declaration"]["audience_groups"]["nicknamed"] = {"selection_criteria": "player.nickname == 'x'"}
    dsl["declaration"]["audience_groups"]["armed"] = {"selection_criteria": "player.selected_target_id >= 1 and player.is_alive == true"}
    cg2 = compile_game("werewolf-(mafia)", 8, dsl=dsl, rules=rules)
    assert list(cg2.audience_errors) == ["nicknamed"] and "armed" in cg2.audience_chains and cg2.table.cmps == [(0, 5, 1)]
    with pytest.raises(DSLCompileError):
        compile_game("werewolf-(mafia)", 8, dsl=dsl, rules=rules, strict_audience=True)
    bad = copy.deepcopy(load_dsl("werewolf-(mafia)"))
    bad["phases"][7]["completion_criteria"]["wait_for"] = "whoever_feels_like_it"
    with pytest.raises(DSLCompileError):
        compile_game("werewolf-(mafia)", 8, dsl=bad, rules=rules)


def test_the_numeric_variant_compiles_to_four_comparison_fields():
    cg = compile_game("two-truths-handicap", 6)
    assert len(cg.table.cmps) == 4 and not cg.audience_errors
    assert len(cg.audience_chains["busy"]) == 2                                  # four clauses: two records
    vote = cg.table.phases[cg.index_of(5)]
    rec = cg.table.preds[vote.actor_pred]
    assert rec[0] & (1 << T.cmp_field_id(T.FAMILY_TTL, cg.table.cmps.index((0, T.CMP_OPS["<"], 2))))
