"""Host adapter on many sessions: the dict state materialised from Oracle B's records (what the drop-in nodes and the
trace exporter emit) packs back to exactly the same record at every step — for the shipped games, the re-vote variant
(whose re-vote counter and tie flag travel in two declared per-player fields, rules `session_fields`) and the aliased
draft schema."""
import numpy as np
import pytest

from helpers import oracle_b_records

CASES = [("werewolf-(mafia)", 5), ("werewolf-(mafia)", 8), ("werewolf-(mafia)", 13), ("werewolf-draft", 7),
         ("werewolf-revote", 8), ("werewolf-revote", 6), ("werewolf-revote", 32),
         ("two-truths-and-a-lie", 3), ("two-truths-and-a-lie", 4), ("two-truths-and-a-lie", 9)]


@pytest.mark.parametrize("game,P", CASES)
def test_record_dict_record_round_trip(games, oracle_for, game, P):
    from game_engine_b200.adapter import SessionCodec
    cg = games(game, P)
    o = oracle_for(cg)
    codec = SessionCodec(cg)
    cap = 2 + 8 * P if cg.family == 2 else 9 * P - 16 + 2 * cg.table.max_revotes * (P - 2)
    checked = ties = 0
    for sid in range(12 if P < 32 else 4):
        recs = oracle_b_records(o, 1000 + sid, 31, cap)
        state = codec.initial_state()
        assert np.array_equal(codec.record_from_state(state), recs[0])
        for k in range(1, len(recs)):
            state.update(codec.step_update(state, recs[k - 1], recs[k], now_ms=0, now_iso=""))
            back = codec.record_from_state(state)
            assert np.array_equal(back, recs[k]), "sid %d step %d\n got=%s\nwant=%s" % (sid, k, back.tolist(), recs[k].tolist())
            checked += 1
            ties += cg.family == 1 and bool(recs[k][7] & 0x80)
        # dict-level invariants of the reference's schema
        ps = state["player_states"]
        assert sorted(ps, key=int) == [str(i + 1) for i in range(P)]
        assert len(state["phase_history"]) == int(recs[-1][2]) | (int(recs[-1][3]) << 8)
        assert all(set(e) == {"phase_id", "phase_name", "timestamp"} for e in state["phase_history"])
        for pid, pa in state["playerActions"].items():
            ids = sorted(int(a["id"]) for a in pa["actions"].values())
            assert ids == list(range(1, len(ids) + 1)) and pa["name"] == ps[pid]["name"]
    assert checked > (12 if P < 32 else 4) * 10
    if cg.table.max_revotes:
        assert ties > 0          # the tie-pending state really occurred and survived the dict


def test_a_table_with_revotes_needs_session_fields(games):
    """Without a place in the dict for the re-vote counter the drop-in nodes would silently take the no-tie branch:
    the compiler refuses such a game instead."""
    from game_engine_b200.compiler import DSLCompileError, compile_game, load_dsl, load_rules
    rules = load_rules("werewolf-revote")
    rules.pop("session_fields")
    with pytest.raises(DSLCompileError):
        compile_game("werewolf-revote", 8, dsl=load_dsl("werewolf-revote"), rules=rules)
