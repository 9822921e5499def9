"""Host adapter on many sessions: the dict state materialised from Oracle B's records (what the drop-in nodes and the
trace exporter emit) packs back to exactly the same record at every step — for the shipped games, the re-vote variant
(except its re-vote counter, which the reference's dict has no field for) and the aliased draft schema."""
import numpy as np
import pytest

from helpers import oracle_b_records

CASES = [("werewolf-(mafia)", 5), ("werewolf-(mafia)", 8), ("werewolf-(mafia)", 13), ("werewolf-draft", 7),
         ("two-truths-and-a-lie", 3), ("two-truths-and-a-lie", 4), ("two-truths-and-a-lie", 9)]


@pytest.mark.parametrize("game,P", CASES)
def test_record_dict_record_round_trip(games, oracle_for, game, P):
    from game_engine_b200.adapter import SessionCodec
    cg = games(game, P)
    o = oracle_for(cg)
    codec = SessionCodec(cg)
    cap = 2 + 8 * P if cg.family == 2 else 9 * P - 16
    checked = 0
    for sid in range(12):
        recs = oracle_b_records(o, 1000 + sid, 31, cap)
        state = codec.initial_state()
        assert np.array_equal(codec.record_from_state(state), recs[0])
        for k in range(1, len(recs)):
            state.update(codec.step_update(state, recs[k - 1], recs[k], now_ms=0, now_iso=""))
            back = codec.record_from_state(state)
            assert np.array_equal(back, recs[k]), "sid %d step %d\n got=%s\nwant=%s" % (sid, k, back.tolist(), recs[k].tolist())
            checked += 1
        # dict-level invariants of the reference's schema
        ps = state["player_states"]
        assert sorted(ps, key=int) == [str(i + 1) for i in range(P)]
        assert len(state["phase_history"]) == int(recs[-1][2]) | (int(recs[-1][3]) << 8)
        assert all(set(e) == {"phase_id", "phase_name", "timestamp"} for e in state["phase_history"])
        for pid, pa in state["playerActions"].items():
            ids = sorted(int(a["id"]) for a in pa["actions"].values())
            assert ids == list(range(1, len(ids) + 1)) and pa["name"] == ps[pid]["name"]
    assert checked > 12 * 10
