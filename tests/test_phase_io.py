"""ge_table_phase_io / ge_table_phase_io_packed: the bytes a step that starts in each phase must move (host-side table
analysis, no GPU needed) — the `necessary_bytes_per_step` of bench.py's roofline."""
import numpy as np
import pytest

WEREWOLF, TTL, REVOTE, DRAFT = "werewolf-(mafia)", "two-truths-and-a-lie", "werewolf-revote", "werewolf-draft"


@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 5), (WEREWOLF, 16), (WEREWOLF, 11), (DRAFT, 8), (REVOTE, 16)])
def test_packed_store_moves_no_more_than_the_canonical_one(games, game, P):
    from game_engine_b200.batch import Table
    cg = games(game, P)
    t = Table(cg)
    canon, packed = t.phase_io(), t.phase_io(packed=True)
    S_pk = 32 if P <= 8 else 48
    kinds = [p.kind for p in cg.table.phases]
    assert len(canon) == len(packed) == len(kinds)
    saw_less = False
    for k, (c, p) in zip(kinds, zip(canon, packed)):
        if k == 3:                                   # terminal: nothing moves
            assert c == (0, 0) and p == (0, 0)
            continue
        assert p[0] >= 16 and p[1] >= 16             # column D0 both ways on every step
        assert p[0] % 16 == 0 and p[0] <= S_pk and p[1] <= S_pk
        assert p[0] <= c[0] and p[1] <= c[1] + 15    # (a directly stored target byte vs a whole column)
        if k in (0, 1) and c == (16, 16):            # header-only UI / timer phases stay header-only
            assert p == (16, 16)
        saw_less |= p[0] + p[1] < c[0] + c[1]
    assert saw_less
    # weighted by a uniform visit histogram the packed store needs fewer bytes per step
    stats = np.zeros(560, dtype=np.uint64)
    stats[260:260 + len(kinds)] = 1
    assert t.necessary_bytes_per_step(stats, packed=True) < t.necessary_bytes_per_step(stats)


def test_packed_io_is_refused_for_tables_the_packed_store_does_not_cover(games):
    from game_engine_b200.batch import Table
    from game_engine_b200.capi import GameEngineError
    for game, P in ((WEREWOLF, 17), (WEREWOLF, 32), (TTL, 4)):
        with pytest.raises(GameEngineError):
            Table(games(game, P)).phase_io(packed=True)
