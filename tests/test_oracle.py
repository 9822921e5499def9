"""Oracle B (scalar C restatement) pinned against known answers and domain properties."""
import numpy as np
import pytest

from game_engine_b200 import table as T

WEREWOLF, TTL = "werewolf-(mafia)", "two-truths-and-a-lie"

# Random123 kat_vectors, philox4x32 10 rounds: (ctr, key, expected)
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_philox_known_answers(games, oracle_for):
    o = oracle_for(games(TTL, 4))
    for ctr, key, want in PHILOX_KAT:
        assert o.philox(key, ctr).tolist() == want


def _fields_w(rec):
    u32 = lambda off: rec[:, off:off + 4].copy().view(np.uint32)[:, 0]
    return dict(phase=rec[:, 0], prev=rec[:, 1], step=rec[:, 2:4].copy().view(np.uint16)[:, 0], winner=rec[:, 4],
                kill=rec[:, 5], protect=rec[:, 6], alive=u32(8), can_vote=u32(12), eligible=u32(16), submitted=u32(20),
                revealed=u32(24), investigated=u32(28), wolf=u32(32), secret=u32(36), lo=u32(40), hi=u32(44))


def _popc(a):
    return np.array([bin(int(x)).count("1") for x in a])


@pytest.mark.parametrize("game,P", [(WEREWOLF, 4), (WEREWOLF, 8), (WEREWOLF, 13), (WEREWOLF, 16), (WEREWOLF, 32),
                                    ("werewolf-revote", 8), ("werewolf-revote", 16),
                                    ("werewolf-draft", 6), ("werewolf-draft", 8), ("werewolf-draft", 32)])
def test_werewolf_properties(games, oracle_for, game, P):
    cg = games(game, P)
    o = oracle_for(cg)
    n, seed = 2000, 11
    rec = o.init(n)
    st = o.new_stats()
    is_term = np.array([ph.kind == T.KIND_TERMINAL for ph in cg.table.phases])
    edges = {i: {b.next for b in ph.branches} for i, ph in enumerate(cg.table.phases)}
    prev_alive = _popc(_fields_w(rec)["alive"])
    for k in range(420 if game != WEREWOLF else 300):
        before = rec.copy()
        o.step(rec, 0, seed, 1, st)
        f, fb = _fields_w(rec), _fields_w(before)
        moved = (f["step"] != fb["step"])
        # history follows YAML edges only; the first step is the phase-0 double visit
        for a, b in set(zip(fb["phase"][moved].tolist(), f["phase"][moved].tolist())):
            assert b in edges[a] or (a == 0 and b == 0), (a, b)
        # terminal sessions are frozen
        np.testing.assert_array_equal(rec[~moved], before[~moved])
        assert is_term[fb["phase"][~moved]].all()
        # alive count never increases, at most one death per step
        alive = _popc(f["alive"])
        assert ((prev_alive - alive) >= 0).all() and ((prev_alive - alive) <= 1).all()
        prev_alive = alive
        if k == 2:      # roles assigned on entry to phase 1 (second step)
            W = cg.table.n_wolves
            assert (_popc(f["wolf"]) == W).all() and (_popc(f["secret"]) == W + 2).all()
            assert (_popc(f["lo"] & f["hi"]) == 1).all() and (_popc(~f["lo"] & f["hi"]) == 1).all()
        # dead players cannot vote or act
        assert ((f["can_vote"] & ~f["alive"]) == 0).all() and ((f["eligible"] & ~f["alive"]) == 0).all()
    f = _fields_w(rec)
    assert is_term[f["phase"]].all(), "every game must end"
    wolves_alive, vill_alive = _popc(f["wolf"] & f["alive"]), _popc(~f["wolf"] & f["alive"])
    assert ((f["winner"] == 1) == (wolves_alive == 0)).all()            # terminal iff win predicate
    assert ((f["winner"] == 2) == ((wolves_alive > 0) & (wolves_alive >= vill_alive))).all()
    o.stats_final(rec, st)
    assert st[1] == 0 and st[2] + st[3] == n and st[4:260].sum() == n and st[260:292].sum() == st[0]
    assert st[0] == f["step"].astype(np.int64).sum()


@pytest.mark.parametrize("P", [3, 4, 7, 12, 32])
def test_ttl_properties(games, oracle_for, P):
    cg = games(TTL, P)
    o = oracle_for(cg)
    n = 500
    rec = o.init(n)
    st = o.new_stats()
    o.step(rec, 5, 3, 2 + 8 * P + 5, st)
    assert (rec[:, 0] == len(cg.phase_ids) - 1).all()
    steps = rec[:, 2:4].copy().view(np.uint16)[:, 0]
    assert (steps == 2 + 8 * P).all()               # 0,0 then 8 phases per speaker, last edge enters 99
    pl = rec[:, 8:8 + 4 * P].reshape(n, P, 4)
    assert (pl[:, :, 1] == 1).all()                 # everybody spoke exactly once
    # each round hands out exactly (P-1) points: +1 per correct voter, +1 to the speaker per wrong voter
    assert (pl[:, :, 0].astype(int).sum(axis=1) == P * (P - 1)).all()
    winner = rec[:, 6]
    best = pl[:, :, 0].argmax(axis=1) + 1           # argmax returns the lowest index among ties
    np.testing.assert_array_equal(winner, best)


def test_step_batching_and_threads_do_not_change_results(games, oracle_for):
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    a, b, c = o.init(3000), o.init(3000), o.init(3000)
    sa, sb, sc = o.new_stats(), o.new_stats(), o.new_stats()
    for _ in range(70):
        o.step(a, 100, 9, 1, sa, threads=1)
    o.step(b, 100, 9, 70, sb, threads=4)
    o.step(c[:1234], 100, 9, 70, sc, threads=2)
    o.step(c[1234:], 100 + 1234, 9, 70, sc, threads=2)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, c)
    np.testing.assert_array_equal(sa, sb)
    np.testing.assert_array_equal(sa, sc)


def test_seed_and_session_id_matter(games, oracle_for):
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    a, b, c = o.init(64), o.init(64), o.init(64)
    o.step(a, 0, 1, 30)
    o.step(b, 0, 2, 30)
    o.step(c, 64, 1, 30)
    assert not np.array_equal(a, b) and not np.array_equal(a, c)


def test_peek_choices_are_what_the_step_records(games, oracle_for):
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    rec = o.init(50)
    for k in range(40):
        before = rec.copy()
        o.step(rec, 7, 21, 1)
        for i in range(50):
            ph = cg.table.phases[before[i, 0]]
            if ph.kind != T.KIND_ACTION or before[i, 2] == 0:
                continue
            ch = o.peek_choices(before[i], 7 + i, 21)
            tg = rec[i, 48:56]
            for p in range(8):
                if ch[p]:
                    assert tg[p] == ch[p]
