"""Human seats (SPEC.md D3h): a person's input replaces the Philox draw of their seat, and a session whose acting
human seats are not all answered stays in its phase while its history still grows (reference
agent/game_agent_v2.py:1144-1170, 1206-1215; agent/prompt/bot_behavior_system_prompt.txt:3, 58-61).

CPU part: properties of Oracle B's restatement.  GPU part: the thread-per-session kernels against it, bit for bit,
with random seats and random (valid, invalid and missing) inputs on every step."""
import numpy as np
import pytest

WEREWOLF, TTL, REVOTE, DRAFT = "werewolf-(mafia)", "two-truths-and-a-lie", "werewolf-revote", "werewolf-draft"


def random_inputs(rng, o, cg, rec, masks, first, seed):
    """Per step: for every session a row of inputs; a third missing, a third what a bot would have chosen in that
    seat (always valid), a third random bytes (often invalid)."""
    n, P = rec.shape[0], cg.n_players
    stride = ((P + 7) // 8) * 8
    ch = np.full((n, stride), 0xFF, dtype=np.uint8)
    kind = rng.integers(0, 3, size=n)
    for i in range(n):
        if masks[i] == 0 or kind[i] == 0:
            continue
        if kind[i] == 1:
            ch[i, :P] = o.peek_choices(rec[i], first + i, seed)          # 0 for seats that do not act
            ch[i, :P][ch[i, :P] == 0] = 0xFF
        else:
            ch[i, :P] = rng.integers(0, P + 3, size=P)
    return ch


@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (TTL, 4)])
def test_no_human_seats_is_the_all_bot_game(games, oracle_for, game, P):
    cg = games(game, P)
    o = oracle_for(cg)
    a, b = o.init(500), o.init(500)
    o.step(a, 7, 3, 40)
    for _ in range(40):
        o.step_humans(b, 7, 3, np.zeros(500, dtype=np.uint32), None)
    np.testing.assert_array_equal(a, b)


def test_a_waiting_session_only_grows_its_history(games, oracle_for):
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    rec = o.init(64)
    masks = np.full(64, 0xFF, dtype=np.uint32)            # every seat human, nobody ever answers
    st = o.new_stats()
    for k in range(10):
        before = rec.copy()
        o.step_humans(rec, 0, 1, masks, None, stats=st)
        if k >= 3:                                        # phase index 2: the wolves' vote waits for the wolves
            assert (rec[:, 0] == 2).all() and (rec[:, 1] == 2).all()
            np.testing.assert_array_equal(rec[:, 4:], before[:, 4:])
            np.testing.assert_array_equal(rec[:, 2].astype(int) | (rec[:, 3].astype(int) << 8), k + 1)
    assert int(st[0]) == 64 * 10 and int(st[260 + 2]) == 64 * 8    # a stay is a counted step and a visit of the phase


def test_a_person_playing_the_bots_choices_changes_nothing(games, oracle_for):
    """Feeding every human seat exactly what its bot would have drawn reproduces the all-bot game."""
    cg = games(REVOTE, 8)
    o = oracle_for(cg)
    n, first, seed = 300, 50, 9
    a, b = o.init(n), o.init(n)
    masks = np.random.default_rng(0).integers(0, 256, size=n).astype(np.uint32)
    for _ in range(60):
        ch = np.full((n, 8), 0xFF, dtype=np.uint8)
        for i in range(n):
            c = o.peek_choices(b[i], first + i, seed)
            ch[i] = np.where(c == 0, 0xFF, c)
        o.step(a, first, seed, 1)
        o.step_humans(b, first, seed, masks, ch)
        # a seat with no legal target (choice 0) never waits, so the two games stay in step
        np.testing.assert_array_equal(a, b)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["tps", "tps_generic"])
@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 5), (WEREWOLF, 16), (REVOTE, 32), (DRAFT, 7), (TTL, 4), (TTL, 9)])
def test_kernels_follow_the_oracle_with_people_at_the_table(games, oracle_for, game, P, kernel):
    from game_engine_b200.batch import SessionBatch, Table
    cg = games(game, P)
    o = oracle_for(cg)
    n, first, seed = 2500, 1000, 13
    rng = np.random.default_rng(P + len(game))
    masks = np.where(rng.random(n) < 0.3, 0, rng.integers(0, 1 << min(P, 31), size=n)).astype(np.uint32)
    masks[:50] = 1                                      # the reference's room: player 1 is the person
    b = SessionBatch(Table(cg), n, first_session_id=first, seed=seed, kernel=kernel)
    b.set_compaction(3, 2) if cg.family == 1 else None
    b.set_human_seats(masks)
    assert b.human_stride == ((P + 7) // 8) * 8
    rec = o.init(n)
    ost = o.new_stats()
    waited = 0
    for k in range(70):
        ch = random_inputs(rng, o, cg, rec, masks, first, seed)
        before = rec.copy()
        o.step_humans(rec, first, seed, masks, ch, stats=ost)
        waited += int(((rec[:, 0] == before[:, 0]) & (rec[:, 2] != before[:, 2]) & (before[:, 2] > 0)).sum())
        b.set_human_choices(ch)
        b.step(1)
        got = b.export_state()
        assert np.array_equal(got, rec), "step %d: first differing session %d" % (k, np.flatnonzero((got != rec).any(axis=1))[0])
    o.stats_final(rec, ost)
    np.testing.assert_array_equal(b.stats(), ost)
    assert waited > 100                                  # people did keep sessions waiting
    # inputs are consumed by the step they were given for: the next step sees "has not acted"
    b.step(1)
    o.step_humans(rec, first, seed, masks, None)
    np.testing.assert_array_equal(b.export_state(), rec)
    # back to all bots
    b.set_human_seats(None)
    b.step(5)
    o.step(rec, first, seed, 5)
    np.testing.assert_array_equal(b.export_state(), rec)


@pytest.mark.gpu
def test_human_seats_need_the_single_step_thread_per_session_path(games):
    from game_engine_b200.batch import SessionBatch, Table
    from game_engine_b200.capi import GameEngineError
    cg = games(WEREWOLF, 8)
    b = SessionBatch(Table(cg), 64, kernel="tps")
    b.set_human_seats(np.ones(64, dtype=np.uint32))
    with pytest.raises(GameEngineError):
        b.run_fused(4)
    with pytest.raises(GameEngineError):
        b.set_kernel("coop")
    with pytest.raises(GameEngineError):
        b.set_human_seats(np.full(64, 1 << 20, dtype=np.uint32))     # seat above the player count
