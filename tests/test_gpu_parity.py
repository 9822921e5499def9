"""GPU parity proper: the CUDA path (through the C ABI) against Oracle B on the same seeded sessions.

Bar: bit-exact canonical records after EVERY step, identical statistics words.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WEREWOLF, TTL, REVOTE, DRAFT = "werewolf-(mafia)", "two-truths-and-a-lie", "werewolf-revote", "werewolf-draft"
CASES = [
    (TTL, 4), (TTL, 3), (TTL, 7), (TTL, 12), (TTL, 32),
    (WEREWOLF, 8), (WEREWOLF, 4), (WEREWOLF, 5), (WEREWOLF, 13), (WEREWOLF, 16), (WEREWOLF, 21), (WEREWOLF, 32),
    (REVOTE, 8), (REVOTE, 32),          # extended game: tie -> re-vote phases (BASELINE config 4)
    (DRAFT, 8), (DRAFT, 19),            # third table (reference game_draft/): generic interpreter kernel, two terminal phases
]


def _batch(cg, n, first, seed, kernel):
    from game_engine_b200.batch import Table, SessionBatch
    t = Table(cg)
    return t, SessionBatch(t, n, first_session_id=first, seed=seed, kernel=kernel)


@pytest.mark.parametrize("kernel", ["tps", "tps_generic", "coop"])
@pytest.mark.parametrize("game,P", CASES)
def test_every_step_bit_exact(games, oracle_for, game, P, kernel):
    cg = games(game, P)
    o = oracle_for(cg)
    n, first, seed = 1000, 12345, 0xC0FFEE      # ragged: 1000 is not a multiple of the 32-session tile
    t, b = _batch(cg, n, first, seed, kernel)
    rec = o.init(n)
    np.testing.assert_array_equal(b.export_state(), rec)
    ost = o.new_stats()
    n_steps = 40 if game == TTL and P <= 4 else 80 if P <= 8 else 140
    if game == TTL:
        n_steps = 2 + 8 * P + 3
    for k in range(n_steps):
        b.step(1)
        o.step(rec, first, seed, 1, ost)
        got = b.export_state()
        if not np.array_equal(got, rec):
            bad = np.nonzero((got != rec).any(axis=1))[0]
            i = int(bad[0])
            raise AssertionError("step %d: %d/%d sessions differ; first sid=%d\n gpu=%s\n cpu=%s"
                                 % (k, len(bad), n, first + i, got[i].tolist(), rec[i].tolist()))
    o.stats_final(rec, ost)
    np.testing.assert_array_equal(b.stats(), ost)
    assert b.counted_steps() == int(ost[0])


@pytest.mark.parametrize("kernel", ["tps", "tps_generic", "coop"])
@pytest.mark.parametrize("game,P,n", [(WEREWOLF, 8, 1 << 16), (WEREWOLF, 32, 1 << 14), (TTL, 4, 1 << 16), (REVOTE, 32, 1 << 13),
                                        (DRAFT, 8, 1 << 15)])
def test_run_to_completion_matches(games, oracle_for, game, P, n, kernel):
    cg = games(game, P)
    o = oracle_for(cg)
    first, seed = 1 << 33, 7          # session ids above 2^32 exercise the high counter word
    t, b = _batch(cg, n, first, seed, kernel)
    rec = o.init(n)
    ost = o.new_stats()
    steps = 400 if game == REVOTE else 256
    b.step(steps)
    o.step(rec, first, seed, steps, ost)
    np.testing.assert_array_equal(b.export_state(), rec)
    o.stats_final(rec, ost)
    gst = b.stats()
    np.testing.assert_array_equal(gst, ost)
    assert gst[1:4].sum() == n
    # every session must have reached the terminal phase within the cap
    kinds = np.array([p.kind for p in cg.table.phases])
    assert (kinds[rec[:, 0]] == 3).all()


@pytest.mark.parametrize("kernel", ["tps", "coop"])
def test_fused_equals_single_steps(games, kernel):
    cg = games(WEREWOLF, 8)
    n = 4096
    t, a = _batch(cg, n, 0, 3, kernel)
    _, b = _batch(cg, n, 0, 3, kernel)
    a.step(37)
    b.run_fused(37)
    np.testing.assert_array_equal(a.export_state(), b.export_state())
    np.testing.assert_array_equal(a.stats(), b.stats())


def test_import_export_round_trip_and_resume(games, oracle_for):
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    n, seed = 777, 99
    rec = o.init(n)
    o.step(rec, 0, seed, 13)
    t, b = _batch(cg, n, 0, seed, "tps")
    b.import_state(rec)
    np.testing.assert_array_equal(b.export_state(), rec)
    # partial export / import windows
    np.testing.assert_array_equal(b.export_state(100, 50), rec[100:150])
    b.step(20)
    o.step(rec, 0, seed, 20)
    np.testing.assert_array_equal(b.export_state(), rec)


def test_run_host_end_to_end(games, oracle_for):
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    n, seed = 5000, 5
    rec = o.init(n)
    t, b = _batch(cg, n, 0, seed, "tps")
    out = np.empty_like(rec)
    st = np.zeros(560, dtype=np.uint64)
    b.run_host(rec, out, 64, st)
    ost = o.new_stats()
    o.step(rec, 0, seed, 64, ost)
    o.stats_final(rec, ost)
    np.testing.assert_array_equal(out, rec)
    np.testing.assert_array_equal(st, ost)


def test_sharding_is_invisible(games):
    """Two shards with contiguous session-id ranges == one batch."""
    cg = games(WEREWOLF, 16)
    n = 3000
    t, whole = _batch(cg, n, 1000, 42, "tps")
    _, lo = _batch(cg, 1700, 1000, 42, "tps")
    _, hi = _batch(cg, n - 1700, 2700, 42, "coop")
    for b in (whole, lo, hi):
        b.step(200)
    np.testing.assert_array_equal(whole.export_state(), np.concatenate([lo.export_state(), hi.export_state()]))
    np.testing.assert_array_equal(whole.stats(), lo.stats() + hi.stats())


def test_bad_import_is_rejected(games):
    from game_engine_b200.capi import GameEngineError
    cg = games(TTL, 4)
    t, b = _batch(cg, 8, 0, 0, "tps")
    rec = b.export_state()
    rec[3, 0] = 200
    with pytest.raises(GameEngineError):
        b.import_state(rec)


@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 32), (TTL, 4)])
def test_compaction_is_invisible(games, oracle_for, game, P):
    """Active-prefix compaction permutes slots on the device; exports, statistics and ids must not notice."""
    cg = games(game, P)
    o = oracle_for(cg)
    n, first, seed = 5000, 77, 31337
    _, on = _batch(cg, n, first, seed, "tps")
    _, every = _batch(cg, n, first, seed, "tps")
    _, off = _batch(cg, n, first, seed, "tps")
    on.set_compaction(4, 2)          # check every 4 steps, compact when >= 1/4 of the prefix is dead
    every.set_compaction(1, 6)       # check every step, compact when >= 1/64 is dead
    off.set_compaction(0)
    rec = o.init(n)
    ost = o.new_stats()
    term = len(cg.phase_ids) - 1
    prev_active = n
    steps = 2 + 8 * P + 4 if game == TTL else 9 * P - 16 + 4
    for k in range(steps):
        for b in (on, every, off):
            b.step(1)
        o.step(rec, first, seed, 1, ost)
        got = every.export_state()
        assert np.array_equal(got, rec), "step %d (compaction every step)" % k
        if k % 5 == 0 or k > steps - 6:
            assert np.array_equal(on.export_state(), rec), "step %d (compaction every 4)" % k
            assert np.array_equal(off.export_state(), rec), "step %d (no compaction)" % k
        act = every.active()
        live = int((rec[:, 0] != term).sum())
        assert live <= act <= prev_active, (k, live, act, prev_active)
        prev_active = act
        assert off.active() == n
    assert every.active() == 0          # every game is over: nothing left to walk
    o.stats_final(rec, ost)
    for b in (on, every, off):
        np.testing.assert_array_equal(b.stats(), ost)
    # partial export windows still address sessions by their original index
    np.testing.assert_array_equal(every.export_state(1234, 321), rec[1234:1555])


def test_import_after_compaction_restores_order(games, oracle_for):
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    n, seed = 3000, 8
    _, b = _batch(cg, n, 0, seed, "tps")
    b.set_compaction(2, 4)
    b.step(30)
    assert b.active() < n
    mid = b.export_state()
    rec = o.init(n)
    o.step(rec, 0, seed, 30)
    np.testing.assert_array_equal(mid, rec)
    b.import_state(mid[100:200], first=100)          # forces the slot order back to identity
    assert b.active() == n
    b.step(40)
    o.step(rec, 0, seed, 40)
    np.testing.assert_array_equal(b.export_state(), rec)
    # switching to the lane-per-player kernel on a compacted batch must also work
    _, c = _batch(cg, n, 0, seed, "tps")
    c.set_compaction(1, 6)
    c.step(25)
    c.set_kernel("coop")
    c.step(45)
    np.testing.assert_array_equal(c.export_state(), rec)


def test_run_host_async_pipeline(games, oracle_for):
    """Several sub-batches driven with run_host_async overlap copies and compute; results equal one oracle run."""
    from game_engine_b200.batch import PinnedBuffer, SessionBatch, Table
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    n, nsub, seed = 8192, 4, 13
    sub = n // nsub
    S = cg.record_size
    tab = Table(cg)
    subs = [SessionBatch(tab, sub, first_session_id=j * sub, seed=seed) for j in range(nsub)]
    pin_in, pin_out, pin_st = PinnedBuffer(n * S), PinnedBuffer(n * S), PinnedBuffer(nsub * 560 * 8)
    rin = pin_in.array.reshape(nsub, sub, S)
    rout = pin_out.array.reshape(nsub, sub, S)
    rst = pin_st.array.view(np.uint64).reshape(nsub, 560)
    rec = o.init(n)
    rin[:] = rec.reshape(nsub, sub, S)
    for j, sb in enumerate(subs):
        sb.set_host_fused(j % 2 == 0)            # fused and launch-per-step sub-batches must agree
        sb.run_host_async(rin[j], rout[j], 56, rst[j])
    for sb in subs:
        sb.sync()
    ost = o.new_stats()
    o.step(rec, 0, seed, 56, ost)
    o.stats_final(rec, ost)
    np.testing.assert_array_equal(rout.reshape(n, S), rec)
    np.testing.assert_array_equal(rst.sum(axis=0), ost)


@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 32)])
def test_audience_masks_on_gpu(games, oracle_for, game, P):
    cg = games(game, P)
    o = oracle_for(cg)
    n, seed = 3000, 4
    _, b = _batch(cg, n, 0, seed, "tps")
    b.set_compaction(2, 4)                    # masks must be reported in original session order after compaction too
    rec = o.init(n)
    for steps in (3, 9, 14, 20):
        b.step(steps)
        o.step(rec, 0, seed, steps)
        want = o.eval_preds(rec, list(cg.audience_preds.values()))
        got = b.eval_preds(list(cg.audience_preds.values()))
        np.testing.assert_array_equal(got, want)
        am = b.audience_masks(100, 50)
        np.testing.assert_array_equal(am["werewolves"], want[100:150, list(cg.audience_preds).index("werewolves")])


@pytest.mark.parametrize("game,P,n", [(REVOTE, 32, 6000), (REVOTE, 8, 5000), (WEREWOLF, 16, 5000)])
def test_phase_regrouping_is_invisible(games, oracle_for, game, P, n):
    """Phase regrouping counting-sorts the slots by phase on the device (and compacts finished games); exports,
    statistics, session ids and audience masks must not notice, whatever the cadence and threshold."""
    cg = games(game, P)
    o = oracle_for(cg)
    first, seed = (1 << 35) + 5, 2027
    _, every = _batch(cg, n, first, seed, "tps")
    _, lazy = _batch(cg, n, first, seed, "tps_generic")
    _, off = _batch(cg, n, first, seed, "tps")
    every.set_regroup(1, 16)         # check every step, regroup as soon as one tile is mixed
    lazy.set_regroup(3, 2)           # check every 3 steps, regroup when >= 1/4 of the tiles are mixed
    off.set_regroup(0)
    rec = o.init(n)
    ost = o.new_stats()
    term = len(cg.phase_ids) - 1
    steps = 9 * P - 16 + (2 * cg.table.max_revotes * (P - 2)) + 4
    prev_active = n
    for k in range(steps):
        for b in (every, lazy, off):
            b.step(1)
        o.step(rec, first, seed, 1, ost)
        assert np.array_equal(every.export_state(), rec), "step %d (regroup every step)" % k
        if k % 7 == 0 or k > steps - 6:
            assert np.array_equal(lazy.export_state(), rec), "step %d (regroup every 3)" % k
            assert np.array_equal(off.export_state(), rec), "step %d (no regroup)" % k
        act = every.active()
        live = int((rec[:, 0] != term).sum())
        assert live <= act <= prev_active, (k, live, act, prev_active)
        prev_active = act
        if k == steps // 3:
            want = o.eval_preds(rec, list(cg.audience_preds.values()))
            np.testing.assert_array_equal(every.eval_preds(list(cg.audience_preds.values())), want)
    assert every.active() == 0
    o.stats_final(rec, ost)
    for b in (every, lazy, off):
        np.testing.assert_array_equal(b.stats(), ost)
    np.testing.assert_array_equal(every.export_state(1000, 777), rec[1000:1777])
    # a regrouped batch can be re-initialised, imported into, and switched to the lane-per-player kernel
    every.reset(first_session_id=first + n)
    rec2 = o.init(n)
    every.step(21)
    o.step(rec2, first + n, seed, 21)
    np.testing.assert_array_equal(every.export_state(), rec2)
    every.set_kernel("coop")
    every.step(9)
    o.step(rec2, first + n, seed, 9)
    np.testing.assert_array_equal(every.export_state(), rec2)
    every.set_kernel("tps")
    every.step(30)
    o.step(rec2, first + n, seed, 30)
    np.testing.assert_array_equal(every.export_state(), rec2)


@pytest.mark.parametrize("game,P,every,regroup", [(WEREWOLF, 8, 4, False), (TTL, 4, 8, False), (REVOTE, 8, 4, True), (DRAFT, 6, 2, False)])
def test_autoreset_runs_epochs_on_the_device(games, oracle_for, game, P, every, regroup):
    """Continuous simulation: when the periodic check finds every game over, the batch restarts on the device with
    session ids first + epoch * stride + i.  The oracle replays the same rule on the host."""
    cg = games(game, P)
    o = oracle_for(cg)
    n, first, seed, stride = 1500, 10_000, 77, 1 << 20
    _, b = _batch(cg, n, first, seed, "tps")
    if regroup:
        b.set_regroup(every, 3)
    else:
        b.set_regroup(0)
        b.set_compaction(every, 2)
    b.set_autoreset(stride)
    kinds = np.array([p.kind for p in cg.table.phases])
    rec = o.init(n)
    ost = o.new_stats()
    harvested = o.new_stats()
    epoch = 0
    steps = 3 * (9 * P - 16 + 2 * cg.table.max_revotes * (P - 2) if game != TTL else 2 + 8 * P) + 40
    for k in range(steps):
        b.step(1)
        o.step(rec, first + epoch * stride, seed, 1, ost)
        if (k + 1) % every == 0 and (kinds[rec[:, 0]] == 3).all():       # the device-side check after this launch
            fin = o.new_stats()
            o.stats_final(rec, fin)                                       # (stats_final SETS the final-state histograms)
            harvested += fin
            rec = o.init(n)
            epoch += 1
        if k % 3 == 0 or epoch:
            assert np.array_equal(b.export_state(), rec), "step %d epoch %d" % (k, epoch)
    assert epoch >= 2 and b.epochs() == epoch
    o.stats_final(rec, ost)
    ost[1:260] += harvested[1:260]
    ost[292:548] += harvested[292:548]
    np.testing.assert_array_equal(b.stats(), ost)
    # a host reset goes back to epoch 0
    b.reset(first_session_id=5)
    assert b.epochs() == 0
    rec = o.init(n)
    b.step(20)
    o.step(rec, 5, seed, 20)
    np.testing.assert_array_equal(b.export_state(), rec)


@pytest.mark.parametrize("game,P,kernel", [(WEREWOLF, 8, "tps"), (WEREWOLF, 8, "tps_generic"), (WEREWOLF, 32, "tps"), (TTL, 4, "tps"),
                                           (DRAFT, 6, "tps"), (WEREWOLF, 16, "tps")])
def test_ring_launch_equals_separate_launches(games, oracle_for, game, P, kernel):
    """ge_step_ring (one launch per pass over a ring of batches, compaction checks batched) == the oracle, with the
    ring's batches at different depths of their games and of different sizes."""
    from game_engine_b200.batch import SessionBatch, Table, step_ring
    from game_engine_b200.capi import GameEngineError
    cg = games(game, P)
    o = oracle_for(cg)
    t = Table(cg)
    sizes = [3000, 777, 4096, 33, 1500]
    firsts = [0, 10000, 20000, 30000, 40000]
    seed = 77
    ring = [SessionBatch(t, n, first_session_id=f, seed=seed, kernel=kernel) for n, f in zip(sizes, firsts)]
    st = ring[0].state_device_ptr() and 0
    import torch
    stream = torch.cuda.Stream()
    for i, b in enumerate(ring):
        b.set_stream(stream.cuda_stream)
        b.set_compaction(3, 2) if cg.family == 1 else None
        b.step(4 * i)                                  # staggered ages
    total = 0
    for rounds in (1, 7, 30, 25):
        step_ring(ring, rounds)
        total += rounds
        for i, b in enumerate(ring):
            rec = o.init(sizes[i])
            ost = o.new_stats()
            o.step(rec, firsts[i], seed, 4 * i + total, ost)
            o.stats_final(rec, ost)
            np.testing.assert_array_equal(b.export_state(), rec)
            np.testing.assert_array_equal(b.stats(), ost)
    # a sub-list is a ring too; mismatched seeds are refused
    step_ring(ring[1:3], 2)
    other = SessionBatch(t, 64, first_session_id=0, seed=seed + 1, kernel=kernel)
    other.set_stream(stream.cuda_stream)
    with pytest.raises(GameEngineError):
        step_ring([ring[0], other], 1)


@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 16), (REVOTE, 32), (TTL, 4)])
def test_programmatic_dependent_launches_change_nothing(games, oracle_for, game, P):
    """GE_OPT_PDL: back-to-back step launches of one stream may overlap the previous kernel's tail; they wait on the device
    before reading its results.  Same records, same statistics — single-batch launches, compaction / regroup launches in
    between, and the ring launch."""
    import torch
    from game_engine_b200.batch import SessionBatch, Table, step_ring
    cg = games(game, P)
    o = oracle_for(cg)
    n, first, seed = 50000, 99, 4
    t, b = _batch(cg, n, first, seed, "tps")
    b.set_option("pdl", 1)
    if game != TTL:
        b.set_compaction(2, 4) if game != REVOTE else b.set_regroup(2, 4)
    rec = o.init(n)
    ost = o.new_stats()
    for chunk in (1, 7, 30, 60):
        b.step(chunk)
        o.step(rec, first, seed, chunk, ost)
        np.testing.assert_array_equal(b.export_state(), rec)
    o.stats_final(rec, ost)
    np.testing.assert_array_equal(b.stats(), ost)
    if game == REVOTE:
        return
    stream = torch.cuda.Stream(device=0)
    ring = [SessionBatch(t, 20000, first_session_id=f, seed=seed) for f in (0, 1 << 34)]
    for r in ring:
        r.set_stream(stream.cuda_stream)
        r.set_option("pdl", 1)
    step_ring(ring, 45)
    for r, f in zip(ring, (0, 1 << 34)):
        want = o.init(20000)
        o.step(want, f, seed, 45)
        np.testing.assert_array_equal(r.export_state(), want)


def test_learned_compaction_schedule_and_progress_hint(games, oracle_for):
    """Compaction checks that did not fire in an earlier epoch are not launched again (the host learns the schedule from a
    log the scan kernel keeps in mapped memory); results never depend on when compaction happens.  The same mapping
    carries the progress hint: an upper bound of the live prefix that only decreases, 0 once every game is over."""
    cg = games(WEREWOLF, 8)
    o = oracle_for(cg)
    n, seed, cap = 1 << 15, 17, 56
    t, b = _batch(cg, n, 0, seed, "tps")
    b.set_compaction(5, 2)
    launches = []
    for epoch in range(1, 8):
        first = epoch * n
        b.reset(first_session_id=first)
        l0 = b.launch_count()
        hints = []
        for _ in range(cap // 4):
            b.step(4)
            b.sync()
            hints.append(b.active_hint())
        launches.append(b.launch_count() - l0)
        rec = o.init(n)
        o.step(rec, first, seed, 4 * (cap // 4))
        np.testing.assert_array_equal(b.export_state(), rec)
        assert all(x >= y for x, y in zip(hints, hints[1:])) and hints[0] <= n and hints[-1] < n // 4, hints
    # the first epoch runs all 11 checks (two launches each), later ones only those that fired; every 8th epoch of the
    # batch observes all of them again (refresh)
    assert launches[0] >= 56 + 2 * 11 and max(launches[1:-1]) <= 56 + 2 * 6 + 2 and launches[-1] == launches[0], launches
    assert b.active() < n // 4
