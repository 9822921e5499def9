"""The PACKED session store (GE_OPT_STORE_PACKED: werewolf tables up to 8 / 16 players keep a record as the 32 / 48 bytes of the
dense wire format in two / three 16-byte columns) against Oracle B, through the C ABI: every byte of every canonical record after every
step, on every path that touches the store — step kernels (specialised and interpreter), fused launches, the ring
launch, compaction, auto-reset, both wire formats of import / export / the host-buffer call, the device-side record
validator — and the transparent conversions for callers the packed layout does not serve (lane-per-player kernels,
human seats, phase regrouping, audience masks)."""
import numpy as np
import pytest

import game_engine_b200.table as T
from test_fuzz_tables import SEEDS, _Blob, random_table
from test_wire import _played, mutate

pytestmark = pytest.mark.gpu

WEREWOLF, TTL, REVOTE, DRAFT = "werewolf-(mafia)", "two-truths-and-a-lie", "werewolf-revote", "werewolf-draft"


def _batch(cg, n, first, seed, kernel="tps", packed=True, table=None):
    from game_engine_b200.batch import SessionBatch, Table
    t = table or Table(cg)
    b = SessionBatch(t, n, first_session_id=first, seed=seed, kernel=kernel)
    b.set_option("store_packed", 1 if packed else 0)          # (on by default for the tables it covers)
    return t, b


def _diff(got, rec, first, k):
    bad = np.nonzero((got != rec).any(axis=1))[0]
    i = int(bad[0])
    return "step %d: %d/%d sessions differ; first sid=%d\n gpu=%s\n cpu=%s" % (k, len(bad), len(rec), first + i, got[i].tolist(), rec[i].tolist())


@pytest.mark.parametrize("kernel", ["tps", "tps_generic"])
@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 4), (WEREWOLF, 5), (WEREWOLF, 7), (DRAFT, 8), (DRAFT, 6), (REVOTE, 8),
                                    (WEREWOLF, 16), (WEREWOLF, 9), (WEREWOLF, 13), (DRAFT, 12), (REVOTE, 16)])
def test_every_step_bit_exact(games, oracle_for, game, P, kernel):
    cg = games(game, P)
    o = oracle_for(cg)
    n, first, seed = 1000, 12345, 0xC0FFEE
    t, b = _batch(cg, n, first, seed, kernel)
    # 32 / 48 bytes per session in HBM — except the re-vote table, whose phase regrouping keeps the canonical columns
    assert b.state_device_bytes() == ((n + 31) // 32) * 32 * ((32 if P <= 8 else 48) if game != REVOTE else cg.record_size)
    rec = o.init(n)
    np.testing.assert_array_equal(b.export_state(), rec)
    ost = o.new_stats()
    for k in range(90 if P <= 8 else 150):
        b.step(1)
        o.step(rec, first, seed, 1, ost)
        got = b.export_state()
        assert np.array_equal(got, rec), _diff(got, rec, first, k)
    o.stats_final(rec, ost)
    np.testing.assert_array_equal(b.stats(), ost)
    assert b.counted_steps() == int(ost[0])


@pytest.mark.parametrize("P", [8, 16])
@pytest.mark.parametrize("kernel", ["tps", "tps_generic"])
def test_run_to_completion_with_compaction(games, oracle_for, kernel, P):
    cg = games(WEREWOLF, P)
    o = oracle_for(cg)
    n, first, seed = 1 << 16, 1 << 33, 7
    t, b = _batch(cg, n, first, seed, kernel)
    b.set_compaction(3, 3)
    rec = o.init(n)
    ost = o.new_stats()
    for chunk in (10, 17, 29, 50, 200):                   # exports of a compacted (permuted) packed store in between
        b.step(chunk)
        o.step(rec, first, seed, chunk, ost)
        np.testing.assert_array_equal(b.export_state(), rec)
        np.testing.assert_array_equal(b.export_state(1000, 333), rec[1000:1333])
    assert b.active() < n
    o.stats_final(rec, ost)
    gst = b.stats()
    np.testing.assert_array_equal(gst, ost)
    assert gst[1:4].sum() == n


@pytest.mark.parametrize("P", [8, 16])
def test_fused_equals_single_steps_and_the_canonical_store(games, P):
    cg = games(WEREWOLF, P)
    n = 4096
    _, a = _batch(cg, n, 0, 3)
    _, b = _batch(cg, n, 0, 3)
    _, c = _batch(cg, n, 0, 3, packed=False)
    assert c.state_device_bytes() == n * cg.record_size and a.state_device_bytes() == n * (32 if P == 8 else 48)
    a.step(67)
    b.run_fused(67)
    c.step(67)
    np.testing.assert_array_equal(a.export_state(), c.export_state())
    np.testing.assert_array_equal(b.export_state(), c.export_state())
    np.testing.assert_array_equal(a.stats(), c.stats())
    np.testing.assert_array_equal(b.stats(), c.stats())


@pytest.mark.parametrize("P", [8, 14])
@pytest.mark.parametrize("wire_fmt", ["canonical", "dense"])
def test_import_export_windows_and_host_buffer_call(games, oracle_for, wire_fmt, P):
    from game_engine_b200 import wire
    cg = games(WEREWOLF, P)
    o = oracle_for(cg)
    n, seed = 5000, 99
    conv = (lambda r: wire.to_dense(cg, r)) if wire_fmt == "dense" else (lambda r: r)
    rec = o.init(n)
    o.step(rec, 0, seed, 13)
    t, b = _batch(cg, n, 0, seed)
    b.set_wire(wire_fmt)
    b.import_state(conv(rec))
    np.testing.assert_array_equal(b.export_state(), conv(rec))
    np.testing.assert_array_equal(b.export_state(100, 50), conv(rec)[100:150])
    b.set_compaction(2, 4)
    b.step(30)
    o.step(rec, 0, seed, 30)
    np.testing.assert_array_equal(b.export_state(), conv(rec))
    # a partial import into a compacted packed store (slot order is restored first)
    fresh = o.init(64)
    b.import_state(conv(fresh), first=200)
    rec[200:264] = fresh
    np.testing.assert_array_equal(b.export_state(), conv(rec))
    b.step(9)
    o.step(rec, 0, seed, 9)
    np.testing.assert_array_equal(b.export_state(), conv(rec))
    # the host-buffer call: records in, fused steps, records + statistics out
    start = o.init(n)
    out = np.empty_like(conv(start))
    st = np.zeros(560, dtype=np.uint64)
    b.reset()
    b.clear_stats()
    b.run_host(conv(start), out, 40, st)
    ost = o.new_stats()
    o.step(start, 0, seed, 40, ost)
    o.stats_final(start, ost)
    np.testing.assert_array_equal(out, conv(start))
    np.testing.assert_array_equal(st, ost)


@pytest.mark.parametrize("P", [8, 16])
def test_ring_launch(games, oracle_for, P):
    import torch
    from game_engine_b200.batch import step_ring
    cg = games(WEREWOLF, P)
    o = oracle_for(cg)
    stream = torch.cuda.Stream(device=0)
    firsts = (0, 7000, 1 << 35)
    ring, tab = [], None
    for f in firsts:
        tab, b = _batch(cg, 3000, f, 11, table=tab)
        b.set_stream(stream.cuda_stream)
        ring.append(b)
    for rounds in (5, 20, 40):
        step_ring(ring, rounds)
    for b, f in zip(ring, firsts):
        rec = o.init(3000)
        o.step(rec, f, 11, 65)
        np.testing.assert_array_equal(b.export_state(), rec)
    # mixed store formats in one ring are refused
    from game_engine_b200.capi import GameEngineError
    ring[1].set_option("store_packed", 0)
    with pytest.raises(GameEngineError):
        step_ring(ring, 1)


@pytest.mark.parametrize("P", [8, 11])
def test_auto_reset(games, oracle_for, P):
    cg = games(WEREWOLF, P)
    o = oracle_for(cg)
    n, first, seed = 2048, 50, 5
    t, b = _batch(cg, n, first, seed)
    b.set_autoreset(1 << 20)
    total = 0
    for _ in range(30 if P == 8 else 60):
        b.step(8)
        total += 8
    epochs = b.epochs()
    assert epochs >= 1
    # replay: epoch e steps sessions first + e * stride + i until every game is over at a compaction check
    got = b.export_state()
    _, c = _batch(cg, n, first, seed, packed=False)
    c.set_autoreset(1 << 20)
    c.step(total)
    assert c.epochs() == epochs
    np.testing.assert_array_equal(got, c.export_state())
    np.testing.assert_array_equal(b.stats(), c.stats())


@pytest.mark.parametrize("P", [8, 16])
def test_switching_to_callers_the_packed_store_does_not_serve(games, oracle_for, P):
    """lane-per-player kernels, audience masks and human seats read the canonical columns: the store converts back and
    forth mid-game without a byte changing."""
    cg = games(WEREWOLF, P)
    o = oracle_for(cg)
    n, first, seed = 3000, 77, 21
    t, b = _batch(cg, n, first, seed)
    b.set_compaction(2, 4)
    rec = o.init(n)
    packed_bytes = b.state_device_bytes()

    def both(k):
        b.step(k)
        o.step(rec, first, seed, k)
        got = b.export_state()
        assert np.array_equal(got, rec), _diff(got, rec, first, k)

    both(12)
    b.set_kernel("coop")
    both(5)
    assert b.state_device_bytes() > packed_bytes
    b.set_kernel("tps")
    both(7)
    assert b.state_device_bytes() == packed_bytes
    masks = b.audience_masks()                       # evaluated on the canonical columns
    _, c = _batch(cg, n, first, seed, packed=False)
    c.import_state(rec)
    ref = c.audience_masks()
    assert set(masks) == set(ref) and len(masks) > 0
    for k in masks:
        np.testing.assert_array_equal(masks[k], ref[k])
    both(6)
    assert b.state_device_bytes() == packed_bytes
    # a person sits down in seat 1 of every session and never answers: the action phases wait
    b.set_human_seats(np.full(n, 1, dtype=np.uint32))
    b.step(3)
    for _ in range(3):
        o.step_humans(rec, first, seed, np.full(n, 1, dtype=np.uint32), np.full((n, P), 0xFF, dtype=np.uint8))
    np.testing.assert_array_equal(b.export_state(), rec)
    assert b.state_device_bytes() > packed_bytes
    b.set_human_seats(None)
    both(10)
    assert b.state_device_bytes() == packed_bytes


def test_other_tables_are_refused(games):
    from game_engine_b200.capi import GameEngineError
    for game, P in ((WEREWOLF, 17), (WEREWOLF, 32), (TTL, 4)):
        from game_engine_b200.batch import SessionBatch, Table
        b = SessionBatch(Table(games(game, P)), 100, first_session_id=0, seed=1)
        with pytest.raises(GameEngineError):
            b.set_option("store_packed", 1)
        b.set_option("store_packed", 0)
        b.step(3)


def test_random_tables_up_to_sixteen_players(oracle_for):
    from game_engine_b200.batch import SessionBatch, Table
    from oracle.oracle import Oracle
    ran = 0
    for seed in list(SEEDS) + [1000 + s for s in range(60)]:
        tab = random_table(seed, T.FAMILY_WEREWOLF)
        if tab.n_players > 16:
            continue
        ran += 1
        cg = _Blob(tab)
        o = Oracle(cg.blob)
        n, first = 777, (seed << 33) + 5
        b = SessionBatch(Table(cg), n, first_session_id=first, seed=seed + 1, kernel="tps")
        b.set_option("store_packed", 1)
        if seed % 2 == 0:
            b.set_compaction(2, 4)
        rec = o.init(n)
        ost = o.new_stats()
        for k in range(48):
            b.step(1)
            o.step(rec, first, seed + 1, 1, ost)
            got = b.export_state()
            assert np.array_equal(got, rec), "table seed %d (P=%d) %s\n phases=%s\n preds=%s" % (
                seed, tab.n_players, _diff(got, rec, first, k), tab.phases, tab.preds)
        o.stats_final(rec, ost)
        np.testing.assert_array_equal(b.stats(), ost, err_msg="table seed %d" % seed)
        b.close()
    assert ran >= 20


@pytest.mark.parametrize("P", [8, 12])
@pytest.mark.parametrize("wire_fmt", ["canonical", "dense"])
def test_device_validator_on_the_packed_store(games, oracle_for, wire_fmt, P):
    from game_engine_b200 import wire
    from game_engine_b200.capi import GameEngineError
    cg = games(WEREWOLF, P)
    o = oracle_for(cg)
    n, seed = 4000, 29
    good = _played(o, n, seed, 70)
    rng = np.random.default_rng(8008)
    mut = mutate(rng, good)
    conv = (lambda r: wire.to_dense(cg, r)) if wire_fmt == "dense" else (lambda r: r)
    if wire_fmt == "dense":                                 # what the dense records can say (mask bits above the byte are gone)
        mut = wire.from_dense(cg, wire.to_dense(cg, mut))
    ok = o.validate_records(mut)
    assert 0 < ok.sum() < n
    t, b = _batch(cg, n, 0, seed)
    b.set_wire(wire_fmt)
    with pytest.raises(GameEngineError) as ei:
        b.import_state(conv(mut))
    assert "%d record(s)" % int((~ok).sum()) in str(ei.value)
    want = mut.copy()
    want[~ok] = o.init(1)[0]
    np.testing.assert_array_equal(b.export_state(), conv(want))
    b.step(6)
    o.step(want, 0, seed, 6)
    np.testing.assert_array_equal(b.export_state(), conv(want))
