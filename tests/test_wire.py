"""Dense wire format (SPEC.md section 5b) and record validation (section 7b).

CPU part: the NumPy converters are inverse to each other on Oracle B's records.  GPU part: the library's export /
import kernels produce exactly those bytes, a host-buffer call gives the same results in either format, and the
device-side validator agrees with Oracle B's on thousands of mutated records; accepted mutants step identically."""
import numpy as np
import pytest

from game_engine_b200 import wire

WEREWOLF, TTL, REVOTE = "werewolf-(mafia)", "two-truths-and-a-lie", "werewolf-revote"


def _played(o, n, seed, steps):
    """Records of n sessions at staggered depths of their games (so every phase and field value occurs)."""
    rec = o.init(n)
    for k in range(steps):
        o.step(rec[(k * n) // steps:], (k * n) // steps, seed, 1)
    return rec


@pytest.mark.parametrize("game,P", [(WEREWOLF, 5), (WEREWOLF, 8), (WEREWOLF, 13), (REVOTE, 16), (WEREWOLF, 24), (TTL, 4)])
def test_dense_is_a_bijection_on_reachable_records(games, oracle_for, game, P):
    cg = games(game, P)
    rec = _played(oracle_for(cg), 1500, 11, 60)
    d = wire.to_dense(cg, rec)
    assert d.shape == (1500, wire.dense_record_size(cg))
    assert wire.dense_padding_is_zero(cg, d).all()
    np.testing.assert_array_equal(wire.from_dense(cg, d), rec)
    if P <= 8 and cg.family == 1:
        assert d.shape[1] == 32
    elif P <= 16 and cg.family == 1:
        assert d.shape[1] == 48
    else:
        assert d.shape[1] == cg.record_size


def mutate(rng, rec):
    """Byte-level mutations of well-formed records: most stay well-formed, many do not."""
    out = rec.copy()
    n, S = out.shape
    for i in range(n):
        for _ in range(int(rng.integers(1, 4))):
            j = int(rng.integers(0, S))
            r = rng.random()
            if r < 0.4:
                out[i, j] ^= 1 << int(rng.integers(0, 8))
            elif r < 0.7:
                out[i, j] = int(rng.integers(0, 256))
            else:
                out[i, j] = int(rng.choice([0, 1, 2, 4, 8, 31, 32, 33, 127, 128, 255]))
    return out


@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 5), (REVOTE, 32), (TTL, 4), (TTL, 7)])
def test_oracle_validator_accepts_every_reachable_record(games, oracle_for, game, P):
    cg = games(game, P)
    o = oracle_for(cg)
    rec = _played(o, 2000, 3, 70)
    assert o.validate_records(rec).all()
    bad = mutate(np.random.default_rng(1), rec)
    ok = o.validate_records(bad)
    assert 0 < ok.sum() < len(ok)            # the mutator produces both kinds


# ------------------------------------------------------------------------------------------------ GPU
def _batch(cg, n, first, seed, kernel="tps"):
    from game_engine_b200.batch import SessionBatch, Table
    t = Table(cg)
    return t, SessionBatch(t, n, first_session_id=first, seed=seed, kernel=kernel)


@pytest.mark.gpu
@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 6), (REVOTE, 16), (WEREWOLF, 11), (WEREWOLF, 32), (TTL, 4)])
def test_library_dense_export_import(games, oracle_for, game, P):
    cg = games(game, P)
    o = oracle_for(cg)
    n, seed = 3001, 17
    t, b = _batch(cg, n, 0, seed)
    assert b.wire_record_size == cg.record_size
    b.step(23)                                   # compaction has permuted the slots by now (werewolf): export undoes it
    canon = b.export_state()
    b.set_wire("dense")
    assert b.wire_record_size == wire.dense_record_size(cg)
    dense = b.export_state()
    np.testing.assert_array_equal(dense, wire.to_dense(cg, canon))
    np.testing.assert_array_equal(b.export_state(1000, 77), wire.to_dense(cg, canon[1000:1077]))
    # dense import into a fresh batch, continue, compare with the oracle
    t2, c = _batch(cg, n, 0, seed)
    c.set_wire("dense")
    c.import_state(dense)
    c.step(9)
    rec = o.init(n)
    o.step(rec, 0, seed, 32)
    c.set_wire("canonical")
    np.testing.assert_array_equal(c.export_state(), rec)


@pytest.mark.gpu
@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 16)])
def test_run_host_dense_equals_canonical(games, oracle_for, game, P):
    cg = games(game, P)
    o = oracle_for(cg)
    n, seed = 4096, 23
    rec = o.init(n)
    t, b = _batch(cg, n, 0, seed)
    b.set_wire("dense")
    b.set_host_fused(True)
    din = wire.to_dense(cg, rec)
    dout = np.empty_like(din)
    st = np.zeros(560, dtype=np.uint64)
    b.run_host(din, dout, 200, st)
    ost = o.new_stats()
    o.step(rec, 0, seed, 200, ost)
    o.stats_final(rec, ost)
    np.testing.assert_array_equal(wire.from_dense(cg, dout), rec)
    np.testing.assert_array_equal(st, ost)
    with pytest.raises(ValueError):              # a canonical-sized buffer is refused when the batch speaks dense
        b.run_host(rec, None, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 5), (WEREWOLF, 14), (REVOTE, 32), (TTL, 4), (TTL, 7)])
def test_device_validator_agrees_with_the_oracle(games, oracle_for, game, P):
    """Every import path rejects exactly the records Oracle B's validator rejects, names the first one, and
    replaces them by initial records; the accepted mutants then step bit for bit like the oracle."""
    from game_engine_b200.capi import GameEngineError
    cg = games(game, P)
    o = oracle_for(cg)
    n, seed = 4000, 29
    good = _played(o, n, seed, 70)
    rng = np.random.default_rng(P * 1000 + cg.family)
    mut = mutate(rng, good)
    ok = o.validate_records(mut)
    assert 0 < ok.sum() < n
    t, b = _batch(cg, n, 0, seed)
    with pytest.raises(GameEngineError) as ei:
        b.import_state(mut)
    msg = str(ei.value)
    assert "%d record(s)" % int((~ok).sum()) in msg and "index %d)" % int(np.flatnonzero(~ok)[0]) in msg, msg
    want = mut.copy()
    want[~ok] = o.init(1)[0]
    np.testing.assert_array_equal(b.export_state(), want)
    # accepted mutants are states no game reaches; the kernels and the oracle still agree on them
    for kernel in ("tps", "tps_generic", "coop"):
        b.set_kernel(kernel)
        b.import_state(want)
        b.step(6)
        ref = want.copy()
        o.step(ref, 0, seed, 6)
        got = b.export_state()
        assert np.array_equal(got, ref), "%s: first differing record %d\n in=%s\ngot=%s\nref=%s" % (
            kernel, np.flatnonzero((got != ref).any(axis=1))[0], want[np.flatnonzero((got != ref).any(axis=1))[0]].tolist(),
            got[np.flatnonzero((got != ref).any(axis=1))[0]].tolist(), ref[np.flatnonzero((got != ref).any(axis=1))[0]].tolist())
    # the asynchronous host-buffer call reports at the next synchronising call
    out = np.empty_like(mut)
    b.run_host_async(mut, out, 0)
    with pytest.raises(GameEngineError):
        b.sync()
    np.testing.assert_array_equal(out, want)
    b.sync()                                     # the verdict is reported once


@pytest.mark.gpu
@pytest.mark.parametrize("game,P", [(WEREWOLF, 8), (WEREWOLF, 7), (REVOTE, 16), (WEREWOLF, 12)])
def test_dense_validator(games, oracle_for, game, P):
    """Dense records: rejected iff the padding is not zero or the canonical record they decode to is malformed."""
    from game_engine_b200.capi import GameEngineError
    cg = games(game, P)
    o = oracle_for(cg)
    n, seed = 4000, 31
    dense = wire.to_dense(cg, _played(o, n, seed, 70))
    mut = mutate(np.random.default_rng(P), dense)
    ok = o.validate_records(wire.from_dense(cg, mut)) & wire.dense_padding_is_zero(cg, mut)
    assert 0 < ok.sum() < n
    t, b = _batch(cg, n, 0, seed)
    b.set_wire("dense")
    with pytest.raises(GameEngineError) as ei:
        b.import_state(mut)
    assert "%d record(s)" % int((~ok).sum()) in str(ei.value), str(ei.value)
    want = mut.copy()
    want[~ok] = wire.to_dense(cg, o.init(1))[0]
    np.testing.assert_array_equal(b.export_state(), want)
