"""BASELINE.json configurations at their FULL sizes on one GPU (configs 2-5: 2^20 x 8 players, 2^24 x 16 players,
2^26 x 32 players with re-votes, 2^28 x TTL), checked through size-independent properties:

* every game ends; the statistics are internally consistent (winners, lengths and visits all add up);
* windows of sessions sampled across the batch equal Oracle B run on those session ids alone (bit-exact);
* sharding is invisible: the statistics of the whole batch equal the sum over two contiguous shards ("checksum of
  checksums") — the property that makes the multi-GPU all-reduce exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    ("werewolf-(mafia)", 8, 1 << 20),
    ("werewolf-(mafia)", 16, 1 << 24),
    ("werewolf-revote", 32, 1 << 26),
    ("two-truths-and-a-lie", 4, 1 << 28),
]


def _cap(cg, game, P):
    return 2 + 8 * P if game.startswith("two") else 9 * P - 16 + 2 * cg.table.max_revotes * (P - 2)


@pytest.mark.parametrize("game,P,n", CASES, ids=["cfg2-1M-p8", "cfg3-16M-p16", "cfg4-64M-p32-revote", "cfg5-256M-ttl"])
def test_full_size_run(games, oracle_for, game, P, n):
    from game_engine_b200.batch import SessionBatch, Table
    cg = games(game, P)
    o = oracle_for(cg)
    tab = Table(cg)
    first, seed, cap = 3 << 40, 20261018, _cap(cg, game, P)
    b = SessionBatch(tab, n, first_session_id=first, seed=seed)
    b.step(cap + 5)                 # the longest possible game, plus no-op launches up to the next compaction check
    st = b.stats()
    kinds = np.array([p.kind for p in cg.table.phases])
    # ---- internal consistency
    assert st[1:4].sum() == n and st[4:260].sum() == n and st[260:292].sum() == st[0]
    if cg.family == 1:
        assert st[1] == 0 and st[2] > 0 and st[3] > 0            # nobody unfinished, both sides win some games
        assert st[292:548].sum() == n
        assert b.active() == 0
    else:
        assert st[1] == 0 and st[2] == n and st[0] == n * cap  # fixed-length games
        assert st[292:548].sum() == n * P and (np.arange(256) * st[292:548]).sum() == n * P * (P - 1)
    mean_len = (np.arange(256) * st[4:260]).sum() / n
    assert abs(mean_len * n - st[0]) < 1e-6 * st[0] or st[259] > 0       # counted steps = sum of game lengths
    # ---- sampled windows == the oracle on those ids
    w = 1024
    for off in (0, n // 3 + 17, n - w):
        got = b.export_state(off, w)
        rec = o.init(w)
        o.step(rec, first + off, seed, cap)
        np.testing.assert_array_equal(got, rec, err_msg="window at %d" % off)
        assert (kinds[got[:, 0]] == 3).all()
    b.close()
    # ---- checksum of checksums: two contiguous shards (what two ranks would hold)
    if n <= 1 << 24:
        cut = n // 2 + 4096
        lo = SessionBatch(tab, cut, first_session_id=first, seed=seed)
        hi = SessionBatch(tab, n - cut, first_session_id=first + cut, seed=seed)
        lo.step(cap)
        hi.step(cap)
        np.testing.assert_array_equal(lo.stats() + hi.stats(), st)
        lo.close()
        hi.close()
