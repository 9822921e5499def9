#!/usr/bin/env python3
"""Bulk pin of Oracle B against the reference's real nodes (Oracle A): thousands of whole games, every step.

The golden fixtures (tests/golden/*.json.gz) pin a few dozen single sessions; SURVEY 7.2(ii) asks for a bulk check
(1 000 two-truths sessions).  This script, run in the build container (needs /root/reference), plays N sessions of a
game with pseudo-random (seed, session id) pairs through the reference's unmodified InitialRouterNode ->
BotBehaviorNode -> PhaseNode -> RefereeNode (driver.run_session) and through Oracle B + the host adapter
(tests/helpers.replay_records), compares the dict state after EVERY step, and writes tests/golden/bulk_pin.json:

    {"cases": [{"game", "players", "sessions", "steps", "mismatches": 0, "pairs_seed", "final_records_sha256"}]}

`final_records_sha256` is the digest of Oracle B's final records over the same (seed, sid) list, so the CPU test suite
(tests/test_bulk_pin.py) re-derives the list, re-runs Oracle B and holds it to what agreed with the reference here —
without the reference.  Test infrastructure only.

    python -m oracle.ref_harness.bulk_pin [--procs 8]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

OUT = os.path.join(ROOT, "tests", "golden", "bulk_pin.json")

# (game, players, sessions, pairs_seed)
CASES = [
    ("two-truths-and-a-lie", 4, 1000, 1),
    ("werewolf-(mafia)", 8, 1000, 2),              # BASELINE config 2's table
    ("werewolf-(mafia)", 16, 200, 3),
    ("werewolf-revote", 8, 300, 4),                # tie -> re-vote loop
    ("werewolf-draft", 8, 200, 5),                 # third table, aliased schema
    ("two-truths-handicap", 5, 200, 6),            # numeric conditions
]


def pairs(pairs_seed: int, n: int):
    """The (seed, sid) list of a case: reproducible, spread over the full 64-bit ranges."""
    rng = np.random.default_rng(pairs_seed)
    seeds = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    sids = rng.integers(0, 1 << 62, size=n, dtype=np.uint64)
    sids[::7] &= np.uint64(0xFFFF)                 # small ids too
    return [(int(a), int(b)) for a, b in zip(seeds, sids)]


def oracle_b_final(cg, o, seed: int, sid: int, cap: int = 400):
    rec = o.init(1)
    kinds = [p.kind for p in cg.table.phases]
    for _ in range(cap):
        if kinds[rec[0, 0]] == 3:
            break
        o.step(rec, sid, seed, 1)
    return rec[0].copy()


def _one(args):
    game, P, seed, sid = args
    import json as _json
    from game_engine_b200 import compile_game
    from helpers import first_diff, oracle_b_records, replay_records
    from oracle.oracle import Oracle
    from oracle.ref_harness.driver import run_session
    cg = compile_game(game, P)
    o = Oracle(cg.blob)
    live = _json.loads(_json.dumps(run_session(game, P, seed, sid)))
    recs = oracle_b_records(o, sid, seed, len(live) - 1)
    mine = _json.loads(_json.dumps(replay_records(cg, recs)))
    bad = None
    for k, (want, got) in enumerate(zip(live, mine)):
        d = first_diff(want, got)
        if d is not None:
            bad = "seed %d sid %d step %d: %s" % (seed, sid, k, d)
            break
    if bad is None and cg.table.phases[recs[-1][0]].kind != 3:
        bad = "seed %d sid %d: the reference's game ended, Oracle B's did not" % (seed, sid)
    return len(live) - 1, bad, bytes(recs[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of every case's sessions (quick runs)")
    a = ap.parse_args()
    out = {"generator": "python -m oracle.ref_harness.bulk_pin", "cases": []}
    with mp.Pool(a.procs) as pool:
        for game, P, n, ps in CASES:
            n_run = max(1, int(n * a.scale))
            t0 = time.time()
            pr = pairs(ps, n)[:n_run]
            res = pool.map(_one, [(game, P, s, i) for s, i in pr], chunksize=4)
            steps = sum(r[0] for r in res)
            bad = [r[1] for r in res if r[1]]
            h = hashlib.sha256()
            for r in res:
                h.update(r[2])
            print("%-22s P=%-2d %5d sessions %7d steps  mismatches %d  (%.0f s)" % (game, P, n_run, steps, len(bad), time.time() - t0), flush=True)
            for b in bad[:5]:
                print("   ", b)
            out["cases"].append({"game": game, "players": P, "sessions": n_run, "steps": steps, "mismatches": len(bad), "pairs_seed": ps,
                                 "pairs_drawn": n, "final_records_sha256": h.hexdigest()})
    if a.scale == 1.0:
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
            f.write("\n")
        print("wrote", OUT)
    return 1 if any(c["mismatches"] for c in out["cases"]) else 0


if __name__ == "__main__":
    sys.exit(main())
