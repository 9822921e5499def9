#!/usr/bin/env python3
"""Generates tests/golden/*.json.gz: traces of the reference's real nodes (Oracle A) for the parity configs.

Run in the build container (needs /root/reference):   python -m oracle.ref_harness.gen_golden
Each fixture: {"game", "players", "seed", "sid", "trace": [state after 0, 1, 2, ... steps]} with wall-clock
timestamps removed (oracle/ref_harness/driver.py:snapshot)."""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness.driver import run_session  # noqa: E402

CASES = [
    # BASELINE.json configs[0]: two-truths-and-a-lie, 1 session, 4 seeded bots; seeds of SURVEY 8d config 1
    ("two-truths-and-a-lie", 4, 0, 0), ("two-truths-and-a-lie", 4, 1, 0), ("two-truths-and-a-lie", 4, 0xC0FFEE, 0),
    ("two-truths-and-a-lie", 3, 2, 1), ("two-truths-and-a-lie", 7, 5, 1 << 40),
    ("werewolf-(mafia)", 8, 0, 0), ("werewolf-(mafia)", 8, 1, 1), ("werewolf-(mafia)", 8, 7, 123456789012),
    ("werewolf-(mafia)", 8, 20261018, 4242), ("werewolf-(mafia)", 5, 3, 3), ("werewolf-(mafia)", 16, 4, 9),
    ("werewolf-(mafia)", 32, 6, 31),
    # more player counts (minimum table sizes, odd counts, the 24-player bucket) and seeds
    ("werewolf-(mafia)", 4, 21, 1000), ("werewolf-(mafia)", 6, 22, 1001), ("werewolf-(mafia)", 7, 23, 1002),
    ("werewolf-(mafia)", 12, 24, 1003), ("werewolf-(mafia)", 24, 25, 1004),
    ("two-truths-and-a-lie", 5, 26, 1005), ("two-truths-and-a-lie", 8, 27, 1006), ("two-truths-and-a-lie", 16, 28, 1007),
    # third table: the reference's earlier 13-phase werewolf generation (game_draft/), aliased state schema
    ("werewolf-draft", 8, 11, 77), ("werewolf-draft", 6, 12, 1 << 36), ("werewolf-draft", 12, 13, 5),
    # BASELINE config 4's table: the repo's extended game (tie -> re-vote phases), run by the reference's own nodes;
    # sid 3010 uses both re-votes of a day and its third vote is tied again (lowest id among the tied dies)
    ("werewolf-revote", 8, 42, 2001), ("werewolf-revote", 8, 46, 3010), ("werewolf-revote", 5, 44, 2003),
    ("werewolf-revote", 16, 45, 7), ("werewolf-revote", 32, 43, 2002),
    # numeric conditions (`player.total_score < 2` decides who votes) and audience groups with >=, >, negated <=, a
    # three-clause `or`: the repo's two-truths variant, conditions evaluated by plain Python in the stub
    ("two-truths-handicap", 4, 51, 4000), ("two-truths-handicap", 6, 52, 4001), ("two-truths-handicap", 9, 53, 4002),
]


# Games with a PERSON in seat 1 (SPEC D3h): the reference's router logs their messages, its PhaseNode waits for them.
# Fixtures go to tests/golden/human/ and carry what the person typed before every step.
HUMAN_CASES = [
    ("werewolf-(mafia)", 6, 5, 77), ("werewolf-(mafia)", 8, 2, 5), ("werewolf-(mafia)", 16, 8, 1), ("werewolf-revote", 8, 46, 3010),
    ("werewolf-draft", 7, 4, 4), ("two-truths-and-a-lie", 4, 3, 9), ("two-truths-and-a-lie", 6, 1, 2),
]


def main_human():
    out_dir = os.path.join(ROOT, "tests", "golden", "human")
    os.makedirs(out_dir, exist_ok=True)
    for game, P, seed, sid in HUMAN_CASES:
        name = "%s_p%d_seed%d_sid%d.json.gz" % (game.replace("(", "").replace(")", ""), P, seed, sid)
        if os.environ.get("GOLDEN_ONLY_MISSING") and os.path.exists(os.path.join(out_dir, name)):
            continue
        trace = run_session(game, P, seed, sid, human=True)
        msgs = trace[0].pop("_human_messages")
        blob = json.dumps({"game": game, "players": P, "seed": seed, "sid": sid, "human_seats": [1], "human_messages": msgs, "trace": trace},
                          ensure_ascii=False, sort_keys=True, separators=(",", ":")).encode("utf-8")
        with gzip.GzipFile(os.path.join(out_dir, name), "wb", mtime=0) as f:
            f.write(blob)
        print("human/%-54s %3d steps  %6d bytes" % (name, len(trace) - 1, os.path.getsize(os.path.join(out_dir, name))))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--human":
        return main_human()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1] if len(sys.argv) > 1 else None          # optional: regenerate one game's fixtures
    for game, P, seed, sid in CASES:
        if only and game != only:
            continue
        name = "%s_p%d_seed%d_sid%d.json.gz" % (game.replace("(", "").replace(")", ""), P, seed, sid)
        if os.environ.get("GOLDEN_ONLY_MISSING") and os.path.exists(os.path.join(out_dir, name)):
            continue
        trace = run_session(game, P, seed, sid)
        name = "%s_p%d_seed%d_sid%d.json.gz" % (game.replace("(", "").replace(")", ""), P, seed, sid)
        blob = json.dumps({"game": game, "players": P, "seed": seed, "sid": sid, "trace": trace}, ensure_ascii=False,
                          sort_keys=True, separators=(",", ":")).encode("utf-8")
        with gzip.GzipFile(os.path.join(out_dir, name), "wb", mtime=0) as f:
            f.write(blob)
        print("%-60s %3d steps  %6d bytes" % (name, len(trace) - 1, os.path.getsize(os.path.join(out_dir, name))))


if __name__ == "__main__":
    main()
