"""Trace export / replay: frames -> on-wire AgentState objects -> JSON lines -> back."""
import glob
import gzip
import json
import os

import numpy as np
import pytest

from helpers import first_diff, normalise, oracle_b_records, pick

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json.gz")))


def _load(path):
    with gzip.open(path, "rt", encoding="utf-8") as f:
        return json.load(f)


@pytest.mark.parametrize("path", pick(GOLDEN, "lie_p3_", "lie_p4_seed0_", "draft_p8_", "mafia_p8_seed0_"), ids=lambda p: os.path.basename(p)[:-8])
def test_materialised_trace_equals_reference_nodes(path, games, oracle_for, tmp_path):
    """The wire objects built from the frames carry exactly the dict state the reference's own nodes produced
    (golden fixtures), plus the UI-facing keys of the TS AgentState."""
    from game_engine_b200 import trace as tr
    g = _load(path)
    cg = games(g["game"], g["players"])
    frames = np.stack(oracle_b_records(oracle_for(cg), g["sid"], g["seed"], len(g["trace"]) - 1))
    states = tr.materialise(cg, frames)
    assert len(states) == len(g["trace"])
    for k, (want, got) in enumerate(zip(g["trace"], states)):
        assert set(tr.WIRE_KEYS) <= set(got)
        d = first_diff(want, json.loads(json.dumps(normalise(got))))
        assert d is None, "state %d: %s" % (k, d)
        assert got["stateVersion"] == k
    last = states[-1]
    if cg.family == 1:
        dead = {pid for pid, ps in last["player_states"].items() if not ps["is_alive"]}
        assert set(last["deadPlayers"]) == dead and dead
    assert last["vote"], "votes were cast during the game"
    assert all(set(v) == {"voteid", "playerid", "option"} for v in last["vote"])
    # JSON lines round trip keeps the records
    p = str(tmp_path / "t.jsonl")
    tr.write_jsonl(p, {"game": g["game"], "players": g["players"], "seed": g["seed"], "session_id": g["sid"]}, states)
    hdr, back = tr.read_jsonl(p)
    assert hdr["trace"] == "game_engine_b200/1" and hdr["seed"] == g["seed"]
    np.testing.assert_array_equal(tr.frames_of(back), frames)


def test_trace_must_start_at_creation(games, oracle_for):
    from game_engine_b200 import trace as tr
    cg = games("werewolf-(mafia)", 8)
    frames = np.stack(oracle_b_records(oracle_for(cg), 0, 0, 5))
    with pytest.raises(ValueError):
        tr.materialise(cg, frames[1:])


@pytest.mark.gpu
@pytest.mark.parametrize("game,P", [("werewolf-(mafia)", 8), ("two-truths-and-a-lie", 4), ("werewolf-revote", 8)])
def test_gpu_trace_and_replay(games, oracle_for, game, P, tmp_path):
    from game_engine_b200 import trace as tr
    from game_engine_b200.batch import SessionBatch, Table
    cg = games(game, P)
    o = oracle_for(cg)
    seed, sid = 77, (1 << 34) + 9
    path = str(tmp_path / "g.jsonl")
    states = tr.export_session(game, P, seed, sid, path=path)
    want = np.stack(oracle_b_records(o, sid, seed, len(states) - 1))
    np.testing.assert_array_equal(tr.frames_of(states), want)
    assert cg.table.phases[want[-1][0]].kind == 3              # ran to the terminal phase, no trailing no-op frames
    assert want[-1][2] | (want[-1][3] << 8) == len(states) - 1
    hdr, back = tr.read_jsonl(path)
    assert tr.replay_check(cg, hdr["seed"], hdr["session_id"], back) == len(states) - 1
    assert tr.replay_check(cg, seed, sid, back, start=len(back) // 2) == len(states) - 1 - len(back) // 2
    # a tampered trace is caught
    bad = json.loads(json.dumps(back))
    rec = bytearray(bytes.fromhex(bad[-2]["record"]))
    rec[8] ^= 1
    bad[-2]["record"] = rec.hex()
    with pytest.raises(ValueError):
        tr.replay_check(cg, seed, sid, bad)
    # window traces of a larger batch: frames of session i equal a single-session run
    tab = Table(cg)
    b = SessionBatch(tab, 500, first_session_id=sid - 3, seed=seed)
    fr = b.trace(12, first=3, count=2)
    np.testing.assert_array_equal(fr[:, 0, :], want[:13])
