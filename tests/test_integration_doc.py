"""The code in INTEGRATION.md must keep working: the batch API walk-through, the trace example and the vendored
ctypes stub are executed as written (GPU), and every C entry point the document names exists (CPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOC = open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8").read()


def test_every_entry_point_named_in_the_document_is_declared():
    header = open(os.path.join(ROOT, "include", "game_engine_b200.h"), encoding="utf-8").read()
    named = set(re.findall(r"\b(ge_[a-z_]+)\b", DOC))
    declared = set(re.findall(r"\b(ge_[a-z_]+)\s*\(", header))
    assert named and named <= declared | {"ge_table", "ge_batch"}, sorted(named - declared)


@pytest.mark.gpu
def test_batch_walkthrough_and_trace_example(tmp_path):
    from game_engine_b200 import compile_game, trace
    from game_engine_b200.batch import SessionBatch, Table, step_many
    cg = compile_game("werewolf-(mafia)", n_players=8)
    tab = Table(cg)
    b = SessionBatch(tab, n_sessions=1 << 12, first_session_id=0, seed=7, device=0)
    b.step(64)
    stats = b.stats()
    records = b.export_state()
    assert records.shape == (1 << 12, 56) and stats[1] == 0 and stats[2] + stats[3] == 1 << 12
    others = [SessionBatch(tab, 1 << 10, first_session_id=(j + 1) << 20, seed=7) for j in range(2)]
    for o in others:
        o.set_grid(3)
    step_many(others, 5)
    assert set(b.audience_masks()) == set(cg.audience_preds)
    assert b.trace(3, first=10, count=4).shape == (4, 4, 56)
    path = str(tmp_path / "room42.jsonl")
    states = trace.export_session("werewolf-(mafia)", players=8, seed=7, session_id=42, path=path)
    hdr, back = trace.read_jsonl(path)
    assert trace.replay_check(cg, hdr["seed"], hdr["session_id"], back) == len(states) - 1


@pytest.mark.gpu
def test_vendored_ctypes_stub_as_written():
    from game_engine_b200 import compile_game
    from game_engine_b200.capi import LIB_PATH
    blob = compile_game("two-truths-and-a-lie", 4).blob
    n, first_sid, seed = 100, 5, 9
    out = np.zeros((n, 24), dtype=np.uint8)
    L = ctypes.CDLL(LIB_PATH)
    vp, u64 = ctypes.c_void_p, ctypes.c_uint64
    L.ge_table_create.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    L.ge_batch_create.argtypes = [vp, ctypes.c_int, u64, u64, u64, ctypes.POINTER(vp)]
    L.ge_step.argtypes = [vp, ctypes.c_int, vp]
    L.ge_export_state.argtypes = [vp, u64, u64, vp]
    L.ge_last_error.restype = ctypes.c_char_p
    tab, bat = vp(), vp()
    assert L.ge_table_create(blob, len(blob), ctypes.byref(tab)) == 0, L.ge_last_error()
    assert L.ge_batch_create(tab, 0, n, first_sid, seed, ctypes.byref(bat)) == 0, L.ge_last_error()
    assert L.ge_step(bat, 1, None) == 0
    assert L.ge_export_state(bat, 0, n, out.ctypes.data) == 0
    assert (out[:, 2] == 1).all()              # every session took its first step
