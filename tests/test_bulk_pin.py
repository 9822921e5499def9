"""Bulk pin: Oracle B == the reference's real nodes (Oracle A) on thousands of whole games, every step.

tests/golden/bulk_pin.json is written by `python -m oracle.ref_harness.bulk_pin` in the build container (needs
/root/reference): per case the number of sessions and steps compared, the mismatches found (0) and the SHA-256 of Oracle
B's final records over the case's reproducible (seed, session id) list.  Here, without the reference, Oracle B is re-run on
the same list and held to that digest — i.e. to what agreed with the reference's nodes — and, on the GPU, the CUDA path is
held to the same records.  With the reference present a sample of every case is replayed live."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import has_reference

PIN = os.path.join(os.path.dirname(__file__), "golden", "bulk_pin.json")


def cases():
    with open(PIN) as f:
        return json.load(f)["cases"]


def _id(c):
    return "%s_p%d" % (c["game"], c["players"])


def test_the_committed_run_found_no_mismatch():
    cs = cases()
    assert sum(c["sessions"] for c in cs) >= 2000 and sum(c["steps"] for c in cs) >= 50000
    assert any(c["game"] == "two-truths-and-a-lie" and c["sessions"] >= 1000 for c in cs)
    assert all(c["mismatches"] == 0 for c in cs)


@pytest.mark.parametrize("c", cases(), ids=_id)
def test_oracle_b_still_produces_the_pinned_records(c, games, oracle_for):
    from oracle.ref_harness.bulk_pin import oracle_b_final, pairs
    cg = games(c["game"], c["players"])
    o = oracle_for(cg)
    h = hashlib.sha256()
    for seed, sid in pairs(c["pairs_seed"], c["pairs_drawn"])[:c["sessions"]]:
        h.update(bytes(oracle_b_final(cg, o, seed, sid)))
    assert h.hexdigest() == c["final_records_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("c", cases(), ids=_id)
def test_cuda_path_produces_the_pinned_records(c, games):
    """One single-session batch per (seed, id) pair would be slow; sessions that share a seed can share a batch, so the
    pairs are replayed as batches of ONE session id each under its own seed, 64 of them per case."""
    from game_engine_b200.batch import SessionBatch, Table
    from oracle.ref_harness.bulk_pin import oracle_b_final, pairs
    from oracle.oracle import Oracle
    cg = games(c["game"], c["players"])
    o = Oracle(cg.blob)
    tab = Table(cg)
    for seed, sid in pairs(c["pairs_seed"], c["pairs_drawn"])[:64]:
        want = oracle_b_final(cg, o, seed, sid)
        b = SessionBatch(tab, 1, first_session_id=sid, seed=seed)
        b.step(400)
        got = b.export_state()[0]
        b.close()
        assert np.array_equal(got, want), (seed, sid, got.tolist(), want.tolist())


@pytest.mark.reference
@pytest.mark.skipif(not has_reference(), reason="needs /root/reference")
@pytest.mark.parametrize("c", cases(), ids=_id)
def test_a_sample_replays_against_the_live_reference(c):
    from oracle.ref_harness.bulk_pin import _one, pairs
    for seed, sid in pairs(c["pairs_seed"], c["pairs_drawn"])[:3]:
        steps, bad, _ = _one((c["game"], c["players"], seed, sid))
        assert bad is None and steps > 0, bad
