#!/usr/bin/env python3
"""The bench's steady-state ring as a short program for ncu (not the contract bench).

The ring of independent batches, the streams, the small persistent grids and the re-initialisation rule are the ones
of bench.py; after a warm-up the profiled region (cudaProfilerStart/Stop) is `--steps` passes over the ring.  Used two ways:

    # physical DRAM bytes, reads AND writes: one counter pass, caches left alone (no flush / invalidate between kernels),
    # so the write-backs of one launch are counted in whichever later launch they reach DRAM in; sum over the region
    ncu --cache-control none --clock-control none --profile-from-start off \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file X.csv python tools/range_profile.py

    # the CONCURRENT state (8 streams, launches co-resident): the whole region as one range, kernels not serialised
    ncu --replay-mode range --cache-control none --clock-control none --section SpeedOfLight ... python tools/range_profile.py

Prints one JSON line with the region's device time (CUDA events), counted steps and launches, so the profile can be
tied to the workload."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from game_engine_b200 import compile_game  # noqa: E402
from game_engine_b200.batch import SessionBatch, Table, step_many, step_ring  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="werewolf-(mafia)")
    ap.add_argument("--players", type=int, default=8)
    ap.add_argument("--sessions", type=int, default=1 << 20)
    ap.add_argument("--ring", type=int, default=8)
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--ctas-per-sm", type=int, default=3)
    ap.add_argument("--steps", type=int, default=0, help="profiled passes over the ring (0 = the longest possible game: every batch goes through one whole cycle)")
    ap.add_argument("--warmup", type=int, default=-1, help="passes before the profiled region (-1 = one whole cycle)")
    ap.add_argument("--cap", type=int, default=0)
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--launch", default="streams", choices=["streams", "ring"], help="one launch per batch on --streams streams, or one ring launch per pass")
    ap.add_argument("--store", default="packed", choices=["canonical", "packed"])
    a = ap.parse_args()
    from bench import game_cap
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cg = compile_game(a.game, a.players)
    cap = a.cap or game_cap(cg.family, a.players, cg.table.max_revotes)
    a.steps = a.steps or cap
    a.warmup = cap if a.warmup < 0 else a.warmup
    tab = Table(cg)
    R, N = a.ring, a.sessions
    merged = a.launch == "ring"
    NS = 1 if merged else max(1, min(a.streams, R))
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    ring = [SessionBatch(tab, N, first_session_id=i * N, seed=1, kernel=a.kernel) for i in range(R)]
    for i, b in enumerate(ring):
        b.set_stream(streams[i % NS].cuda_stream)
        b.set_grid(0 if merged else a.ctas_per_sm)
        if cg.family == 1 and a.players <= 16:
            b.set_option("store_packed", 1 if a.store == "packed" else 0)
    age, epoch = [0] * R, [0] * R
    for i, b in enumerate(ring):
        pre = (i * cap) // R
        if pre:
            b.step(pre)
        age[i] = pre

    def passes(n):
        for _ in range(n):
            for i, b in enumerate(ring):
                if age[i] >= cap:
                    epoch[i] += 1
                    b.reset(first_session_id=(epoch[i] * R + i) * N)
                    age[i] = 0
            (step_ring if merged else step_many)(ring, 1)
            for i in range(R):
                age[i] += 1

    passes(a.warmup)
    torch.cuda.synchronize()
    c0 = sum(b.counted_steps() for b in ring)
    l0 = sum(b.launch_count() for b in ring)
    s0 = streams[0]
    torch.cuda.profiler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(s0)
    for st in streams[1:]:
        st.wait_event(ev0)
    passes(a.steps)
    for st in streams[1:]:
        e = torch.cuda.Event()
        e.record(st)
        s0.wait_event(e)
    ev1.record(s0)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1)
    counted = sum(b.counted_steps() for b in ring) - c0
    launches = sum(b.launch_count() for b in ring) - l0
    S = cg.record_size
    print(json.dumps({"game": a.game, "players": a.players, "sessions_per_batch": N, "ring": R, "launch": a.launch, "streams": NS,
                      "ctas_per_sm": a.ctas_per_sm, "passes": a.steps, "region_ms": ms, "counted_steps": counted,
                      "launches": launches, "steps_per_s": counted / (ms * 1e-3),
                      "record_bytes": ring[0].state_device_bytes() // (((N + 31) // 32) * 32), "store": a.store,
                      "algorithmic_bytes": counted * 2 * S}), flush=True)
    for b in ring:
        b.close()
    np.zeros(1)


if __name__ == "__main__":
    main()
