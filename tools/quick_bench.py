#!/usr/bin/env python3
"""Exploration timer (not the contract bench): per-launch times of the step kernels."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from game_engine_b200 import compile_game  # noqa: E402
from game_engine_b200.batch import Table, SessionBatch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="werewolf-(mafia)")
    ap.add_argument("--players", type=int, default=8)
    ap.add_argument("--sessions", type=int, default=1 << 20)
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--kernels", default="tps,coop")
    a = ap.parse_args()
    torch.cuda.init()
    cg = compile_game(a.game, a.players)
    S = cg.record_size
    tab = Table(cg)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    assert st != 0
    for kern in a.kernels.split(","):
        ring = [SessionBatch(tab, a.sessions, first_session_id=i * a.sessions, seed=1, kernel=kern) for i in range(a.ring)]
        for b in ring:
            b.step(3, st)
        torch.cuda.synchronize()
        for b in ring:
            b.reset()
            b.sync()
        torch.cuda.synchronize()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)] for _ in ring]
        prev = [b.counted_steps() for b in ring]
        per_step = []
        t0 = time.time()
        for k in range(a.steps):
            for i, b in enumerate(ring):
                evs[i][k].record()
                b.step(1, st)
                if k == a.steps - 1:
                    pass
            # events: record the end marker once per (ring, step) using next step's start
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        torch.cuda.synchronize()
        wall = time.time() - t0
        total = sum(b.counted_steps() for b in ring) - sum(prev)
        # per-step times: difference between consecutive event records in issue order
        order = [evs[i][k] for k in range(a.steps) for i in range(len(ring))] + [end]
        times = [order[j].elapsed_time(order[j + 1]) for j in range(len(order) - 1)]
        gpu_ms = order[0].elapsed_time(end)
        print("kernel=%s game=%s P=%d N=%d ring=%d steps=%d" % (kern, a.game, a.players, a.sessions, a.ring, a.steps))
        print("  total counted=%d gpu_ms=%.3f wall_ms=%.1f  steps/s=%.3e  alg GB/s=%.1f"
              % (total, gpu_ms, wall * 1e3, total / (gpu_ms * 1e-3), total * 2 * S / (gpu_ms * 1e-3) / 1e9))
        by_step = [sum(times[k * len(ring):(k + 1) * len(ring)]) / len(ring) for k in range(a.steps)]
        print("  us per launch by step:", " ".join("%.1f" % (t * 1e3) for t in by_step))
        full = a.sessions * 2 * S
        print("  alg GB/s by step (if all live):", " ".join("%.0f" % (full / (t * 1e-3) / 1e9) for t in by_step[:12]))
        for b in ring:
            b.close()


if __name__ == "__main__":
    main()
