#!/usr/bin/env python3
"""Phase-regrouping experiment: whole games of a de-synchronising table with different regroup settings.

    python tools/regroup_bench.py [--game werewolf-revote --players 32 --log2n 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from game_engine_b200 import compile_game  # noqa: E402
from game_engine_b200.batch import SessionBatch, Table  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="werewolf-revote")
    ap.add_argument("--players", type=int, default=32)
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--settings", default="0:0,8:3,4:3,2:3,1:3,4:1,4:5")
    a = ap.parse_args()
    torch.cuda.init()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    cg = compile_game(a.game, a.players)
    P = a.players
    cap = 9 * P - 16 + 2 * cg.table.max_revotes * (P - 2)
    n = 1 << a.log2n
    tab = Table(cg)
    print("| regroup every | mixed shift | counted steps | ms | steps/s | launches | unfinished |")
    print("|---|---|---|---|---|---|---|")
    for s in a.settings.split(","):
        every, shift = (int(x) for x in s.split(":"))
        bs = [SessionBatch(tab, n, first_session_id=i * n, seed=5, kernel="tps") for i in range(a.ring)]
        for b in bs:
            b.set_stream(stream.cuda_stream)
            b.set_regroup(every, shift)
            b.step(3)
            b.reset()
            b.clear_stats()
        torch.cuda.synchronize()
        c0 = sum(b.counted_steps() for b in bs)
        l0 = sum(b.launch_count() for b in bs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(cap):
            for b in bs:
                b.step(1)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        counted = sum(b.counted_steps() for b in bs) - c0
        launches = sum(b.launch_count() for b in bs) - l0
        unfinished = sum(int(b.stats()[1]) for b in bs)
        print("| %d | %d | %d | %.3f | %.3e | %d | %d |" % (every, shift, counted, ms, counted / (ms * 1e-3), launches, unfinished), flush=True)
        for b in bs:
            b.close()


if __name__ == "__main__":
    main()
