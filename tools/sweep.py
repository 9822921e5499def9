#!/usr/bin/env python3
"""Sweeps for BASELINE.json configs 3-5 on ONE GPU: batch-size sweep (two-truths-and-a-lie, 2^10 .. 2^26
sessions) and the larger werewolf tables (P = 16, 32, re-vote variant).  Whole games from fresh sessions,
ring of batches so inputs come from HBM when the footprint allows; prints a markdown table.

    python tools/sweep.py [--out profiles/r01_sweep.md]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from game_engine_b200 import compile_game  # noqa: E402
from game_engine_b200.batch import SessionBatch, Table  # noqa: E402


def run(game, P, n, kernel, cap, ring, stream):
    cg = compile_game(game, P)
    tab = Table(cg)
    fused = kernel.endswith("-fused")           # whole game in ONE launch per batch (ge_run_fused), state in registers
    kernel = kernel.replace("-fused", "")
    bs = [SessionBatch(tab, n, first_session_id=i * n, seed=5, kernel=kernel) for i in range(ring)]
    for b in bs:
        b.set_stream(stream.cuda_stream)
        b.step(3)
        b.reset()
        b.clear_stats()
    torch.cuda.synchronize()
    c0 = sum(b.counted_steps() for b in bs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if fused:
        for b in bs:
            b.run_fused(cap)
    else:
        for _ in range(cap):
            for b in bs:
                b.step(1)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    counted = sum(b.counted_steps() for b in bs) - c0
    st = bs[0].stats()
    for b in bs:
        b.close()
    return counted, ms, cg.record_size, int(st[1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--max-log2", type=int, default=26)
    a = ap.parse_args()
    torch.cuda.init()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    rows = []
    for lg in range(10, a.max_log2 + 1, 2):
        n = 1 << lg
        ring = max(1, min(8, (512 << 20) // (n * 24)))
        c, ms, S, unfinished = run("two-truths-and-a-lie", 4, n, "tps", 34, ring, stream)
        rows.append(("two-truths-and-a-lie", 4, n, ring, "tps", c, ms, S, unfinished))
        if lg <= 20:                            # launch-bound sizes: the fused launch is the right call
            c, ms, S, unfinished = run("two-truths-and-a-lie", 4, n, "tps-fused", 34, ring, stream)
            rows.append(("two-truths-and-a-lie", 4, n, ring, "tps-fused", c, ms, S, unfinished))
    for game, P, n, cap in (("werewolf-(mafia)", 8, 1 << 20, 56), ("werewolf-(mafia)", 16, 1 << 20, 128),
                            ("werewolf-(mafia)", 32, 1 << 20, 272), ("werewolf-revote", 32, 1 << 20, 392),
                            ("werewolf-(mafia)", 16, 1 << 24, 128), ("werewolf-(mafia)", 32, 1 << 23, 272)):
        for kernel in ("tps", "coop"):
            if kernel == "coop" and n > (1 << 20):
                continue
            ring = max(1, min(4, (512 << 20) // (n * (48 + P))))
            c, ms, S, unfinished = run(game, P, n, kernel, cap, ring, stream)
            rows.append((game, P, n, ring, kernel, c, ms, S, unfinished))
    lines = ["| game | P | sessions/batch | ring | kernel | counted steps | ms | steps/s | algorithmic GB/s | unfinished |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    for game, P, n, ring, kernel, c, ms, S, unf in rows:
        lines.append("| %s | %d | 2^%d | %d | %s | %d | %.3f | %.3e | %.0f | %d |"
                     % (game, P, n.bit_length() - 1, ring, kernel, c, ms, c / (ms * 1e-3), c * 2 * S / (ms * 1e-3) / 1e9, unf))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        with open(a.out, "w") as f:
            f.write("# sweep (tools/sweep.py): whole games from fresh sessions, one B200, CUDA events on the launching stream\n\n" + txt + "\n")


if __name__ == "__main__":
    main()
