#!/usr/bin/env python3
"""BASELINE config 5: two-truths-and-a-lie batch-size sweep, 2^10 .. 2^28 sessions in total, at 1 / 2 / 4 / 8 GPUs, next to
the stubbed CPU harness (Oracle B) on the same N.

    python tools/ttl_sweep.py --out profiles/r02_ttl_sweep_g1.md                 # one GPU, with the CPU column
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/ttl_sweep.py --no-cpu ...

Whole games from fresh sessions (34 steps: every game has the same length).  The total is sharded by contiguous
session ids over the ranks; every rank steps its shard with single-step launches (and, up to 2^22 sessions per rank,
also with one fused launch); time = max over ranks of the CUDA-event time between two barriers.  The CPU column runs
Oracle B with OpenMP on all host threads for N up to 2^24 and is extrapolated linearly (labelled) above."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from game_engine_b200 import compile_game  # noqa: E402
from game_engine_b200.batch import SessionBatch, Table  # noqa: E402
from game_engine_b200.parallel import shard_range  # noqa: E402


def render(rows, world):
    lines = ["# two-truths-and-a-lie batch-size sweep (tools/ttl_sweep.py), %d GPU(s): whole games (34 steps) from fresh sessions" % world, "",
             "GPU time = max over ranks of the CUDA-event time between barriers, best of 3; CPU = Oracle B (oracle/ge_oracle.c, OpenMP, all host "
             "threads; rows above 2^24 sessions are linear extrapolations of the 2^24 run and say so).", "",
             "| total sessions | single-step launches: steps/s | algorithmic GB/s per GPU | fused launch: steps/s | CPU steps/s | best GPU / CPU |",
             "|---|---|---|---|---|---|"]
    for r in rows:
        f = ("%.3e" % r["fused"]["steps_per_s"]) if "fused" in r else "—"
        c = r.get("cpu")
        best_gpu = max(r["single_step"]["steps_per_s"], r.get("fused", {}).get("steps_per_s", 0))
        cs = ("%.3e (%d threads%s)" % (c["steps_per_s"], c["threads"], ", extrapolated" if c["extrapolated"] else "")) if c else "—"
        ratio = ("%.0fx" % (best_gpu / c["steps_per_s"])) if c else "—"
        lines.append("| 2^%d | %.3e | %.0f | %s | %s | %s |" % (r["log2_sessions"], r["single_step"]["steps_per_s"], r["single_step"]["alg_GBps_per_gpu"], f, cs, ratio))
    return "\n".join(lines) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--render", default="", help="render a saved JSON-lines output of this tool to --out (no GPU needed)")
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=28)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.render:
        rows = [json.loads(x) for x in open(a.render) if x.startswith("{")]
        open(a.out, "w").write(render(rows, rows[0]["gpus"]))
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cg = compile_game("two-truths-and-a-lie", 4)
    tab = Table(cg)
    cap = 34
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def timed(n_total, fused):
        lo, n = shard_range(n_total, world, rank)
        best = None
        for rep in range(3):
            b = SessionBatch(tab, max(n, 1), first_session_id=lo, seed=5, device=local)
            b.set_stream(stream.cuda_stream)
            if rep == 0:
                b.step(2)
                b.reset()
                b.clear_stats()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if n > 0:
                b.run_fused(cap) if fused else b.step(cap)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            counted = b.counted_steps() if n > 0 else 0
            b.close()
            t = torch.tensor([ms, float(counted)], dtype=torch.float64, device=dev)
            if world > 1:
                tmax, tsum = t.clone(), t.clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
                ms, counted = float(tmax[0]), float(tsum[1])
            if best is None or ms < best[0]:
                best = (ms, counted)
        return best

    cpu = {}
    if rank == 0 and not a.no_cpu:
        from oracle import oracle as _o
        try:
            _o.build(native=True)
            native = True
        except Exception:
            native = False
        from oracle.oracle import Oracle
        o = Oracle(cg.blob, native=native)
        threads = max(int(o.max_threads()), len(os.sched_getaffinity(0)))
        for lg in range(a.min_log2, min(a.max_log2, 24) + 1, 2):
            n = 1 << lg
            best = None
            for _ in range(3 if lg <= 20 else 1):
                rec = o.init(n)
                st = o.new_stats()
                t0 = time.perf_counter()
                o.step(rec, 0, 5, cap, st, threads=threads)
                dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, int(st[0]))
            cpu[lg] = {"steps_per_s": best[1] / best[0], "seconds": best[0], "threads": threads, "extrapolated": False}
        if 24 in cpu:
            for lg in range(26, a.max_log2 + 1, 2):
                cpu[lg] = {"steps_per_s": cpu[24]["steps_per_s"], "seconds": cpu[24]["seconds"] * (1 << (lg - 24)), "threads": threads, "extrapolated": True}
    rows = []
    for lg in range(a.min_log2, a.max_log2 + 1, 2):
        n = 1 << lg
        ms, counted = timed(n, False)
        row = {"log2_sessions": lg, "gpus": world, "single_step": {"ms": ms, "steps_per_s": counted / (ms * 1e-3), "alg_GBps_per_gpu": counted * 48 / (ms * 1e-3) / 1e9 / world}}
        if n // world <= (1 << 22):
            ms2, c2 = timed(n, True)
            row["fused"] = {"ms": ms2, "steps_per_s": c2 / (ms2 * 1e-3)}
        if lg in cpu:
            row["cpu"] = cpu[lg]
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if rank == 0 and a.out:
        open(a.out, "w").write(render(rows, world))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
