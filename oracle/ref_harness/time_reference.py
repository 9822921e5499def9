#!/usr/bin/env python3
"""Times Oracle A — the reference's REAL nodes (InitialRouterNode -> BotBehaviorNode -> PhaseNode -> RefereeNode of
/root/reference/agent/game_agent_v2.py) with the LLM replaced by the rule-following stub — on this machine's CPU.

This is BASELINE.json configs[0] (two-truths-and-a-lie, 4 seeded bots, LLM calls stubbed, reference stepping on
CPU) and the same for 8-player werewolf.  It needs /root/reference, so it runs in the build container only; the
result is recorded in profiles/.  Test infrastructure (lives under oracle/).

    python -m oracle.ref_harness.time_reference [--sessions 40]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness.driver import run_session  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sessions", type=int, default=40)
    a = ap.parse_args()
    print("| game | players | sessions | session-phase-steps | seconds | steps/s (1 core, no LLM latency) |")
    print("|---|---|---|---|---|---|")
    for game, P in (("two-truths-and-a-lie", 4), ("werewolf-(mafia)", 8)):
        run_session(game, P, 0, 0)                      # import + warm-up
        steps = 0
        t0 = time.perf_counter()
        for sid in range(a.sessions):
            steps += len(run_session(game, P, 1, sid)) - 1
        dt = time.perf_counter() - t0
        print("| %s | %d | %d | %d | %.2f | %.0f |" % (game, P, a.sessions, steps, dt, steps / dt))


if __name__ == "__main__":
    main()
