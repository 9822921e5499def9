#!/usr/bin/env python3
"""bench.py — session-phase-steps/sec of the batched referee/phase step (contract in the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...     # CPU arm (oracle port, all host threads)
    torchrun --nproc-per-node N ... bench.py --gpus N ...           # one rank per GPU, NCCL

Workload (BASELINE.json configs[1]): werewolf-(mafia).yaml, 8 players, 2^20 sessions per batch per GPU, synthetic
(Philox bots).  To keep the inputs out of L2 the resident data set is a RING of batches whose footprint exceeds
2x L2.  A "step" is ONE PASS of the hot path over that data set: one launch of the step kernel per batch of the
ring (every non-terminal session advances one phase; the state is read and written once) — the same unit in the
CUDA arm and in the CPU reference arm.  To make the number independent of K the ring is a steady state: batch i
starts i*G/R steps into its games and a batch that has been stepped G times (the longest possible game) is
re-initialised with fresh session ids (the re-initialisation kernels are inside the timed region and counted in
gpu_launches).  value = counted session-phase-steps (terminal sessions do not count) of all ranks /
max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "session_phase_steps_per_sec"
UNIT = "session-phase-steps/s"
FALLBACK_HBM_GBS = 6650.0          # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
# longest possible game in steps (SPEC.md): werewolf 11 + 9*(P-3) (one death per day down to 2 players), TTL 2 + 8*P


def game_cap(family: int, players: int, max_revotes: int = 0) -> int:
    # longest possible game: werewolf 11 + 9*(P-3) steps (+2 per re-vote, per day), TTL 2 + 8*P
    return 9 * players - 16 + 2 * max_revotes * (players - 2) if family == 1 else 2 + 8 * players


# BASELINE.json configs[1..4] ("config 2..5" in BASELINE.md section 4).  `total` sessions are sharded over the ranks; each
# rank keeps a ring of R independent replicas of its shard (different session ids) so that its inputs come from HBM.
CONFIGS = {
    # (ring of 16 batches, one stream each: a short timed run (K = 20) is then 320 launches / 2.2 ms and the streams that run
    # out of work early in such a window are a smaller share of the machine — 7.3e10 instead of 6.8e10 with a ring of 8 on 8
    # streams; the long-run value is the same within 1 %: 7.53e10 against 7.56e10.  16 batches on 8 streams: 7.17 / 7.48e10)
    2: dict(game="werewolf-(mafia)", players=8, total=None, sessions=1 << 20, ring=16, streams=16, ctas_per_sm=3),
    3: dict(game="werewolf-(mafia)", players=16, total=1 << 24, ring=4, streams=4, ctas_per_sm=3),
    4: dict(game="werewolf-revote", players=32, total=1 << 26, ring=1, streams=1, ctas_per_sm=0),
    5: dict(game="two-truths-and-a-lie", players=4, total=1 << 28, ring=2, streams=2, ctas_per_sm=0),
}


def apply_config(a, world: int):
    """--config N presets game / players / sessions per batch / ring / streams (explicit flags still win)."""
    c = CONFIGS[a.config]
    given = {k for k in ("game", "players", "sessions", "ring", "streams", "ctas_per_sm") if getattr(a, k) is not None}
    for k in ("game", "players", "ring", "streams", "ctas_per_sm"):
        if k not in given:
            setattr(a, k, c[k])
    if "sessions" not in given:
        if c["total"] is None:
            a.sessions = c["sessions"]                       # config 2: 2^20 sessions per batch on every GPU (weak scaling)
        else:
            # the shard of this rank; the ring holds R independent replicas of it so that the resident data set stays
            # larger than L2 (2^24 sessions x 64 B over 8 GPUs is 134 MB per GPU: one replica would sit in the 126 MB L2)
            a.sessions = max(1 << 16, c["total"] // max(1, world))
    return a


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.md section 4 configuration: 2 = werewolf 8 players, 2^20 sessions per batch per GPU (the headline, "
                         "default); 3 = werewolf 16 players, 2^24 sessions sharded over the GPUs; 4 = werewolf-revote 32 players, "
                         "2^26 sessions sharded; 5 = two-truths-and-a-lie, 2^28 sessions sharded")
    ap.add_argument("--steps", type=int, default=2500, help="timed steps; one step = one pass over the ring (one launch per batch)")
    ap.add_argument("--warmup", type=int, default=25)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--game", default=None)
    ap.add_argument("--players", type=int, default=None)
    ap.add_argument("--sessions", type=int, default=None, help="sessions per batch (per GPU)")
    ap.add_argument("--ring", type=int, default=None, help="batches in the ring (footprint must exceed L2)")
    ap.add_argument("--cap", type=int, default=0, help="steps before a batch is re-initialised (0 = by player count)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "tps", "tps_generic", "coop"])
    ap.add_argument("--launch", default="streams", choices=["streams", "ring"],
                    help="streams: one step launch per batch, the ring's batches spread over --streams CUDA streams; "
                         "ring: ONE launch per pass over the ring on one stream (ge_step_ring, full-occupancy grid)")
    ap.add_argument("--ring-streams", type=int, default=1, help="--launch ring: streams the ring is split over (one ring launch per stream per pass)")
    ap.add_argument("--ring-ctas", type=int, default=4, help="--launch ring with several streams: CTAs per SM of each ring launch")
    ap.add_argument("--streams", type=int, default=None, help="CUDA streams the ring's independent batches are spread over")
    ap.add_argument("--ctas-per-sm", type=int, default=None,
                    help="persistent grid of a step launch = SMs x this (0 = occupancy limit); small grids let the launches "
                         "of the ring's other batches be resident at the same time")
    ap.add_argument("--autoreset", action="store_true",
                    help="re-initialise a batch on the device as soon as all its games are over (ge_batch_set_autoreset) instead "
                         "of from the host after `cap` steps; measured: no gain at 2^20 sessions (some game always runs to the cap)")
    ap.add_argument("--compaction", default="", help="active-prefix compaction 'every,shift' (default: the library's choice for the family)")
    ap.add_argument("--regroup", default="", help="phase regrouping 'every,shift' (default: the library's choice for the table)")
    ap.add_argument("--light-bulk", action="store_true",
                    help="A/B: header-only launches fetch their tiles with cp.async.bulk + mbarrier (ge_batch_set_option)")
    ap.add_argument("--no-pdl", action="store_true", help="A/B: plain stream-ordered step launches instead of programmatic dependent launches (GE_OPT_PDL 0)")
    ap.add_argument("--store", default="packed", choices=["canonical", "packed"],
                    help="session store in HBM: packed (werewolf tables up to 8 players keep a 32-byte record in two 16-byte columns, "
                         "the library's default for them; other tables are canonical either way) or canonical columns for every table "
                         "(ge_batch_set_option GE_OPT_STORE_PACKED 0: the A/B)")
    ap.add_argument("--setup-passes", type=int, default=-1,
                    help="untimed passes over the ring BEFORE the warm-up that bring it to its steady state (every batch through whole "
                         "game cycles with fresh ids, so the compaction checks follow the learned schedule); -1 = two game cycles")
    ap.add_argument("--head-start-us", type=int, default=3000,
                    help="length of the spin kernel the timed launches are queued behind (host head start; 0 = none)")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--e2e-calls", type=int, default=24)
    ap.add_argument("--e2e-subs", type=int, default=4, help="pipelined sub-batches of the end-to-end call")
    ap.add_argument("--wire", default="dense", choices=["dense", "canonical"],
                    help="record format of the end-to-end call's host buffers (dense: 32 / 48 bytes per session for werewolf "
                         "tables up to 8 / 16 players, SPEC section 5b; other tables are canonical either way)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    return apply_config(a, int(os.environ.get("WORLD_SIZE", "1")))


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (pynvml; nvidia-smi fallback)."""

    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def run(self):
        if self._nvml is None:
            return
        nv = self._nvml
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        if not s:       # fallback: one nvidia-smi sample (outside the timed region; said so in the record)
            try:
                import subprocess
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [int(x) for x in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 0, "note": "nvidia-smi after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------- helpers
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


L2_BYTES = 126 << 20


def ncu_traffic(game: str, players: int, kernel: str, sessions: int, ring: int, record_bytes: int):
    """Physical DRAM traffic per counted step of this workload from the committed ncu captures (profiles/traffic.json,
    written by tools/phys_traffic.py), or None.  An entry is used only for the same table, player count and kernel,
    and only when its resident data set and ours are on the same side of the L2 capacity (bytes per step do not
    depend on the batch size once the ring is several times larger than L2 — cfg 5: 42.4 B at 2^24 and 42.5 B at
    2^27 sessions per batch — but a ring that fits in L2 moves almost nothing: 6.6 B at 4 x 2^19 x 64 B)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None
    exact = t.get("%s_p%d_%s_n%d" % (game, players, kernel, sessions))
    if exact:
        return exact
    mine = ring * sessions * record_bytes
    best = None
    for k, e in t.items():
        if not k.startswith("%s_p%d_%s_n" % (game, players, kernel)):
            continue
        theirs = e.get("resident_bytes", 0)
        if mine > 2 * L2_BYTES and theirs > 2 * L2_BYTES:
            if best is None or abs(theirs - mine) < abs(best.get("resident_bytes", 0) - mine):
                best = dict(e, source=e["source"] + " [measured at %d sessions per batch]" % e.get("sessions_per_batch", 0))
    return best


def build_oracle_native():
    from oracle import oracle as _o
    try:
        _o.build(native=True)
        return True
    except Exception:
        _o.build(native=False)
        return False


def host_threads(o) -> int:
    """Threads the CPU arm uses: every core this process may run on.  torchrun exports OMP_NUM_THREADS=1, which
    would otherwise throttle the baseline to one thread."""
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    return max(int(o.max_threads()), cores)


def cpu_oracle_rate(cg, n_sessions: int, cap: int, seed: int, threads: int, native: bool):
    """Times Oracle B (the CPU port) on n_sessions full games; returns (steps, seconds)."""
    from oracle.oracle import Oracle
    o = Oracle(cg.blob, native=native)
    rec = o.init(n_sessions)
    st = o.new_stats()
    t0 = time.perf_counter()
    o.step(rec, 0, seed, cap, st, threads=threads)
    return int(st[0]), time.perf_counter() - t0


def cpu_baseline(cg, cap: int, seed: int, target_seconds: float):
    """Oracle B on whole games of the same workload: best of 5 on all host threads, plus a single-thread figure
    (BASELINE.md section 4); the five runs together take about `target_seconds`."""
    native = build_oracle_native()
    from oracle.oracle import Oracle
    threads = host_threads(Oracle(cg.blob, native=native))
    steps, dt = cpu_oracle_rate(cg, 1 << 13, cap, seed, threads, native)          # calibration (also warms the thread pool)
    per_session = dt / (1 << 13)
    n = int(max(1 << 13, min(1 << 22, target_seconds / 6.0 / max(per_session, 1e-9))))
    runs = [cpu_oracle_rate(cg, n, cap, seed, threads, native) for _ in range(5)]
    steps, dt = min(runs, key=lambda r: r[1])
    n1 = max(1 << 10, n // max(1, threads))
    s1, d1 = min((cpu_oracle_rate(cg, n1, cap, seed, 1, native) for _ in range(3)), key=lambda r: r[1])
    return {
        "value": steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": "%d sessions x %d steps (whole games) of the same workload, Oracle B (oracle/ge_oracle.c, gcc -O3%s, OpenMP), best of 5 "
                  "(%.2f s; all five: %s)" % (n, cap, " -march=native" if native else "", dt, " ".join("%.2f" % r[1] for r in runs)),
        "single_thread": {"value": s1 / d1, "unit": UNIT, "sample": "%d sessions, best of 3, %.2f s" % (n1, d1)},
        "ns_per_step": 1e9 * dt / max(1, steps),
    }


# ----------------------------------------------------------------------------------------- reference arm
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from game_engine_b200 import compile_game
    from oracle.oracle import Oracle
    cg = compile_game(a.game, a.players)
    cap = a.cap or game_cap(cg.family, a.players, cg.table.max_revotes)
    native = build_oracle_native()
    o = Oracle(cg.blob, native=native)
    threads = host_threads(o)
    # size the per-step sample so that K+W steps take about a minute and a half (first call discarded: thread
    # start-up and page faults); never below 2^16 sessions per step, where OpenMP overhead would dominate
    cpu_oracle_rate(cg, 1 << 15, cap, a.seed, threads, native)
    steps, dt = cpu_oracle_rate(cg, 1 << 17, cap, a.seed, threads, native)
    per_sess_step = dt / ((1 << 17) * cap)                 # seconds per session per pass (terminal passes are cheap)
    n = int(min(a.sessions, max(1 << 16, 90.0 / max(per_sess_step * (a.steps + a.warmup), 1e-12))))
    # same steady-state ring as the CUDA arm: R sub-batches staggered through their games
    R = a.ring
    n = max(R, n // R * R)
    sub = n // R
    recs = [o.init(sub) for _ in range(R)]
    st = o.new_stats()
    age, epoch = [0] * R, [0] * R
    for i in range(R):
        pre = (i * cap) // R
        if pre:
            o.step(recs[i], i * sub, a.seed, pre, None, threads=threads)
        age[i] = pre

    def one_step():
        # one pass over the whole sample = one step of every sub-batch
        for i in range(R):
            if age[i] >= cap:
                epoch[i] += 1
                recs[i] = o.init(sub)
                age[i] = 0
            o.step(recs[i], (epoch[i] * R + i) * sub, a.seed, 1, st, threads=threads)
            age[i] += 1

    for _ in range(a.warmup):
        one_step()
    c0 = int(st[0])
    t0 = time.perf_counter()
    for _ in range(a.steps):
        one_step()
    dt = time.perf_counter() - t0
    counted = int(st[0]) - c0
    value = counted / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "%s, %d players, sample of %d sessions per step (bounded sample of the 2^20-session batch), "
                               "re-initialised every %d steps" % (a.game, a.players, n, cap),
                   "note": "the reference repo has no CPU implementation of this path that can run without an LLM; this arm "
                           "times the CPU port of the same rules (oracle/ge_oracle.c) on all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d sessions per step, %d steps, Oracle B%s" % (n, a.steps, " -march=native" if native else "")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------- our arm
def stream_delay(device: int, stream: int, microseconds: int) -> None:
    """Keeps a CUDA stream busy for about `microseconds` with a one-thread spin kernel from the bench-only helper
    library (tools/benchaux, built by game_engine_b200/build.py; deliberately not part of the product ABI)."""
    import ctypes
    path = os.path.join(ROOT, "tools", "benchaux", "libge_benchaux.so")
    if not os.path.exists(path):
        from game_engine_b200 import build as _b
        _b.build_aux()
    lib = ctypes.CDLL(path)
    lib.bx_stream_delay.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint]
    if lib.bx_stream_delay(int(device), ctypes.c_void_p(stream), int(microseconds)) != 0:
        raise RuntimeError("bx_stream_delay failed")


class _CudaArray:
    """Minimal __cuda_array_interface__ view so torch can wrap the library's device statistics buffer."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def bind_to_gpu_numa(index: int):
    """Multi-rank runs: pin this process to the CPUs nearest to its GPU (NVML's ideal CPU affinity) before anything
    is allocated, so that the pinned host buffers of the end-to-end call sit on the GPU's own NUMA node (what
    `numactl` does for a production launcher).  Returns the CPU list, or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in (64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1) if c in allowed)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def run_ours(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    # stdout carries exactly ONE JSON line: whatever libraries print there (the image sets NCCL_DEBUG=VERSION, so NCCL
    # prints a banner with printf) is sent to stderr; the result line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from game_engine_b200 import compile_game
    from game_engine_b200.batch import PinnedBuffer, SessionBatch, Table

    cg = compile_game(a.game, a.players)
    S = cg.record_size
    cap = a.cap or game_cap(cg.family, a.players, cg.table.max_revotes)
    N, R = a.sessions, a.ring
    tab = Table(cg)
    packable = cg.family == 1 and a.players <= 16
    packed = a.store == "packed" and packable and a.kernel != "coop" and cg.table.max_revotes == 0
    S_store = (32 if a.players <= 8 else 48) if packed else S                    # bytes per session resident in HBM
    # the ring's batches are independent sessions: batch i runs on stream i % NS so that one batch's launch
    # ramp / tail and near-empty late-game launches overlap with another batch's work
    merged = a.launch == "ring" and a.kernel != "coop"
    NS = max(1, min(a.ring_streams if merged else a.streams, R))
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    stream = streams[0]
    torch.cuda.set_stream(stream)

    # session ids: unique across ranks, ring slots and epochs
    def sid_base(epoch, slot):
        return ((epoch * world + rank) * R + slot) * N

    ring = [SessionBatch(tab, N, first_session_id=sid_base(0, i), seed=a.seed, device=local_rank, kernel=a.kernel) for i in range(R)]
    for i, b in enumerate(ring):
        b.set_stream(streams[i % NS].cuda_stream)
        b.set_grid((a.ring_ctas if NS > 1 else 0) if merged else a.ctas_per_sm)
        if a.light_bulk:
            b.set_option("light_bulk", 1)
        if packable:
            b.set_option("store_packed", 1 if packed else 0)
        if a.no_pdl:
            b.set_option("pdl", 0)
        if a.regroup:
            b.set_regroup(*[int(x) for x in a.regroup.split(",")])
        if a.compaction:
            b.set_compaction(*[int(x) for x in a.compaction.split(",")])
    # Steady state: a batch whose games are all over starts over with fresh session ids — from the host after `cap`
    # steps (the longest possible game; default) or, with --autoreset, on the device as soon as the periodic
    # compaction check finds no live session (ge_batch_set_autoreset).
    auto = a.autoreset and a.kernel != "coop"
    if auto:
        for b in ring:
            b.set_autoreset(world * R * N)           # sid_base(epoch + 1, slot) - sid_base(epoch, slot)
    host_cap = (1 << 60) if auto else cap
    age = [0] * R
    epoch = [0] * R
    for i, b in enumerate(ring):                     # stagger: batch i starts i*cap/R steps into its games
        pre = (i * cap) // R
        if pre:
            b.step(pre)
        age[i] = pre
    from game_engine_b200.batch import step_many, step_ring
    k_global = 0
    resets = 0

    def ring_groups(batches):
        """--launch ring with several streams: the batches of one stream form one ring launch"""
        groups = {}
        for b in batches:
            groups.setdefault(ring.index(b) % NS, []).append(b)
        return list(groups.values())

    def step_each(batches):
        """one step of each batch of the list: a launch per batch, or (--launch ring) one launch per stream's share of the list"""
        if merged and batches:
            for g in ring_groups(batches):
                step_ring(g, 1)
        else:
            for b in batches:
                b.step(1)

    def run_steps(n):
        """n launches, round-robin over the ring starting at slot k_global % R; a batch that has been stepped
        `cap` times (the longest possible game) is first re-initialised with fresh session ids.  Whole rounds
        that need no re-initialisation go to the library in one ge_step_many call (keeps Python out of the way)."""
        nonlocal k_global, resets
        while n > 0:
            i = k_global % R
            if age[i] >= host_cap:
                epoch[i] += 1
                ring[i].reset(first_session_id=sid_base(epoch[i], i))
                age[i] = 0
                resets += 1
            run = 0                                   # launches until some batch hits the cap, walking from slot i
            while run < n and age[(i + run) % R] + (run // R) < host_cap:
                run += 1
            run = max(run, 1)
            head = min(run, (R - i) % R)              # finish the current round first
            step_each([ring[(i + j) % R] for j in range(head)])
            full, tail = divmod(run - head, R)
            if full and merged:
                for _ in range(full):
                    for g in ring_groups(ring):
                        step_ring(g, 1)
            elif full:
                step_many(ring, full)
            step_each(ring[:tail])
            for j in range(run):
                age[(i + j) % R] += 1
            k_global += run
            n -= run

    # steady state first (untimed, not part of the warm-up count): two whole game cycles per batch, after which every batch
    # has been re-initialised at least once and the host has seen which compaction checks fire (learned schedule).  A
    # short timed run (K = 20) otherwise measures the ring's first, unrepresentative cycle: -8 %.
    setup_passes = (2 * cap if a.kernel != "coop" else 0) if a.setup_passes < 0 else a.setup_passes
    if setup_passes:
        run_steps(setup_passes * R)
        torch.cuda.synchronize()
    run_steps(max(3, a.warmup) * R)
    torch.cuda.synchronize()
    agg = torch.zeros(560, dtype=torch.int64, device=dev)
    snap_pin = PinnedBuffer(8 * R)                           # counter snapshots at the start of the timed region
    snap = snap_pin.array.view(np.uint64)
    epochs0 = sum(b.epochs() for b in ring) if auto else 0

    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # The timed launches are queued BEHIND a short spin kernel on stream 0 (every other stream waits for ev0, which
    # follows it), so that the host is hundreds of launches ahead when the device starts the timed region — as it is
    # in steady state — instead of feeding an empty GPU one launch at a time: without it a short run (small K)
    # measures the host's start-up, not the device.  The spin itself is outside [ev0, ev1].
    for i, b in enumerate(ring):
        b.counted_steps_async(snap[i:i + 1])
    launches0 = sum(b.launch_count() for b in ring)
    resets0 = resets
    stream_delay(local_rank, stream.cuda_stream, a.head_start_us)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for st in streams[1:]:
        st.wait_event(ev0)
    t_enq0 = time.perf_counter()
    run_steps(a.steps * R)                                   # one step = one pass over the ring = R launches
    enqueue_ms = (time.perf_counter() - t_enq0) * 1e3        # host time to enqueue K steps
    for st in streams[1:]:                                   # join: ev1 fires when the work of ALL streams is done
        e = torch.cuda.Event()
        e.record(st)
        stream.wait_event(e)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)

    counted = sum(b.counted_steps() for b in ring) - int(snap[:R].sum())
    launches = sum(b.launch_count() for b in ring) - launches0
    resets = (sum(b.epochs() for b in ring) - epochs0) if auto else resets - resets0

    # the job's only exchange step: statistics all-reduce (win rate + phase-length histogram) over NCCL
    t_ar0 = time.perf_counter()
    for b in ring:
        b.set_stream(stream.cuda_stream)
        b.stats_refresh()
        agg += torch.as_tensor(_CudaArray(b.stats_device_ptr(), 560), device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    allreduce_ms = (time.perf_counter() - t_ar0) * 1e3
    stats = agg.cpu().numpy()

    tm = torch.tensor([ms, float(counted), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = tm.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tm.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_max, counted_all, launches_all = float(tmax[0]), float(tsum[1]), int(tsum[2])
    else:
        ms_max, counted_all, launches_all = ms, float(counted), int(launches)
    value = counted_all / (ms_max * 1e-3)

    # ---- fused mode, reported separately (never the roofline number): whole games of one fresh 2^20-session batch per
    #      launch, state in registers across the steps (ge_run_fused) — what a run-to-completion user gets on the device
    fused = None
    if not a.no_e2e:
        fb = [SessionBatch(tab, N, first_session_id=sid_base((1 << 19) + j, 0), seed=a.seed, device=local_rank, kernel=a.kernel)
              for j in range(4)]
        for j, b in enumerate(fb):
            b.set_stream(streams[j % NS].cuda_stream)
            if packable:
                b.set_option("store_packed", 1 if packed else 0)
            b.run_fused(cap)
        torch.cuda.synchronize()
        f0 = sum(b.counted_steps() for b in fb)
        for b in fb:
            b.reset()
        torch.cuda.synchronize()
        fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fe0.record(stream)
        for st in streams[1:]:
            st.wait_event(fe0)
        for b in fb:
            b.run_fused(cap)
        for st in streams[1:]:
            e = torch.cuda.Event()
            e.record(st)
            stream.wait_event(e)
        fe1.record(stream)
        torch.cuda.synchronize()
        f_counted = sum(b.counted_steps() for b in fb) - f0
        fused = {"value": f_counted / (fe0.elapsed_time(fe1) * 1e-3), "unit": UNIT,
                 "note": "ge_run_fused: %d steps per launch, state in registers; 4 fresh batches of %d sessions; per GPU, "
                         "not part of value / roofline" % (cap, N)}
        for b in fb:
            b.close()

    # ---- e2e: the public host-buffer call (H2D of the initial records, `cap` steps, D2H of the final
    #      records + statistics), pinned host memory, wall clock around synchronous calls
    e2e = None
    if not a.no_e2e:
        # The public host-buffer call on the same 2^20 sessions, split into NSUB sub-batches driven with
        # run_host_async so that H2D, the steps and D2H of different sub-batches overlap (two copy engines + SMs).
        NSUB = max(1, a.e2e_subs)
        sub = min(N, 1 << 20) // NSUB
        NE = sub * NSUB                                       # sessions per end-to-end call (bounded: 2 x NE x W bytes of pinned memory)
        subs = [SessionBatch(tab, sub, first_session_id=sid_base(1 << 20, 0) + j * sub, seed=a.seed, device=local_rank, kernel=a.kernel)
                for j in range(NSUB)]
        for sb in subs:
            sb.set_wire(a.wire)
            if packable:
                sb.set_option("store_packed", 1 if packed else 0)
        W = subs[0].wire_record_size                          # bytes per session on the wire
        pin_in, pin_out, pin_st = PinnedBuffer(NE * W), PinnedBuffer(NE * W), PinnedBuffer(NSUB * 560 * 8)
        rin = pin_in.array.reshape(NSUB, sub, W)
        rout = pin_out.array.reshape(NSUB, sub, W)
        rst = pin_st.array.view(np.uint64).reshape(NSUB, 560)
        for j, sb in enumerate(subs):
            sb.export_state(out=rin[j])                       # initial records in the wire format, produced by the library
            sb.set_host_fused(True)                           # run-to-completion call: one fused launch per sub-batch

        def e2e_stream(n_calls, n_steps, src, records_out=True):
            """n_calls back-to-back calls; call c+1 of a sub-batch is enqueued behind call c on the same stream, so the
            H2D of one call overlaps the D2H of the previous one (a server draining a queue of requests).  Every call
            moves its inputs host->device and its results device->host; statistics accumulate across the calls."""
            for sb in subs:
                sb.clear_stats()
            for c in range(n_calls):
                for j, sb in enumerate(subs):
                    sb.run_host_async(src[j], rout[j] if records_out else None, n_steps, rst[j])
            for sb in subs:
                sb.sync()
            return int(rst[:, 0].sum())

        e2e_stream(2, cap, rin)                               # warm-up
        if world > 1:
            dist.barrier()
        # three repeats of the timed stream of calls, the MEDIAN is reported (host-side jitter moves a 20 ms measurement by
        # +-7 % from one repeat to the next on the pool's boxes); all three are listed in e2e.repeats
        reps = []
        for _ in range(3):
            t0 = time.perf_counter()
            e_counted = e2e_stream(a.e2e_calls, cap, rin)
            reps.append((time.perf_counter() - t0, e_counted))
        dt, e_counted = sorted(reps)[1]
        tl = time.perf_counter()
        e2e_stream(1, cap, rin)                               # one isolated call: latency, pipeline fill and drain exposed
        lat_ms = (time.perf_counter() - tl) * 1e3
        # statistics-only output (a Monte-Carlo caller that wants win rates, not final records): records in, 4.5 KB out
        e2e_stream(2, cap, rin, records_out=False)
        ts = time.perf_counter()
        so_counted = e2e_stream(a.e2e_calls, cap, rin, records_out=False)
        dts = time.perf_counter() - ts
        # single-step variant: every session-phase-step round-trips through host memory
        e2e_stream(1, 1, rin)
        t1 = time.perf_counter()
        s_counted = e2e_stream(a.e2e_calls, 1, rout)
        dt1 = time.perf_counter() - t1
        ed = torch.tensor([dt, float(e_counted), dt1, float(s_counted), dts, float(so_counted)], dtype=torch.float64, device=dev)
        if world > 1:
            emax = ed.clone()
            dist.all_reduce(emax, op=dist.ReduceOp.MAX)
            esum = ed.clone()
            dist.all_reduce(esum, op=dist.ReduceOp.SUM)
            dt, e_counted, dt1, s_counted = float(emax[0]), float(esum[1]), float(emax[2]), float(esum[3])
            dts, so_counted = float(emax[4]), float(esum[5])
        e2e = {
            "value": e_counted / dt, "unit": UNIT,
            "h2d_bytes_per_step": NE * W, "d2h_bytes_per_step": NE * W + NSUB * 560 * 8,
            "wire": {"format": a.wire, "record_bytes": W, "canonical_record_bytes": S},
            "call": "%d back-to-back calls, each %d x SessionBatch.run_host_async (ge_run_host_async): pinned host records in -> "
                    "%d steps -> records + stats out for %d sessions in %d pipelined sub-batches; one sync at the end; bytes "
                    "are per call" % (a.e2e_calls, NSUB, cap, NE, NSUB),
            "calls": a.e2e_calls, "ms_per_call": dt / a.e2e_calls * 1e3, "single_call_latency_ms": lat_ms,
            "repeats": {"values": [c / t for t, c in reps], "reported": "median of 3 (this rank)"},
            "statistics_only_output": {"value": so_counted / dts, "unit": UNIT, "ms_per_call": dts / a.e2e_calls * 1e3,
                                       "d2h_bytes_per_step": NSUB * 560 * 8,
                                       "note": "the same call with records_out = NULL: records in, the statistics (win rates, histograms) out"},
            "single_step_round_trip": {"value": s_counted / dt1, "unit": UNIT, "ms_per_call": dt1 / a.e2e_calls * 1e3,
                                       "note": "n_steps=1 per call: every step crosses PCIe twice"},
        }
        for sb in subs:
            sb.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_src = hbm_peak()
    B = 2 * S
    per_gpu = counted_all / world / (ms_max * 1e-3)                       # counted steps per second per GPU
    achieved = per_gpu * B / 1e9
    kern = ring[0].kernel
    step_launches = a.steps * (NS if merged else R)                       # step-kernel launches per GPU in the timed region
    # physical DRAM traffic: bytes per counted step from the committed ncu capture of THIS workload (same table, players,
    # kernel and batch size; anything else is refused), times the live step rate
    tr = ncu_traffic(a.game, a.players, kern + ("+packed" if packed else ""), N, R, S_store)
    nec = tab.necessary_bytes_per_step(stats, packed)                             # visit-weighted columns that must move (ge_table_phase_io)
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (tr["bytes_per_step"] * counted_all / world / step_launches) if tr else None, "peak_source": peak_src,
            "frac_of_nominal_8000": achieved / 8000.0,         # BASELINE.md section 2 asks for the nominal figure too
            "algorithmic_bytes_per_step": B, "kernel": "k_%s_%s_%s" % ("ring" if merged else "step", "w" if cg.family == 1 else "t", "tps" if kern.startswith("tps") else kern),
            "steps_per_launch": counted_all / world / step_launches, "launches_per_step": NS if merged else R,
            "necessary_bytes_per_step": nec, "necessary_gbs": per_gpu * nec / 1e9, "frac_necessary": per_gpu * nec / 1e9 / peak,
            "note": "achieved = counted steps x 2S (SURVEY 8d: the whole record read and written every step); necessary = the columns "
                    "each phase must move (ge_table_phase_io) weighted by the visit histogram; dram_physical = what crossed the "
                    "DRAM pins (ncu, reads AND writes, caches not flushed between kernels)"}
    if tr:
        roof["dram_physical"] = {"bytes_per_step": tr["bytes_per_step"], "read_bytes_per_step": tr["read_bytes_per_step"],
                                 "write_bytes_per_step": tr["write_bytes_per_step"], "gbs": per_gpu * tr["bytes_per_step"] / 1e9,
                                 "frac_of_measured_peak": per_gpu * tr["bytes_per_step"] / 1e9 / peak,
                                 "frac_of_nominal_8000": per_gpu * tr["bytes_per_step"] / 1e9 / 8000.0,
                                 "issue_frac": per_gpu * tr["warp_instructions_per_step"] / (148 * 4 * 1.965e9),
                                 "source": tr["source"]}
    else:
        roof["dram_physical"] = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {
            "workload": "%s.yaml, %d players, %d sessions per batch per GPU, Philox bots" % (a.game, a.players, N),
            "baseline_config": a.config,
            "light_path": "cp.async.bulk + mbarrier" if a.light_bulk else "LDG.128", "kernel": kern, "launch": ("%d ring launch(es) per pass (ge_step_ring)" % NS) if merged else "one launch per batch", "streams": NS,
            "ctas_per_sm": "occupancy limit" if merged else a.ctas_per_sm, "ring_batches": R, "ring_bytes": R * N * S_store, "store": ("packed (%d bytes per session in HBM)" % S_store) if packed else "canonical columns", "l2_policy": "inputs larger than L2 (ring of batches, round-robin)",
            "steps_before_reinit": "when every game of the batch is over (device-side auto-reset, checked every 8 steps)" if auto else cap,
            "reinits_in_timed_region": resets, "setup_passes": setup_passes, "record_bytes": S, "seed": a.seed,
            "parallelism": "dp%d (independent session shards, one NCCL all-reduce of the statistics)" % world,
        },
        "roofline": roof,
        "e2e": e2e,
        "fused": fused,
        "gpu_launches": launches_all,
        "clocks": clocks,
        "stats_allreduce_ms": allreduce_ms,
        "host_enqueue_ms": enqueue_ms,
        "cpu_binding": ("%d CPUs near the GPU (NVML affinity)" % len(numa_cpus)) if numa_cpus else "none",
        "win_rate": {"villagers": float(stats[2]) / max(1.0, float(stats[2] + stats[3])),
                     "werewolves": float(stats[3]) / max(1.0, float(stats[2] + stats[3])),
                     "sessions_finished": int(stats[2] + stats[3])} if cg.family == 1 else None,
    }
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cg, cap, a.seed, a.cpu_seconds)
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
