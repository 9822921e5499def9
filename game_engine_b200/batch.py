"""SessionBatch — host handle over N independent game sessions resident in HBM.

Batch-level mirror of the reference's per-room graph run (reference agent/game_agent_v2.py:1571-1587:
InitialRouterNode -> BotBehaviorNode -> PhaseNode -> RefereeNode): `step()` is one such run for every
non-terminal session.  Everything numeric happens in libgame_engine_b200.so; NumPy arrays here are only
host buffers handed across the C ABI.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import capi
from .compiler import CompiledGame


class Table:
    """Owns a ge_table handle."""

    def __init__(self, game: CompiledGame):
        self.game = game
        L = capi.lib()
        self._h = ctypes.c_void_p()
        self._blob = ctypes.create_string_buffer(game.blob, len(game.blob))
        capi.check(L.ge_table_create(ctypes.cast(self._blob, ctypes.c_void_p), len(game.blob), ctypes.byref(self._h)))
        self.record_size = int(L.ge_table_record_size(self._h))
        assert self.record_size == game.record_size

    def phase_io(self, packed: bool = False):
        """[(read_bytes, write_bytes)] per phase index: what a step starting there must move per session (packed: with
        the packed session store, GE_OPT_STORE_PACKED)."""
        out = []
        fn = capi.lib().ge_table_phase_io_packed if packed else capi.lib().ge_table_phase_io
        for i in range(len(self.game.phase_ids)):
            r, w = ctypes.c_uint32(), ctypes.c_uint32()
            capi.check(fn(self._h, i, ctypes.byref(r), ctypes.byref(w)))
            out.append((int(r.value), int(w.value)))
        return out

    def necessary_bytes_per_step(self, stats, packed: bool = False) -> float:
        """Visit-weighted necessary DRAM bytes per session-phase-step (stats = the u64[560] statistics: words
        260.. are visits per phase index; a visit to a non-terminal phase is followed by one step that starts there)."""
        io = self.phase_io(packed)
        num = den = 0.0
        for i, (r, w) in enumerate(io):
            if r + w == 0:
                continue
            v = float(stats[260 + i])
            num += v * (r + w)
            den += v
        return num / den if den else 0.0

    def close(self) -> None:
        if self._h:
            capi.lib().ge_table_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SessionBatch:
    def __init__(self, table: Table, n_sessions: int, first_session_id: int = 0, seed: int = 0, device: int = 0,
                 kernel: str = "auto"):
        self.table = table
        self.n = int(n_sessions)
        self.first_session_id = int(first_session_id)
        self.seed = int(seed)
        self.device = int(device)
        L = capi.lib()
        self._h = ctypes.c_void_p()
        capi.check(L.ge_batch_create(table._h, device, self.n, self.first_session_id, self.seed, ctypes.byref(self._h)))
        self.set_kernel(kernel)

    # ---- control
    def set_stream(self, stream: int) -> None:
        """Bind to an external CUDA stream handle (e.g. torch.cuda.Stream().cuda_stream); 0 = own stream."""
        capi.check(capi.lib().ge_batch_set_stream(self._h, ctypes.c_void_p(stream)))

    def set_wire(self, wire: str) -> None:
        """Record format of the host-buffer calls: "canonical" (SPEC section 5) or "dense" (section 5b: 32 / 48 bytes
        for werewolf tables up to 8 / 16 players; other tables keep the canonical record)."""
        capi.check(capi.lib().ge_batch_set_wire(self._h, capi.WIRE_NAMES[wire]))

    @property
    def wire_record_size(self) -> int:
        return int(capi.lib().ge_batch_wire_size(self._h))

    @property
    def human_stride(self) -> int:
        return int(capi.lib().ge_table_human_stride(self.table._h))

    def set_human_seats(self, masks: Optional[np.ndarray]) -> None:
        """Seats played by people: uint32[n], bit p-1 = player p (None = all bots).  SPEC.md D3h."""
        if masks is None:
            capi.check(capi.lib().ge_batch_set_human_seats(self._h, None))
            return
        m = np.ascontiguousarray(masks, dtype=np.uint32)
        assert m.size == self.n
        capi.check(capi.lib().ge_batch_set_human_seats(self._h, m.ctypes.data))

    def set_human_choices(self, choices: np.ndarray) -> None:
        """Inputs of the human seats for the NEXT step: uint8[n, human_stride], 0xFF = has not acted."""
        c = np.ascontiguousarray(choices, dtype=np.uint8)
        assert c.size == self.n * self.human_stride
        self._hc_keepalive = c                      # the copy is asynchronous
        capi.check(capi.lib().ge_batch_set_human_choices(self._h, c.ctypes.data))

    def set_kernel(self, kernel: str) -> None:
        capi.check(capi.lib().ge_batch_set_kernel(self._h, capi.KERNEL_NAMES[kernel]))

    @property
    def kernel(self) -> str:
        k = capi.lib().ge_batch_get_kernel(self._h)
        return {capi.KERNEL_COOP: "coop", capi.KERNEL_TPS: "tps", capi.KERNEL_TPS_GENERIC: "tps_generic"}[k]

    def reset(self, first_session_id: Optional[int] = None, seed: Optional[int] = None) -> None:
        if first_session_id is not None:
            self.first_session_id = int(first_session_id)
        if seed is not None:
            self.seed = int(seed)
        capi.check(capi.lib().ge_batch_reset(self._h, self.first_session_id, self.seed))

    def set_compaction(self, every_n_steps: int, min_dead_shift: int = 2) -> None:
        """Active-prefix compaction: check every n steps, compact when >= 1/2^shift of the prefix is dead
        (0 steps = off; default (5, 2) for the werewolf family, off for TTL)."""
        capi.check(capi.lib().ge_batch_set_compaction(self._h, int(every_n_steps), int(min_dead_shift)))

    def set_option(self, option: str, value: int) -> None:
        """Tuning options: "light_bulk" (header-only launches fetch their tiles with cp.async.bulk + mbarrier);
        "store_packed" (werewolf tables up to 8 players: 32-byte records in HBM, two 16-byte columns)."""
        capi.check(capi.lib().ge_batch_set_option(self._h, {"light_bulk": capi.OPT_LIGHT_BULK, "store_packed": capi.OPT_STORE_PACKED, "pdl": capi.OPT_PDL}[option], int(value)))

    def set_grid(self, ctas_per_sm: int) -> None:
        """Persistent grid of the step launches = SMs x ctas_per_sm (0 = as many as fit).  Smaller grids let the
        launches of other batches on other streams co-reside."""
        capi.check(capi.lib().ge_batch_set_grid(self._h, int(ctas_per_sm)))

    def set_autoreset(self, sid_stride: int) -> None:
        """Continuous simulation: when every game of the batch is over it restarts on the device with session ids
        first_session_id + epoch * sid_stride + i (0 = off; needs compaction or regrouping on)."""
        capi.check(capi.lib().ge_batch_set_autoreset(self._h, int(sid_stride)))

    def epochs(self) -> int:
        """Device-side re-initialisations so far (synchronises)."""
        v = ctypes.c_uint64()
        capi.check(capi.lib().ge_batch_epochs(self._h, ctypes.byref(v)))
        return int(v.value)

    def set_regroup(self, every_n_steps: int, min_mixed_shift: int = 3) -> None:
        """Phase regrouping: check every n steps, counting-sort the active prefix by phase when >= 1/2^shift of
        the tiles are mixed (0 steps = off; on by default for tables with a tie -> re-vote loop)."""
        capi.check(capi.lib().ge_batch_set_regroup(self._h, int(every_n_steps), int(min_mixed_shift)))

    def active(self) -> int:
        """Length of the slot prefix that can still hold live sessions (synchronises)."""
        v = ctypes.c_uint64()
        capi.check(capi.lib().ge_batch_active(self._h, ctypes.byref(v)))
        return int(v.value)

    def active_hint(self) -> int:
        """Non-blocking upper bound of active() (refreshed after each compaction); 0 = every game is over."""
        v = ctypes.c_uint64()
        capi.check(capi.lib().ge_batch_active_hint(self._h, ctypes.byref(v)))
        return int(v.value)

    def clear_stats(self) -> None:
        capi.check(capi.lib().ge_batch_clear_stats(self._h))

    def step(self, n_steps: int = 1, stream: int = 0) -> None:
        """n_steps single-step launches (asynchronous)."""
        capi.check(capi.lib().ge_step(self._h, int(n_steps), ctypes.c_void_p(stream)))

    def run_fused(self, n_steps: int, stream: int = 0) -> None:
        capi.check(capi.lib().ge_run_fused(self._h, int(n_steps), ctypes.c_void_p(stream)))

    def sync(self) -> None:
        capi.check(capi.lib().ge_sync(self._h))

    # ---- state
    def export_state(self, first: int = 0, count: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        count = self.n - first if count is None else int(count)
        S = self.wire_record_size
        if out is None:
            out = np.empty((count, S), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size == count * S
        capi.check(capi.lib().ge_export_state(self._h, int(first), count, out.ctypes.data))
        return out

    def import_state(self, records: np.ndarray, first: int = 0) -> None:
        rec = np.ascontiguousarray(records, dtype=np.uint8)
        S = self.wire_record_size
        assert rec.size % S == 0
        capi.check(capi.lib().ge_import_state(self._h, int(first), rec.size // S, rec.ctypes.data))

    def trace(self, n_steps: int, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Steps the batch n_steps times and returns uint8[n_steps + 1, count, S]: the canonical records of the
        window after every step (frame 0 = before the first step); see trace.py for the JSON form."""
        count = self.n - first if count is None else int(count)
        out = np.empty((int(n_steps) + 1, count, self.table.record_size), dtype=np.uint8)
        capi.check(capi.lib().ge_trace(self._h, int(first), count, int(n_steps), out.ctypes.data))
        return out

    def run_host(self, records_in: Optional[np.ndarray], records_out: Optional[np.ndarray], n_steps: int,
                 stats_out: Optional[np.ndarray] = None) -> None:
        """End-to-end call with host buffers (H2D, n_steps steps, D2H); synchronous."""
        pin, pout, pst = self._host_args(records_in, records_out, stats_out)
        capi.check(capi.lib().ge_run_host(self._h, pin, pout, int(n_steps), pst))

    def _host_args(self, records_in, records_out, stats_out):
        """Pointers of the host-buffer call.  The library reads / writes n * S bytes behind them, so anything but a
        C-contiguous uint8 array of exactly that size is refused here."""
        need = self.n * self.wire_record_size
        for name, arr in (("records_in", records_in), ("records_out", records_out)):
            if arr is not None and not (isinstance(arr, np.ndarray) and arr.dtype == np.uint8 and arr.flags.c_contiguous and arr.size == need):
                raise ValueError("%s must be a C-contiguous uint8 array of %d bytes (%d sessions x %d)" % (name, need, self.n, self.wire_record_size))
        if records_out is not None and not records_out.flags.writeable:
            raise ValueError("records_out must be writeable")
        if stats_out is not None and not (isinstance(stats_out, np.ndarray) and stats_out.dtype == np.uint64 and stats_out.flags.c_contiguous
                                          and stats_out.size >= capi.STATS_LEN):
            raise ValueError("stats_out must be a C-contiguous uint64 array of at least %d words" % capi.STATS_LEN)
        return (records_in.ctypes.data if records_in is not None else None,
                records_out.ctypes.data if records_out is not None else None,
                stats_out.ctypes.data if stats_out is not None else None)

    def set_host_fused(self, on: bool) -> None:
        """Host-buffer calls apply their n_steps in one fused launch (state in registers) when on."""
        capi.check(capi.lib().ge_batch_set_host_fused(self._h, 1 if on else 0))

    def run_host_async(self, records_in: Optional[np.ndarray], records_out: Optional[np.ndarray], n_steps: int,
                       stats_out: Optional[np.ndarray] = None) -> None:
        """run_host without the final synchronisation (pinned buffers; call sync() before reading them)."""
        pin, pout, pst = self._host_args(records_in, records_out, stats_out)
        capi.check(capi.lib().ge_run_host_async(self._h, pin, pout, int(n_steps), pst))

    def eval_preds(self, preds, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Lane masks [count, len(preds)] of DNF predicates (tuples (pos0, neg0, pos1, neg1)), e.g. the compiled
        audience groups of a game (CompiledGame.audience_preds)."""
        count = self.n - first if count is None else int(count)
        arr = np.ascontiguousarray(np.array(list(preds), dtype=np.uint16).reshape(-1, 4))
        out = np.zeros((count, arr.shape[0]), dtype=np.uint32)
        capi.check(capi.lib().ge_eval_preds(self._h, arr.ctypes.data, arr.shape[0], int(first), count, out.ctypes.data))
        return out

    def audience_masks(self, first: int = 0, count: Optional[int] = None) -> dict:
        """{group name: uint32[count]} for every audience group of the game's declaration."""
        aud = self.table.game.audience_chains
        if not aud:
            return {}
        flat, span = [], {}
        for name, chain in aud.items():                      # a long criterion is a run of records: OR its columns
            span[name] = (len(flat), len(chain))
            flat += chain
        out = {}
        for lo in range(0, len(flat), 32):                   # the library takes up to 32 records per call
            m = self.eval_preds(flat[lo:lo + 32], first, count)
            for name, (a, n) in span.items():
                for j in range(max(a, lo), min(a + n, lo + 32)):
                    out[name] = out.get(name, 0) | m[:, j - lo]
        return {name: np.asarray(out[name], dtype=np.uint32).copy() for name in aud}

    # ---- statistics
    def stats(self) -> np.ndarray:
        out = np.zeros(capi.STATS_LEN, dtype=np.uint64)
        capi.check(capi.lib().ge_stats(self._h, out.ctypes.data, capi.STATS_LEN))
        return out

    def stats_refresh(self, stream: int = 0) -> None:
        capi.check(capi.lib().ge_stats_refresh(self._h, ctypes.c_void_p(stream)))

    def stats_device_ptr(self) -> int:
        return int(capi.lib().ge_stats_device_ptr(self._h) or 0)

    def counted_steps(self) -> int:
        v = ctypes.c_uint64()
        capi.check(capi.lib().ge_counted_steps(self._h, ctypes.byref(v)))
        return int(v.value)

    def counted_steps_async(self, pinned_u64: np.ndarray) -> None:
        """Enqueues a copy of the counted-steps word (as of this point of the batch's stream) into a one-element
        uint64 view of a PinnedBuffer; read it after sync()."""
        assert pinned_u64.dtype == np.uint64 and pinned_u64.size == 1
        capi.check(capi.lib().ge_counted_steps_async(self._h, pinned_u64.ctypes.data))

    def launch_count(self) -> int:
        return int(capi.lib().ge_launch_count(self._h))

    def state_device_ptr(self) -> int:
        return int(capi.lib().ge_state_device_ptr(self._h) or 0)

    def state_device_bytes(self) -> int:
        return int(capi.lib().ge_state_device_bytes(self._h))

    def close(self) -> None:
        if self._h:
            capi.lib().ge_batch_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def step_many(batches, n_rounds: int = 1) -> None:
    """n_rounds round-robin passes: one step of every batch in order (one C call; see ge_step_many)."""
    arr = (ctypes.c_void_p * len(batches))(*[b._h for b in batches])
    capi.check(capi.lib().ge_step_many(arr, len(batches), int(n_rounds)))


def step_ring(batches, n_rounds: int = 1) -> None:
    """n_rounds passes over a ring of batches, ONE launch per pass (ge_step_ring): same results as step_many for
    batches that share table, device, seed, kernel and stream."""
    arr = (ctypes.c_void_p * len(batches))(*[b._h for b in batches])
    capi.check(capi.lib().ge_step_ring(arr, len(batches), int(n_rounds)))


class PinnedBuffer:
    """Page-locked host memory from the library (cudaHostAlloc) exposed as a NumPy array."""

    def __init__(self, nbytes: int):
        self._p = ctypes.c_void_p()
        capi.check(capi.lib().ge_host_alloc(ctypes.byref(self._p), int(nbytes)))
        self.nbytes = int(nbytes)
        self.array = np.ctypeslib.as_array(ctypes.cast(self._p, ctypes.POINTER(ctypes.c_uint8)), shape=(self.nbytes,))

    def close(self) -> None:
        if self._p:
            self.array = None
            capi.lib().ge_host_free(self._p)
            self._p = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
