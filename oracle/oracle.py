"""ctypes wrapper of Oracle B (oracle/ge_oracle.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
STATS_LEN = 560


def build(native: bool = False, quiet: bool = True) -> str:
    """Compile ge_oracle.c with gcc; returns the path of the shared library."""
    out = os.path.join(_HERE, "_native" if native else "", "libge_oracle.so")
    src = os.path.join(_HERE, "ge_oracle.c")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cmd = ["make", "-C", _HERE] + (["NATIVE=1"] if native else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL if quiet else None)
    return out


class Oracle:
    def __init__(self, blob: bytes, native: bool = False):
        path = os.path.join(_HERE, "_native" if native else "", "libge_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "ge_oracle.c")):
            path = build(native)
        self.lib = ctypes.CDLL(path)
        L = self.lib
        u8p, u64, sz = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t
        L.ge_cpu_table_check.argtypes = [u8p, sz]
        L.ge_cpu_record_size.argtypes = [u8p, sz]
        L.ge_cpu_record_size.restype = sz
        L.ge_cpu_init.argtypes = [u8p, sz, u8p, u64]
        L.ge_cpu_step.argtypes = [u8p, sz, u8p, u64, u64, u64, ctypes.c_int, u8p, ctypes.c_int]
        L.ge_cpu_step_h.argtypes = [u8p, sz, u8p, u64, u64, u64, ctypes.c_int, u8p, ctypes.c_int, u8p, u8p, sz]
        L.ge_cpu_stats_final.argtypes = [u8p, sz, u8p, u64, u8p]
        L.ge_cpu_peek_choices.argtypes = [u8p, sz, u8p, u64, u64, u8p]
        L.ge_cpu_philox.argtypes = [u8p, u8p, u8p]
        L.ge_cpu_validate_records.argtypes = [u8p, sz, u8p, u64, u8p]
        L.ge_cpu_validate_records.restype = ctypes.c_long
        L.ge_cpu_eval_preds.argtypes = [u8p, sz, u8p, u64, u8p, ctypes.c_int, u8p]
        self.blob = bytes(blob)
        self._blob_buf = ctypes.create_string_buffer(self.blob, len(self.blob))
        self._bp = ctypes.cast(self._blob_buf, ctypes.c_void_p)
        if L.ge_cpu_table_check(self._bp, len(self.blob)) != 0:
            raise ValueError("oracle rejected the table blob")
        self.record_size = int(L.ge_cpu_record_size(self._bp, len(self.blob)))
        self.n_players = self.blob[8]

    def max_threads(self) -> int:
        return int(self.lib.ge_cpu_max_threads())

    def init(self, n: int) -> np.ndarray:
        rec = np.zeros((n, self.record_size), dtype=np.uint8)
        self.lib.ge_cpu_init(self._bp, len(self.blob), rec.ctypes.data, n)
        return rec

    def step(self, rec: np.ndarray, first_sid: int, seed: int, n_steps: int = 1, stats: np.ndarray | None = None,
             threads: int = 0) -> None:
        assert rec.dtype == np.uint8 and rec.flags.c_contiguous and rec.shape[1] == self.record_size
        sp = stats.ctypes.data if stats is not None else None
        rc = self.lib.ge_cpu_step(self._bp, len(self.blob), rec.ctypes.data, rec.shape[0], first_sid, seed, n_steps, sp, threads)
        if rc != 0:
            raise RuntimeError("ge_cpu_step failed")

    def step_humans(self, rec: np.ndarray, first_sid: int, seed: int, masks: np.ndarray, choices: np.ndarray | None,
                    n_steps: int = 1, stats: np.ndarray | None = None) -> None:
        """step() with human seats (SPEC D3h): masks uint32[n]; choices uint8[n, stride] (0xFF = has not acted) apply
        to the first of the n_steps steps; choices None = nobody has acted."""
        assert rec.dtype == np.uint8 and rec.flags.c_contiguous and rec.shape[1] == self.record_size
        m = np.ascontiguousarray(masks, dtype=np.uint32)
        assert m.size == rec.shape[0]
        if choices is None:
            choices = np.full((rec.shape[0], 32), 0xFF, dtype=np.uint8)
        c = np.ascontiguousarray(choices, dtype=np.uint8).reshape(rec.shape[0], -1)
        sp = stats.ctypes.data if stats is not None else None
        rc = self.lib.ge_cpu_step_h(self._bp, len(self.blob), rec.ctypes.data, rec.shape[0], first_sid, seed, n_steps, sp, 0,
                                    m.ctypes.data, c.ctypes.data, c.shape[1])
        if rc != 0:
            raise RuntimeError("ge_cpu_step_h failed")

    def stats_final(self, rec: np.ndarray, stats: np.ndarray) -> None:
        self.lib.ge_cpu_stats_final(self._bp, len(self.blob), rec.ctypes.data, rec.shape[0], stats.ctypes.data)

    def peek_choices(self, record: np.ndarray, sid: int, seed: int) -> np.ndarray:
        out = np.zeros(self.n_players, dtype=np.uint8)
        r = np.ascontiguousarray(record, dtype=np.uint8)
        self.lib.ge_cpu_peek_choices(self._bp, len(self.blob), r.ctypes.data, sid, seed, out.ctypes.data)
        return out

    def eval_preds(self, rec: np.ndarray, preds) -> np.ndarray:
        arr = np.ascontiguousarray(np.array(list(preds), dtype=np.uint16).reshape(-1, 4))
        out = np.zeros((rec.shape[0], arr.shape[0]), dtype=np.uint32)
        self.lib.ge_cpu_eval_preds(self._bp, len(self.blob), rec.ctypes.data, rec.shape[0], arr.ctypes.data, arr.shape[0], out.ctypes.data)
        return out

    def validate_records(self, rec: np.ndarray) -> np.ndarray:
        """ok[i] = True when record i is well-formed (SPEC.md section 7b), the rule set of the library's import paths."""
        rec = np.ascontiguousarray(rec, dtype=np.uint8).reshape(-1, self.record_size)
        ok = np.zeros(rec.shape[0], dtype=np.uint8)
        if self.lib.ge_cpu_validate_records(self._bp, len(self.blob), rec.ctypes.data, rec.shape[0], ok.ctypes.data) < 0:
            raise RuntimeError("ge_cpu_validate_records failed")
        return ok.astype(bool)

    def philox(self, key, ctr):
        k = np.asarray(key, dtype=np.uint32)
        c = np.asarray(ctr, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        self.lib.ge_cpu_philox(k.ctypes.data, c.ctypes.data, o.ctypes.data)
        return o

    def new_stats(self) -> np.ndarray:
        return np.zeros(STATS_LEN, dtype=np.uint64)
