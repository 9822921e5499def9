"""ctypes binding of include/game_engine_b200.h (libgame_engine_b200.so).

The library is the ONLY compute path.  If it is missing or a symbol is absent, importing this module
raises — there is deliberately no Python/NumPy fallback (a silent fallback would void every parity and
performance claim).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GE_LIB selects another build of the same library (A/B experiments with build-time knobs); default: the in-tree build
LIB_PATH = os.environ.get("GE_LIB") or os.path.join(_HERE, "libgame_engine_b200.so")

STATS_LEN = 560
GE_OK, GE_ERR_ARG, GE_ERR_CUDA, GE_ERR_UNSUPPORTED, GE_ERR_NOMEM = 0, -1, -2, -3, -4
KERNEL_AUTO, KERNEL_COOP, KERNEL_TPS, KERNEL_TPS_GENERIC = 0, 1, 2, 3
OPT_LIGHT_BULK = 1
OPT_STORE_PACKED = 2
OPT_PDL = 3
WIRE_CANONICAL, WIRE_DENSE = 0, 1
WIRE_NAMES = {"canonical": WIRE_CANONICAL, "dense": WIRE_DENSE}
KERNEL_NAMES = {"auto": KERNEL_AUTO, "coop": KERNEL_COOP, "tps": KERNEL_TPS, "tps_generic": KERNEL_TPS_GENERIC}

# every symbol include/game_engine_b200.h declares: (name, restype, argtypes)
_vp, _u64, _sz, _int = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int
SYMBOLS = [
    ("ge_table_create", _int, [_vp, _sz, ctypes.POINTER(_vp)]),
    ("ge_table_destroy", None, [_vp]),
    ("ge_table_record_size", _sz, [_vp]),
    ("ge_table_n_players", _int, [_vp]),
    ("ge_table_phase_io", _int, [_vp, _int, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]),
    ("ge_table_phase_io_packed", _int, [_vp, _int, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]),
    ("ge_batch_create", _int, [_vp, _int, _u64, _u64, _u64, ctypes.POINTER(_vp)]),
    ("ge_batch_reset", _int, [_vp, _u64, _u64]),
    ("ge_batch_clear_stats", _int, [_vp]),
    ("ge_batch_destroy", None, [_vp]),
    ("ge_batch_set_stream", _int, [_vp, _vp]),
    ("ge_batch_set_kernel", _int, [_vp, _int]),
    ("ge_table_wire_size", _sz, [_vp, _int]),
    ("ge_batch_set_wire", _int, [_vp, _int]),
    ("ge_batch_wire_size", _sz, [_vp]),
    ("ge_batch_set_compaction", _int, [_vp, _int, _int]),
    ("ge_batch_set_regroup", _int, [_vp, _int, _int]),
    ("ge_batch_set_grid", _int, [_vp, _int]),
    ("ge_batch_set_autoreset", _int, [_vp, _u64]),
    ("ge_batch_epochs", _int, [_vp, ctypes.POINTER(_u64)]),
    ("ge_batch_active", _int, [_vp, ctypes.POINTER(_u64)]),
    ("ge_batch_active_hint", _int, [_vp, ctypes.POINTER(_u64)]),
    ("ge_batch_get_kernel", _int, [_vp]),
    ("ge_batch_set_option", _int, [_vp, _int, _int]),
    ("ge_batch_set_human_seats", _int, [_vp, _vp]),
    ("ge_batch_set_human_choices", _int, [_vp, _vp]),
    ("ge_table_human_stride", _sz, [_vp]),
    ("ge_step", _int, [_vp, _int, _vp]),
    ("ge_step_many", _int, [ctypes.POINTER(_vp), _int, _int]),
    ("ge_step_ring", _int, [ctypes.POINTER(_vp), _int, _int]),
    ("ge_run_fused", _int, [_vp, _int, _vp]),
    ("ge_sync", _int, [_vp]),
    ("ge_export_state", _int, [_vp, _u64, _u64, _vp]),
    ("ge_import_state", _int, [_vp, _u64, _u64, _vp]),
    ("ge_trace", _int, [_vp, _u64, _u64, _int, _vp]),
    ("ge_run_host", _int, [_vp, _vp, _vp, _int, _vp]),
    ("ge_run_host_async", _int, [_vp, _vp, _vp, _int, _vp]),
    ("ge_batch_set_host_fused", _int, [_vp, _int]),
    ("ge_host_alloc", _int, [ctypes.POINTER(_vp), _sz]),
    ("ge_host_free", None, [_vp]),
    ("ge_eval_preds", _int, [_vp, _vp, _int, _u64, _u64, _vp]),
    ("ge_stats_refresh", _int, [_vp, _vp]),
    ("ge_stats", _int, [_vp, _vp, _sz]),
    ("ge_stats_device_ptr", _vp, [_vp]),
    ("ge_counted_steps", _int, [_vp, ctypes.POINTER(_u64)]),
    ("ge_counted_steps_async", _int, [_vp, _vp]),
    ("ge_state_device_ptr", _vp, [_vp]),
    ("ge_state_device_bytes", _sz, [_vp]),
    ("ge_launch_count", _u64, [_vp]),
    ("ge_last_error", ctypes.c_char_p, []),
    ("ge_version", ctypes.c_char_p, []),
]


class GameEngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("game_engine_b200 error %d: %s" % (code, msg))
        self.code = code


def load(path: str = LIB_PATH) -> ctypes.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            "CUDA library %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = load(os.environ.get("GE_LIB") or LIB_PATH)       # GE_LIB: an A/B build of the same library (build.py GE_LIB_OUT)
    return _lib


def check(rc: int) -> None:
    if rc != GE_OK:
        raise GameEngineError(rc, lib().ge_last_error().decode("utf-8", "replace"))
