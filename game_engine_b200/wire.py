"""Host-side view of the record formats that cross the C ABI (SPEC.md sections 5 and 5b).

The library converts between them inside its import / export kernels; these NumPy helpers exist for callers that
hold records on the host (building requests, reading results, tests).  Pure byte shuffling — no game rules here."""
from __future__ import annotations

import numpy as np

from . import table as T
from .compiler import CompiledGame


def dense_record_size(cg: CompiledGame) -> int:
    """Bytes per session in the dense wire format; tables it does not cover keep the canonical size."""
    if cg.family == T.FAMILY_WEREWOLF and cg.n_players <= 16:
        return 32 if cg.n_players <= 8 else 48
    return cg.record_size


def has_dense(cg: CompiledGame) -> bool:
    return dense_record_size(cg) != cg.record_size


def to_dense(cg: CompiledGame, canonical: np.ndarray) -> np.ndarray:
    """canonical uint8[n, S] -> dense uint8[n, W] (masks as u8 up to 8 players, u16 up to 16)."""
    rec = np.ascontiguousarray(canonical, dtype=np.uint8).reshape(-1, cg.record_size)
    if not has_dense(cg):
        return rec.copy()
    n = rec.shape[0]
    out = np.zeros((n, dense_record_size(cg)), dtype=np.uint8)
    out[:, 0:8] = rec[:, 0:8]
    masks = rec[:, 8:48].reshape(n, 10, 4)
    if cg.n_players <= 8:
        out[:, 8:18] = masks[:, :, 0]
        out[:, 20:28] = rec[:, 48:56]
    else:
        out[:, 8:28] = masks[:, :, 0:2].reshape(n, 20)
        out[:, 32:48] = rec[:, 48:64]
    return out


def from_dense(cg: CompiledGame, dense: np.ndarray) -> np.ndarray:
    """dense uint8[n, W] -> canonical uint8[n, S]."""
    d = np.ascontiguousarray(dense, dtype=np.uint8).reshape(-1, dense_record_size(cg))
    if not has_dense(cg):
        return d.copy()
    n = d.shape[0]
    out = np.zeros((n, cg.record_size), dtype=np.uint8)
    out[:, 0:8] = d[:, 0:8]
    masks = out[:, 8:48].reshape(n, 10, 4)
    if cg.n_players <= 8:
        masks[:, :, 0] = d[:, 8:18]
        out[:, 48:56] = d[:, 20:28]
    else:
        masks[:, :, 0:2] = d[:, 8:28].reshape(n, 10, 2)
        out[:, 48:64] = d[:, 32:48]
    return out


def dense_padding_is_zero(cg: CompiledGame, dense: np.ndarray) -> np.ndarray:
    """bool[n]: the reserved bytes of each dense record are zero (the library rejects records where they are not)."""
    d = np.ascontiguousarray(dense, dtype=np.uint8).reshape(-1, dense_record_size(cg))
    if not has_dense(cg):
        return np.ones(d.shape[0], dtype=bool)
    pad = np.concatenate([d[:, 18:20], d[:, 28:32]], axis=1) if cg.n_players <= 8 else d[:, 28:32]
    return ~pad.any(axis=1)
