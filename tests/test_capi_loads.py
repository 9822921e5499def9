"""The C-ABI library loads on a CPU-only box and exports every symbol include/game_engine_b200.h declares.
No compute calls here (there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "game_engine_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ge_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from game_engine_b200 import capi
    lib = capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    bound = {s[0] for s in capi.SYMBOLS}
    assert bound == set(declared), (bound ^ set(declared))


def test_table_create_validates_without_a_gpu(games):
    from game_engine_b200 import capi
    L = capi.lib()
    cg = games("werewolf-(mafia)", 8)
    h = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(cg.blob, len(cg.blob))
    assert L.ge_table_create(ctypes.cast(buf, ctypes.c_void_p), len(cg.blob), ctypes.byref(h)) == 0
    assert L.ge_table_record_size(h) == 56 and L.ge_table_n_players(h) == 8
    L.ge_table_destroy(h)
    bad = bytearray(cg.blob)
    bad[0] = ord("X")
    buf = ctypes.create_string_buffer(bytes(bad), len(bad))
    assert L.ge_table_create(ctypes.cast(buf, ctypes.c_void_p), len(bad), ctypes.byref(h)) == capi.GE_ERR_ARG
    assert b"magic" in L.ge_last_error()
    # truncated blob and out-of-range branch target
    assert L.ge_table_create(ctypes.cast(buf, ctypes.c_void_p), 40, ctypes.byref(h)) == capi.GE_ERR_ARG
    bad = bytearray(cg.blob)
    bad[32 + 16 + 1] = 200            # phase 0, branch 0, next
    buf = ctypes.create_string_buffer(bytes(bad), len(bad))
    assert L.ge_table_create(ctypes.cast(buf, ctypes.c_void_p), len(bad), ctypes.byref(h)) == capi.GE_ERR_ARG


def test_compute_fails_loudly_without_a_gpu(games):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from game_engine_b200.batch import SessionBatch, Table
    from game_engine_b200.capi import GameEngineError
    with pytest.raises(GameEngineError):
        SessionBatch(Table(games("two-truths-and-a-lie", 4)), 16)


def test_product_never_imports_the_oracle():
    """The package must not import, link or call anything under oracle/ (a product path through the oracle
    would void every parity claim)."""
    pkg = os.path.join(ROOT, "game_engine_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|libge_oracle|ge_cpu_|ref_harness", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn), encoding="utf-8") as f:
                    assert not bad.search(f.read()), fn
