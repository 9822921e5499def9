#!/usr/bin/env python3
"""Oracle B under AddressSanitizer + UBSan on random well-formed tables, the shipped tables and ~1500 byte-mutated blobs
(rejected ones exercise the validator, accepted ones are stepped).  Test infrastructure; the checker must itself be
memory-safe before its verdicts are trusted.

    make -C oracle asan
    LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 \
        python oracle/asan_check.py
Last run (build container): clean, 783 tables stepped, 803 rejected.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from game_engine_b200 import table as T, compile_game
from test_fuzz_tables import random_table
lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libge_oracle_asan.so"))
u8p, u64, sz = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t
lib.ge_cpu_table_check.argtypes = [u8p, sz]
lib.ge_cpu_record_size.argtypes = [u8p, sz]; lib.ge_cpu_record_size.restype = sz
lib.ge_cpu_init.argtypes = [u8p, sz, u8p, u64]
lib.ge_cpu_step.argtypes = [u8p, sz, u8p, u64, u64, u64, ctypes.c_int, u8p, ctypes.c_int]
lib.ge_cpu_stats_final.argtypes = [u8p, sz, u8p, u64, u8p]
rng = np.random.default_rng(7)
blobs = [random_table(s, f).pack() for s in range(40) for f in (1, 2)]
blobs += [compile_game(g, p).blob for g, p in (("werewolf-(mafia)", 8), ("werewolf-(mafia)", 32), ("werewolf-revote", 32), ("werewolf-draft", 8), ("two-truths-and-a-lie", 4), ("two-truths-and-a-lie", 32))]
ran = rej = 0
for i in range(len(blobs) + 1500):
    if i < len(blobs):
        blob = blobs[i]
    else:
        b = bytearray(blobs[i % len(blobs)])
        for _ in range(int(rng.integers(1, 5))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if rng.random() < 0.15:
            b = b[: int(rng.integers(1, len(b)))]
        blob = bytes(b)
    buf = ctypes.create_string_buffer(blob, len(blob))
    p = ctypes.cast(buf, ctypes.c_void_p)
    if lib.ge_cpu_table_check(p, len(blob)) != 0:
        rej += 1
        continue
    S = lib.ge_cpu_record_size(p, len(blob))
    n = 64
    rec = np.zeros((n, S), dtype=np.uint8)
    st = np.zeros(560, dtype=np.uint64)
    lib.ge_cpu_init(p, len(blob), rec.ctypes.data, n)
    lib.ge_cpu_step(p, len(blob), rec.ctypes.data, n, 5, i, 60, st.ctypes.data, 2)
    lib.ge_cpu_stats_final(p, len(blob), rec.ctypes.data, n, st.ctypes.data)
    ran += 1
print("asan/ubsan clean: ran", ran, "tables, rejected", rej)
