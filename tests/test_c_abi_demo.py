"""The C ABI from plain C: examples/c_abi_demo.c compiles against include/game_engine_b200.h with gcc and links the
shared library (CPU); on the GPU box it runs and prints the same statistics as the Python binding."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "game_engine_b200")


def _build(tmp_path):
    exe = str(tmp_path / "ge_demo")
    cmd = ["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"),
           "-o", exe, "-L" + LIBDIR, "-lgame_engine_b200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_demo_compiles_and_links_as_c(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 2 and "usage:" in out.stderr           # no table given: usage, no CUDA touched


@pytest.mark.gpu
def test_demo_matches_python_binding(tmp_path, games):
    from game_engine_b200.batch import SessionBatch, Table
    cg = games("werewolf-(mafia)", 8)
    blob = tmp_path / "w8.getb"
    blob.write_bytes(cg.blob)
    exe = _build(tmp_path)
    out = subprocess.run([exe, str(blob), "20000", "56", "7"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    m = re.search(r"counted=(\d+) winners=\[(\d+),(\d+),(\d+)\]", out.stdout)
    b = SessionBatch(Table(cg), 20000, first_session_id=0, seed=7)
    b.step(56)
    st = b.stats()
    assert [int(x) for x in m.groups()] == [int(st[0]), int(st[1]), int(st[2]), int(st[3])]
