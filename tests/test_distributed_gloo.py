"""N>1 host logic on CPU: world_size-2 gloo.  Each rank owns a contiguous session shard, produces its
statistics words (the CPU oracle stands in for the per-rank producer; there is no GPU here) and the
SUM all-reduce must equal the statistics of the unsharded run, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, seed, steps, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from game_engine_b200 import compile_game
    from game_engine_b200.parallel import allreduce_stats, shard_range
    from oracle.oracle import Oracle
    cg = compile_game("werewolf-(mafia)", 16)
    o = Oracle(cg.blob)
    first, count = shard_range(n_total, world, rank)
    rec = o.init(count)
    st = o.new_stats()
    o.step(rec, first, seed, steps, st, threads=1)
    o.stats_final(rec, st)
    t = torch.from_numpy(st.view(np.int64).copy())
    allreduce_stats(t)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), t.numpy())
    np.save(os.path.join(out_dir, "state%d.npy" % rank), rec)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_exactly():
    from game_engine_b200.parallel import shard_range
    for n in (1, 7, 1000, 1 << 20, (1 << 24) + 3):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_range(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


@pytest.mark.timeout(300)
def test_world_size_2_allreduce_equals_unsharded(tmp_path, games, oracle_for):
    n_total, seed, steps, world = 3001, 17, 140, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, seed, steps, str(tmp_path)), nprocs=world, join=True)
    cg = games("werewolf-(mafia)", 16)
    o = oracle_for(cg)
    rec = o.init(n_total)
    st = o.new_stats()
    o.step(rec, 0, seed, steps, st)
    o.stats_final(rec, st)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(r0, r1)                                  # every rank holds the global result
    np.testing.assert_array_equal(r0.view(np.uint64), st)                  # and it equals the 1-rank histogram
    both = np.concatenate([np.load(tmp_path / "state0.npy"), np.load(tmp_path / "state1.npy")])
    np.testing.assert_array_equal(both, rec)                               # sharding is invisible in the states
