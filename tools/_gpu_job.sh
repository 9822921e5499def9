python -m pytest tests -m gpu -q > gpurun_out/r02o_tests.log 2>&1; echo tests rc=$?; tail -15 gpurun_out/r02o_tests.log | cut -c1-600
python - <<PY
import numpy as np
from game_engine_b200 import compile_game
from game_engine_b200.batch import SessionBatch, Table
from oracle.oracle import Oracle
for P in (8, 16, 32):
    cg = compile_game("werewolf-(mafia)", P); o = Oracle(cg.blob)
    n = 100003
    for kernel in ("tps", "tps_generic"):
        b = SessionBatch(Table(cg), n, first_session_id=5, seed=3, kernel=kernel)
        b.set_option("light_bulk", 1)
        b.step(70)
        rec = o.init(n); o.step(rec, 5, 3, 70)
        print("bulk parity", P, kernel, bool(np.array_equal(b.export_state(), rec)))
PY
for i in 1 2; do python bench.py --steps 1000 --no-cpu-baseline --no-e2e > gpurun_out/r02o_bench_ldg$i.json 2>>gpurun_out/r02o_bench.err; python bench.py --steps 1000 --light-bulk --no-cpu-baseline --no-e2e > gpurun_out/r02o_bench_bulk$i.json 2>>gpurun_out/r02o_bench.err; done; tail -3 gpurun_out/r02o_bench.err
python -c "
import json
for m in ('ldg1','bulk1','ldg2','bulk2'):
    d=json.load(open('gpurun_out/r02o_bench_%s.json'%m)); print(m, '%.4e'%d['value'], d['config']['light_path'])
"
