N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$RUN tools/pcie_probe.py > gpurun_out/r02r_pcie_n$N.json 2> gpurun_out/r02r_pcie_n$N.err; echo pcie rc=$?; cut -c1-300 gpurun_out/r02r_pcie_n$N.json
$RUN bench.py --gpus $N --steps 1000 --no-cpu-baseline 2> gpurun_out/r02r_cfg2_n$N.err | grep '^{' > gpurun_out/r02r_cfg2_n$N.json; echo cfg2 rc=$?
for c in 3 4 5; do $RUN bench.py --gpus $N --config $c --steps 300 --no-e2e --no-cpu-baseline 2> gpurun_out/r02r_cfg${c}_n$N.err | grep '^{' > gpurun_out/r02r_cfg${c}_n$N.json; echo cfg$c rc=$?; done
$RUN tools/ttl_sweep.py --no-cpu > gpurun_out/r02r_ttl_sweep_g$N.jsonl 2> gpurun_out/r02r_ttl_sweep_g$N.err; echo sweep rc=$?
python - <<PY
import json
for c in (2,3,4,5):
    try:
        d=json.loads(open('gpurun_out/r02r_cfg%d_n$N.json'%c).read().strip().splitlines()[-1]); e=d.get('e2e') or {}
        print(c, d['config']['workload'][:70], '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'phys', (d['roofline'].get('dram_physical') or {}).get('gbs'), 'e2e %.3e'%e.get('value',0), 'allreduce %.1f ms'%d.get('stats_allreduce_ms',0))
    except Exception as ex: print(c, 'FAILED', ex)
PY
tail -3 gpurun_out/r02r_cfg4_n$N.err
