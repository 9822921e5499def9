N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$RUN bench.py --gpus $N --steps 1000 --no-cpu-baseline 2> gpurun_out/r02x_cfg2_n$N.err | grep '^{' > gpurun_out/r02x_cfg2_n$N.json; echo cfg2 rc=$?
for c in 3 5; do $RUN bench.py --gpus $N --config $c --steps 300 --no-e2e --no-cpu-baseline 2> gpurun_out/r02x_cfg${c}_n$N.err | grep '^{' > gpurun_out/r02x_cfg${c}_n$N.json; echo cfg$c rc=$?; done
$RUN tools/ttl_sweep.py --no-cpu 2> gpurun_out/r02x_ttl_sweep_g$N.err | grep '^{' > gpurun_out/r02x_ttl_sweep_g$N.jsonl; echo sweep rc=$?
python - <<PY
import json
for c in (2,3,5):
    try:
        d=json.loads(open('gpurun_out/r02x_cfg%d_n$N.json'%c).read().strip().splitlines()[-1]); e=d.get('e2e') or {}
        print(c, d['config']['workload'][:70], '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%e.get('value',0))
    except Exception as ex: print(c, 'FAILED', ex)
PY
