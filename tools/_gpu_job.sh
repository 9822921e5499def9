B="python bench.py --config 4 --sessions 8388608 --steps 400 --no-e2e --no-cpu-baseline"
for rg in 5,3 6,3 8,3 10,3 5,2 8,2 5,1 8,1 12,2; do $B --regroup $rg 2>>gpurun_out/r02w.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('regroup $rg', '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'], d['gpu_launches'])"; done
tail -3 gpurun_out/r02w.err
