# config 4 experiments on one GPU (2^23 sessions, the N=8 shard size): regroup cadence and occupancy of the 32-player kernel
B="python bench.py --config 4 --sessions 8388608 --steps 300 --no-e2e --no-cpu-baseline"
for rg in 5,3 3,3 2,3 1,3 1,2 1,4 2,4; do $B --regroup $rg 2>>gpurun_out/r02s.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('regroup $rg', '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'], d['gpu_launches'])"; done
for v in 5 6; do GE_LIB=$PWD/tools/_variants/libge_occ$v.so $B 2>>gpurun_out/r02s.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('occ $v cfg4', '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'])"; done
for v in 5 6; do GE_LIB=$PWD/tools/_variants/libge_occ$v.so python bench.py --players 32 --steps 300 --no-e2e --no-cpu-baseline 2>>gpurun_out/r02s.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('occ $v werewolf P32', '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'])"; done
python bench.py --players 32 --steps 300 --no-e2e --no-cpu-baseline 2>>gpurun_out/r02s.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('occ 4 werewolf P32', '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'])"
tail -3 gpurun_out/r02s.err
