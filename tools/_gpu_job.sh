python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_fuzz_tables.py tests/test_humans.py -m gpu -x -q > gpurun_out/r02v_tests.log 2>&1; echo tests rc=$?; tail -8 gpurun_out/r02v_tests.log | cut -c1-400
B="python bench.py --steps 300 --no-e2e --no-cpu-baseline"
$B --config 4 --sessions 8388608 2>>gpurun_out/r02v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 n23', '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'])"
$B --config 4 2>>gpurun_out/r02v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 n26', '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'])"
$B --game werewolf-revote --players 8 2>>gpurun_out/r02v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('revote p8', '%.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'])"
python bench.py --steps 1000 --no-e2e --no-cpu-baseline 2>>gpurun_out/r02v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg2', '%.4e'%d['value'])"
bash tools/capture_traffic.sh r02v_cfg4_n23 --game werewolf-revote --players 32 --sessions 8388608 --ring 1 --streams 1 --ctas-per-sm 0
tail -3 gpurun_out/r02v.err
