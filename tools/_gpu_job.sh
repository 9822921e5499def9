python -m pytest tests/test_wire.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1; echo tests rc=$?; tail -8 gpurun_out/r02q_tests.log | cut -c1-400
for i in 1 2; do python bench.py --steps 300 --no-cpu-baseline > gpurun_out/r02q_bench$i.json 2>>gpurun_out/r02q_bench.err; done
python -c "
import json
for m in ('1','2'):
    d=json.load(open('gpurun_out/r02q_bench%s.json'%m)); e=d['e2e']; print(m, '%.4e'%d['value'], 'e2e %.3e'%e['value'], e['ms_per_call'], 'single %.3e'%e['single_step_round_trip']['value'])
"
