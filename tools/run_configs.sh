#!/bin/bash
# BASELINE.json configs 2-5 on N GPUs of one box (one rank per GPU, torchrun); JSON lines go to gpurun_out/.
#   bash tools/run_configs.sh 8 [tag]
N=${1:-8}; TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" -gt 1 ]; then
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
else
  RUN="python"
fi
run() {  # name, args...
  name=$1; shift
  $RUN bench.py --gpus $N "$@" 2>$OUT/${TAG}_${name}_n${N}.err | grep '^{' > $OUT/${TAG}_${name}_n${N}.json
  python - "$OUT/${TAG}_${name}_n${N}.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-10s N=%d  %.4e steps/s  %.3f of HBM/GPU  e2e %s  allreduce %.1f ms  %s" % (
        sys.argv[2], d["n_gpus"], d["value"], d["roofline"]["frac"],
        ("%.3e" % d["e2e"]["value"]) if d.get("e2e") else "-", d.get("stats_allreduce_ms", 0), d["config"]["workload"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
S3=$(( (1 << 24) / N )); S4=$(( (1 << 26) / N )); S5=$(( (1 << 28) / N ))
run cfg2 
run cfg3 --players 16 --sessions $S3 --ring 4 --streams 4 --steps 500 --no-e2e
run cfg4 --game werewolf-revote --players 32 --sessions $S4 --ring 1 --streams 1 --ctas-per-sm 0 --steps 400 --warmup 25 --no-e2e
run cfg5 --game two-truths-and-a-lie --players 4 --sessions $S5 --ring 2 --streams 2 --ctas-per-sm 0 --steps 200 --warmup 20 --no-e2e
