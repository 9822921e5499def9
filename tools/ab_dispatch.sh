#!/bin/bash
# A/B of compile-time variants of the library on one box: bash tools/ab_dispatch.sh <variant .so> [label]
# (build the variant with GE_EXTRA_NVCC=-D... GE_LIB_OUT=<path> python -m game_engine_b200.build; capi.py loads $GE_LIB)
V=${1:-$PWD/game_engine_b200/libge_ab_chain.so}; VL=${2:-variant}
run() { # label, lib, args...
  label=$1; shift; lib=$1; shift
  GE_LIB=$lib python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_tmp.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/ab_tmp.json'));print('$label', '%.4e' % d['value'])"
}
L=$PWD/game_engine_b200/libgame_engine_b200.so
for i in 1 2; do
run "cfg2 default" $L --steps 1000
run "cfg2 $VL" $V --steps 1000
done
run "cfg2-canonical-store default" $L --steps 1000 --store canonical
run "cfg2-canonical-store $VL" $V --steps 1000 --store canonical
run "cfg3 default" $L --config 3 --steps 300
run "cfg3 $VL" $V --config 3 --steps 300
run "cfg4 default" $L --config 4 --steps 200
run "cfg4 $VL" $V --config 4 --steps 200
run "p32 default" $L --players 32 --steps 500
run "p32 $VL" $V --players 32 --steps 500
