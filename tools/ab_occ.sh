#!/bin/bash
# A/B of the occupancy the 8- and 16-player step kernels are compiled for (__launch_bounds__ via -DGE_P8_CTAS / -DGE_P16_CTAS)
run() { label=$1; shift; lib=$1; shift
  GE_LIB=$lib python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_tmp.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/ab_tmp.json'));print('$label', '%.4e' % d['value'])"; }
D=$PWD/game_engine_b200/libgame_engine_b200.so; A=$PWD/game_engine_b200/libge_ab_occ9.so; B=$PWD/game_engine_b200/libge_ab_occ10.so
for i in 1 2; do
run "cfg2 8 CTAs/SM (default) x3" $D --steps 1000
run "cfg2 9 CTAs/SM           x3" $A --steps 1000
run "cfg2 10 CTAs/SM          x3" $B --steps 1000
done
run "cfg2 10 CTAs/SM, grids of 4" $B --steps 1000 --ctas-per-sm 4
run "cfg2 10 CTAs/SM, grids of 5" $B --steps 1000 --ctas-per-sm 5
run "cfg2 10 CTAs/SM, grids of 2" $B --steps 1000 --ctas-per-sm 2
run "cfg2 9 CTAs/SM, grids of 4 " $A --steps 1000 --ctas-per-sm 4
run "cfg2 1 stream full: 8 " $D --steps 1000 --streams 1 --ctas-per-sm 0
run "cfg2 1 stream full: 10" $B --steps 1000 --streams 1 --ctas-per-sm 0
run "cfg2 ring: 8 " $D --steps 1000 --launch ring
run "cfg2 ring: 10" $B --steps 1000 --launch ring
run "cfg3 6 CTAs/SM (default)" $D --config 3 --steps 300
run "cfg3 7 CTAs/SM          " $A --config 3 --steps 300
run "cfg3 8 CTAs/SM          " $B --config 3 --steps 300
