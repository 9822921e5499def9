#!/bin/bash
# A/B of programmatic dependent launches (the default; bench.py --no-pdl switches them off) in the launch modes
run() { label=$1; shift
  python bench.py --no-cpu-baseline --no-e2e --steps 1000 "$@" > gpurun_out/ab_tmp.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/ab_tmp.json'));print('$label', '%.4e' % d['value'])"; }
for i in 1 2; do
run "8 streams x 3 CTAs/SM        " --no-pdl
run "8 streams x 3 CTAs/SM  + PDL "
done
run "1 stream, full grids         " --streams 1 --ctas-per-sm 0 --no-pdl
run "1 stream, full grids   + PDL " --streams 1 --ctas-per-sm 0
run "2 streams, full grids        " --streams 2 --ctas-per-sm 0 --no-pdl
run "2 streams, full grids  + PDL " --streams 2 --ctas-per-sm 0
run "4 streams x 4 CTAs/SM        " --streams 4 --ctas-per-sm 4 --no-pdl
run "4 streams x 4 CTAs/SM  + PDL " --streams 4 --ctas-per-sm 4
run "ring launch                  " --launch ring --no-pdl
run "ring launch            + PDL " --launch ring
run "cfg5 TTL                     " --config 5 --steps 100 --no-pdl
run "cfg5 TTL               + PDL " --config 5 --steps 100
run "cfg3                         " --config 3 --steps 300 --no-pdl
run "cfg3                   + PDL " --config 3 --steps 300
