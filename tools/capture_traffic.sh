#!/bin/bash
# Physical DRAM traffic of one workload (run on the GPU box): plain run, then the same command under ncu with the caches
# left alone.  Usage: bash tools/capture_traffic.sh <tag> <range_profile.py args...>; outputs gpurun_out/<tag>_{plain.json,phys.csv,phys.log}
TAG=$1; shift
python tools/range_profile.py "$@" > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
timeout 900 ncu --cache-control none --clock-control none --profile-from-start off \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum \
    --csv --log-file gpurun_out/${TAG}_phys.csv python tools/range_profile.py "$@" > gpurun_out/${TAG}_phys.log 2>&1
echo "$TAG rc=$? $(cat gpurun_out/${TAG}_plain.json | cut -c1-200)"
