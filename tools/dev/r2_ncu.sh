#!/bin/bash
# ncu --set full capture of the lane kernel on one mix (after the same command ran clean without ncu)
cd "$(dirname "$0")/../.."
mix=${1:-wmix}; streams=${2:-262144}; tag=${3:-r2}
mkdir -p gpurun_out/$tag
python bench.py --streams $streams --steps 1 --warmup 3 --no-cpu --no-e2e --mix $mix > gpurun_out/$tag/plain_$mix.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:scan_lane -s 3 -c 1 -f -o gpurun_out/$tag/prof_$mix python bench.py --streams $streams --steps 1 --warmup 3 --no-cpu --no-e2e --mix $mix > gpurun_out/$tag/ncu_$mix.log 2>&1
tail -2 gpurun_out/$tag/ncu_$mix.log
