#!/bin/bash
# quick pass times of the main build (and variants) on the headline mixes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['matches_per_step_rank0'], d['rescanned_streams'])"; }
for mix in ${MIXES:-wmix whi wlo uniform adv}; do
  echo -n "main $mix "; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>gpurun_out/r2/err_$mix.txt | one
done | tee gpurun_out/r2/quick_main.txt
for f in $(ls regex_fpga_b200/lib/variants/*.so 2>/dev/null); do
  for mix in ${VMIXES:-wmix whi}; do
  echo -n "$(basename $f) $mix "; RFB_LIB=$PWD/$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one
  done
done | tee gpurun_out/r2/quick_variants.txt
