#!/bin/bash
# round-2 dev sweep: GPU parity tests, then the pass time of the main build and of every variant on several mixes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2/gpu.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2/pytest_gpu.log
tail -3 gpurun_out/r2/pytest_gpu.log
fi
one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['matches_per_step_rank0'], d['rescanned_streams'])"; }
for mix in wmix whi wlo uniform adv; do
  echo -n "main $mix "; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one
done | tee gpurun_out/r2/sweep_main.txt
for mix in wmix whi wlo; do
  echo -n "nocalib $mix "; RFB_NO_CALIBRATE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one
  echo -n "ring16 $mix "; RFB_RING_CAP=16 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one
  echo -n "hot0 $mix "; RFB_HOT_ROWS=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one
done | tee gpurun_out/r2/sweep_env.txt
for f in $(ls regex_fpga_b200/lib/variants/*.so 2>/dev/null); do
  for mix in wmix whi; do
  echo -n "$(basename $f) $mix "; RFB_LIB=$PWD/$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one
  done
done | tee gpurun_out/r2/sweep_variants.txt
