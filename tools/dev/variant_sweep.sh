#!/bin/bash
# lane-kernel pass time for every library variant under regex_fpga_b200/lib/variants (or the names given); dev tool
cd "$(dirname "$0")/../.."
for f in ${@:-$(ls regex_fpga_b200/lib/variants/*.so)}; do
  echo -n "$(basename $f) "; RFB_LIB=$PWD/$f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e ${MIX:+--mix $MIX} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['matches_per_step_rank0'])"
done
