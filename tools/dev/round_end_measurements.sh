#!/bin/bash
# Round-end evidence, one B200: bench lines for every mix, the ncu launch list of the bench command, one ncu --set full
# capture of the lane kernel per mix at the bench's own size, config 5.  Everything lands in gpurun_out/final/.
cd "$(dirname "$0")/../.."
set -x
mkdir -p gpurun_out/final
python bench.py --steps 20 --warmup 3 > gpurun_out/final/r2_bench_n1.json 2>gpurun_out/final/bench_n1.err
for m in wlo whi uniform adv wsplice; do python bench.py --steps 5 --warmup 3 --no-cpu --mix $m > gpurun_out/final/r2_bench_n1_$m.json 2>/dev/null; done
python bench.py --steps 5 --warmup 3 --no-cpu --ruleset l7_filter > gpurun_out/final/r2_bench_n1_l7filter.json 2>/dev/null
python tools/dev/config5_bench.py 262144 > gpurun_out/final/config5.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/r2_launches_bench_steps2_warmup3.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final/ncu1.log 2>&1
for m in wmix whi wlo; do
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --mix $m > gpurun_out/final/plain_$m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_lane -s 3 -c 1 -f -o gpurun_out/final/prof_r2_$m python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --mix $m > gpurun_out/final/ncu_$m.log 2>&1
done
tail -2 gpurun_out/final/ncu_wmix.log
cut -c1-400 gpurun_out/final/r2_bench_n1.json
