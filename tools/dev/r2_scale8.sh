#!/bin/bash
# 8-GPU box: pinned H2D probe at N = 1/2/4/8 (with and without CPU binding), then bench.py at N = 8 and 4 with its parity block
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2/topo.txt 2>&1
lscpu | grep -E "NUMA|Model name|Socket|^CPU\(s\)" > gpurun_out/r2/lscpu.txt 2>&1
: > gpurun_out/r2/h2d_probe.jsonl
python tools/dev/h2d_probe_mp.py --packed 2>/dev/null | tail -1 >> gpurun_out/r2/h2d_probe.jsonl
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2951$n tools/dev/h2d_probe_mp.py --packed 2>/dev/null | grep '^{' >> gpurun_out/r2/h2d_probe.jsonl
  $TR --nproc-per-node $n --master-port 2952$n tools/dev/h2d_probe_mp.py --packed --bind 2>/dev/null | grep '^{' >> gpurun_out/r2/h2d_probe.jsonl
done
cat gpurun_out/r2/h2d_probe.jsonl
for n in 8 4; do
  $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 2>gpurun_out/r2/bench_n$n.err | grep '^{' > gpurun_out/r2/bench_n$n.json
  python -c "import json; d=json.loads(open('gpurun_out/r2/bench_n$n.json').read().strip().splitlines()[-1]); print($n, d['value'], d['e2e']['value'], d['parity'])"
done
