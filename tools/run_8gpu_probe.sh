#!/bin/bash
# One call on an N-GPU box: topology, host-copy ceiling at 1/2/4/8 ranks (bound / unbound), bench at 8 / 4 / 2.
N=${1:-8}
mkdir -p gpurun_out
bash tools/probe_box.sh > /dev/null 2>&1; cp gpurun_out/probe_box.txt gpurun_out/probe_box_${N}gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/hostbw_${N}gpu.jsonl
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  $TR --nproc-per-node $n --master-port $((29500+n)) tools/probe_hostbw.py 2>/dev/null | grep '^{' >> gpurun_out/hostbw_${N}gpu.jsonl
done
$TR --nproc-per-node $N --master-port 29600 tools/probe_hostbw.py --nobind 2>/dev/null | grep '^{' >> gpurun_out/hostbw_${N}gpu.jsonl
cat gpurun_out/hostbw_${N}gpu.jsonl
for n in $N 4 2; do
  [ $n -le $N ] || continue
  $TR --nproc-per-node $n --master-port $((29700+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_bench_${n}gpu.log 2>&1
  grep '^{' gpurun_out/r02_bench_${n}gpu.log | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N', l['n_gpus'], 'value', round(l['value']), 'e2e', round(l['e2e']['value']), 'e2e_metrics_only', round(l['e2e_metrics_only']['value']), 'cohort', round(l['cohort_c5']['value']), round(l['cohort_c5']['e2e']['value']))"
done
