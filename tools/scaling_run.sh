#!/bin/bash
# bench.py at N = 1, 2, 4, 8 GPUs of one box, one JSON line each -> gpurun_out/final/scaling.jsonl
out=gpurun_out/final
mkdir -p $out
: > $out/scaling.jsonl
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline >> $out/scaling.jsonl 2> $out/scaling_1.err
port=29600
for n in 2 4 8; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 20 --warmup 3 >> $out/scaling.jsonl 2> $out/scaling_$n.err || echo "N=$n failed"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 \
    bench.py --gpus 8 --steps 5 --warmup 3 --workload cfg5 > $out/bench_cfg5_8gpu.json 2> $out/bench_cfg5_8gpu.err || echo "cfg5 N=8 failed"
python -c "
import json
for l in open('$out/scaling.jsonl'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))
d=json.loads(open('$out/bench_cfg5_8gpu.json').read().strip().splitlines()[-1]); print('cfg5', d['n_gpus'], round(d['value'],1), d['scaling'], round(d['ms_per_step'],3), round(d['e2e']['value'],1))
"
