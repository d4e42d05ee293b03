#!/bin/bash
# round 2, 8-GPU call: bench.py under torchrun at N = 8 (cfg2 weak + cfg5 strong + devices_api)
mkdir -p gpurun_out/r2n8
O=gpurun_out/r2n8
nvidia-smi -L > $O/gpus.txt
nproc >> $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "exit $?" >> $O/bench_n8.err
tail -n 3 $O/bench_n8.err
