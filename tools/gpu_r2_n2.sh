#!/bin/bash
# round 2, 2-GPU call: the one-process devices= API on two GPUs, bench.py under torchrun (cfg2 weak + cfg5 strong + devices_api)
mkdir -p gpurun_out/r2n2
O=gpurun_out/r2n2
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_gpu or rle" > $O/pytest_two_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_two_gpu.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "exit $?" >> $O/bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --configs none > $O/bench_n1.json 2> $O/bench_n1.err; echo "exit $?" >> $O/bench_n1.err
for f in pytest_two_gpu.txt bench_n2.err bench_n1.err; do echo "== $f"; tail -n 5 $O/$f; done
