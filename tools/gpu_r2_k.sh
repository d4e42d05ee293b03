#!/bin/bash
# round 2, GPU call K: long fill A/B (textbook vs U-free chain; pinned parameters + 16-byte checkpoint stores),
# and the batched kernels with running row-0 values (parity + timing)
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
GA_LONG_BLOCKS_PER_SM=1 timeout 800 python tools/tune_long.py 100000 libbase_u0.so libbase_u1.so libad_u0.so libad_u1.so libbase_u0.so libbase_u1.so libad_u0.so libad_u1.so 2>&1 | grep -v '"blocks_per_sm": 2' > $O/tune.log
cut -c1-175 $O/tune.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do python tools/prof_batch.py cfg2 1000000 | tail -1; done
python tools/prof_batch.py cfg5 1000000 | tail -1
