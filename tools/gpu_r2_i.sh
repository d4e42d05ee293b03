#!/bin/bash
# round 2, GPU call I: long fill with TMA-staged sequence tiles + warp-uniform spin loops (parity, fuzz),
# and source-level counters of the batched cfg2 fill (do its shuffles ever take the divergent path?)
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 300 python -m pytest tests/test_gpu_long.py -x -q -m gpu 2>&1 | tail -2
timeout 150 python tests/fuzz/fuzz_long.py 100 31 2>&1 | tail -1
python tools/prof_long.py 100000 | tail -1
python tools/prof_batch.py cfg2 200000 > $O/plain.txt 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -c 1 -o /tmp/bf python tools/prof_batch.py cfg2 200000 > $O/ncu.txt 2>&1
ncu -i /tmp/bf.ncu-rep --page source --csv > $O/cfg2_fill_src.csv 2>/dev/null
tail -1 $O/plain.txt; ls -la $O
