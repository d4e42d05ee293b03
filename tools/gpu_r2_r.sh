#!/bin/bash
# source-level counters of the batched walk kernel (cfg2)
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -c 1 -o /tmp/wk python tools/prof_batch.py cfg2 250000 > $O/ncu.txt 2>&1
ncu -i /tmp/wk.ncu-rep --page raw --csv > $O/walk_raw.csv 2>/dev/null
ncu -i /tmp/wk.ncu-rep --page source --csv > $O/walk_src.csv 2>/dev/null
ls -la $O
