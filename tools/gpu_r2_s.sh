#!/bin/bash
# source-level counters of the protein fill kernel (cfg3)
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -c 1 -o /tmp/pf python tools/prof_batch.py cfg3 20000 > $O/ncu.txt 2>&1
ncu -i /tmp/pf.ncu-rep --page source --csv > $O/cfg3_fill_src.csv 2>/dev/null
ls -la $O
