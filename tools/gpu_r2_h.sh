#!/bin/bash
# round 2, GPU call H: where does a band sweep with TMA-staged letters lose its time? (source-level counters)
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
export GA_B200_LIB=$PWD/globalign_b200/_variants/libx3.so
python tools/prof_long.py 20000 > $O/plain.txt 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:long_fill -c 1 -o /tmp/lf python tools/prof_long.py 20000 > $O/ncu.txt 2>&1
ncu -i /tmp/lf.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i /tmp/lf.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
tail -2 $O/plain.txt; ls -la $O
