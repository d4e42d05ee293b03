#!/bin/bash
# round 2, GPU call J: source-level counters of the long-pair fill (where does a lone warp stall?)
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
GA_LONG_BLOCKS_PER_SM=1 timeout 300 python tools/tune_long.py 100000 libshfl.so libxsmem.so 2>&1 | grep -v "blocks_per_sm\": 2" | cut -c1-175
python tools/prof_long.py 100000 > $O/plain.txt 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:long_fill -c 1 -o /tmp/lf python tools/prof_long.py 100000 > $O/ncu.txt 2>&1
ncu -i /tmp/lf.ncu-rep --page raw --csv > $O/long_fill_raw.csv 2>/dev/null
ncu -i /tmp/lf.ncu-rep --page details --csv > $O/long_fill_details.csv 2>/dev/null
ncu -i /tmp/lf.ncu-rep --page source --csv > $O/long_fill_src.csv 2>/dev/null
tail -1 $O/plain.txt; ls -la $O
