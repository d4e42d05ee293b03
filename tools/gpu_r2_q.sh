#!/bin/bash
# source-level counters of the X-propagation kernel of the long-pair traceback
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 600 ncu --set full --clock-control none --import-source on -k regex:long_xprop -c 1 -o /tmp/xp python tools/prof_long.py 100000 > $O/ncu.txt 2>&1
ncu -i /tmp/xp.ncu-rep --page raw --csv > $O/xprop_raw.csv 2>/dev/null
ncu -i /tmp/xp.ncu-rep --page source --csv > $O/xprop_src.csv 2>/dev/null
ls -la $O
