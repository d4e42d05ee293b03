#!/bin/bash
# round 2, GPU call T: ncu capture of the final protein fill + walk (cfg3 sample), same form as gpu_r2_e.sh
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
M="sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum"
python tools/prof_batch.py cfg3 20000 > $O/cfg3_plain.txt 2>&1 && \
timeout 600 ncu --set full --metrics $M --clock-control none -k regex:"fill_kernel|walk_kernel" -s 4 -c 2 -o /tmp/cfg3 python tools/prof_batch.py cfg3 20000 > $O/cfg3_ncu.txt 2>&1
ncu -i /tmp/cfg3.ncu-rep --page raw --csv > $O/cfg3_raw.csv 2>/dev/null
ncu -i /tmp/cfg3.ncu-rep --page details --csv > $O/cfg3_details.csv 2>/dev/null
tail -1 $O/cfg3_plain.txt; ls -la $O
