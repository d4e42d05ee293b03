#!/bin/bash
# round 2, GPU call D: long-pair fill v4 (branch-free blocks) + ncu captures of the batched kernels
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
M="sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum"
timeout 300 python -m pytest tests/test_gpu_long.py -x -q -m gpu > $O/pytest_long.txt 2>&1; echo "exit $?" >> $O/pytest_long.txt
timeout 900 python tools/diag_long.py > $O/diag_long.jsonl 2> $O/diag_long.err
timeout 120 python tests/fuzz/fuzz_long.py 60 9 > $O/fuzz_long.txt 2>&1
python tools/prof_long.py > $O/prof_long_plain.txt 2>&1 && \
timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:long_fill -c 1 -o $O/long_fill python tools/prof_long.py > $O/ncu_long.txt 2>&1
python tools/prof_batch.py cfg2 250000 > $O/prof_cfg2_plain.txt 2>&1 && \
timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"fill_kernel|walk_kernel" -s 4 -c 2 -o $O/cfg2 python tools/prof_batch.py cfg2 250000 > $O/ncu_cfg2.txt 2>&1
python tools/prof_batch.py cfg3 20000 > $O/prof_cfg3_plain.txt 2>&1 && \
timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"fill_kernel|walk_kernel" -s 4 -c 2 -o $O/cfg3 python tools/prof_batch.py cfg3 20000 > $O/ncu_cfg3.txt 2>&1
python tools/prof_batch.py cfg5 100000 > $O/prof_cfg5_plain.txt 2>&1 && \
timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"fill_kernel|walk_kernel" -s 4 -c 2 -o $O/cfg5 python tools/prof_batch.py cfg5 100000 > $O/ncu_cfg5.txt 2>&1
for f in pytest_long.txt fuzz_long.txt prof_cfg2_plain.txt prof_cfg3_plain.txt prof_cfg5_plain.txt; do echo "== $f"; tail -n 4 $O/$f; done
cat $O/diag_long.jsonl
