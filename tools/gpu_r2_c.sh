#!/bin/bash
# round 2, GPU call C: long-pair fill v3 (pipelined shuffles), RLE output, bench e2e with RLE
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 300 python -m pytest tests/test_gpu_long.py -x -q -m gpu > $O/pytest_long.txt 2>&1; echo "exit $?" >> $O/pytest_long.txt
timeout 900 python tools/diag_long.py > $O/diag_long.jsonl 2> $O/diag_long.err
timeout 120 python tests/fuzz/fuzz_long.py 60 7 > $O/fuzz_long.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_long.py > $O/pytest_rest.txt 2>&1; echo "exit $?" >> $O/pytest_rest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --configs cfg3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "exit $?" >> $O/bench_cfg2.err
python tools/prof_long.py > $O/prof_long_plain.txt 2>&1 && \
timeout 600 ncu --set full --metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum --clock-control none --import-source on -k regex:long_fill -c 1 -o $O/long_fill python tools/prof_long.py > $O/ncu_long.txt 2>&1
for f in pytest_long.txt fuzz_long.txt pytest_rest.txt bench_cfg2.err; do echo "== $f"; tail -n 6 $O/$f; done
cat $O/diag_long.jsonl
