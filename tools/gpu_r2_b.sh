#!/bin/bash
# round 2, GPU call B: long-pair fill v2 -- diagnostics, parity, first run of the new bench, ncu
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 300 python -m pytest tests/test_gpu_long.py -x -q -m gpu > $O/pytest_long.txt 2>&1; echo "exit $?" >> $O/pytest_long.txt
timeout 900 python tools/diag_long.py > $O/diag_long.jsonl 2> $O/diag_long.err
timeout 120 python tests/fuzz/fuzz_long.py 60 5 > $O/fuzz_long.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_long.py > $O/pytest_rest.txt 2>&1; echo "exit $?" >> $O/pytest_rest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_all.json 2> $O/bench_all.err; echo "exit $?" >> $O/bench_all.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "exit $?" >> $O/bench_ref.err
python tools/prof_long.py > $O/prof_long_plain.txt 2>&1 && \
timeout 600 ncu --set full --metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum --clock-control none --import-source on -k regex:long_fill -c 1 -o $O/long_fill python tools/prof_long.py > $O/ncu_long.txt 2>&1
for f in pytest_long.txt fuzz_long.txt pytest_rest.txt bench_all.err bench_ref.err; do echo "== $f"; tail -n 4 $O/$f; done
cat $O/diag_long.jsonl
