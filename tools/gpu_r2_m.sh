#!/bin/bash
# round 2, GPU call M: the shipped long-pair fill (shuffle hand-off, U-free chain, pair table, TMA sequence
# tiles, warp-uniform spin loops): parity, fuzz, timing, ncu capture + launch list
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
M="sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum"
timeout 300 python -m pytest tests/test_gpu_long.py -x -q -m gpu > $O/pytest_long.txt 2>&1; echo "exit $?" >> $O/pytest_long.txt
timeout 200 python tests/fuzz/fuzz_long.py 120 53 > $O/fuzz_long.txt 2>&1
GA_LONG_PAIR_TABLE=0 timeout 100 python tests/fuzz/fuzz_long.py 40 59 > $O/fuzz_long_single_table.txt 2>&1
timeout 600 python bench.py --workload cfg4 --configs cfg1 --steps 10 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "exit $?" >> $O/bench_cfg4.err
python tools/prof_long.py 100000 > $O/long_fill_plain.txt 2>&1 && \
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:long_fill -c 1 -o /tmp/long_fill python tools/prof_long.py 100000 > $O/long_fill_ncu.txt 2>&1
ncu -i /tmp/long_fill.ncu-rep --page raw --csv > $O/long_fill_raw.csv 2>/dev/null
ncu -i /tmp/long_fill.ncu-rep --page details --csv > $O/long_fill_details.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_long_100kb.csv python tools/prof_long.py 100000 > $O/launches.log 2>&1
for f in pytest_long.txt fuzz_long.txt fuzz_long_single_table.txt bench_cfg4.err long_fill_plain.txt; do echo "== $f"; tail -n 3 $O/$f | cut -c1-300; done
