#!/bin/bash
# round 2, GPU call G: final long-pair fill (6 rows x 1 chain per lane, mid-block prefetch): parity, timing, ncu
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
M="sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum"
timeout 300 python -m pytest tests/test_gpu_long.py -x -q -m gpu > $O/pytest_long.txt 2>&1; echo "exit $?" >> $O/pytest_long.txt
timeout 300 python tools/diag_long.py "" > $O/diag_long.jsonl 2> $O/diag_long.err
timeout 200 python tests/fuzz/fuzz_long.py 120 13 > $O/fuzz_long.txt 2>&1
timeout 600 python bench.py --workload cfg4 --configs cfg1 --steps 10 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "exit $?" >> $O/bench_cfg4.err
python tools/prof_long.py 100000 > $O/long_fill_plain.txt 2>&1 && \
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:long_fill -c 1 -o /tmp/long_fill python tools/prof_long.py 100000 > $O/long_fill_ncu.txt 2>&1
ncu -i /tmp/long_fill.ncu-rep --page raw --csv > $O/long_fill_raw.csv 2>/dev/null
ncu -i /tmp/long_fill.ncu-rep --page details --csv > $O/long_fill_details.csv 2>/dev/null
ncu -i /tmp/long_fill.ncu-rep --page source --csv > $O/long_fill_src.csv 2>/dev/null
for f in pytest_long.txt fuzz_long.txt bench_cfg4.err long_fill_plain.txt; do echo "== $f"; tail -n 4 $O/$f; done
cat $O/diag_long.jsonl
