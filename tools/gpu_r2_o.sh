#!/bin/bash
# round 2, GPU call O: launch list of the default bench command (cfg2 headline) -- the kernels' share of the step
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 600 python bench.py --steps 2 --warmup 3 --configs none > $O/bench_plain.json 2> $O/bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_cfg2.csv python bench.py --steps 2 --warmup 3 --configs none > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
echo "exit $?"; tail -c 300 $O/bench_plain.json; ls -la $O
GA_LONG_BLOCKS_PER_SM=1 timeout 500 python tools/tune_long.py 100000 libbase.so liballck.so libs8.so 2>&1 | grep -v "blocks_per_sm\": 2" | cut -c1-175
