#!/bin/bash
# round 2, GPU call E: latency microbenchmark, ncu captures exported to CSV on the box (reports are too big to travel)
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
M="sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum"
tools/micro/latency > $O/latency.txt 2>&1
cap() {   # name, source page too?, command...
  local name=$1 src=$2; shift 2
  "$@" > $O/${name}_plain.txt 2>&1 && \
  timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c $COUNT -o /tmp/$name "$@" > $O/${name}_ncu.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details --csv > $O/${name}_details.csv 2>/dev/null
  if [ "$src" = 1 ]; then ncu -i /tmp/$name.ncu-rep --page source --csv > $O/${name}_src.csv 2>/dev/null; fi
  rm -f /tmp/$name.ncu-rep
}
KREGEX=long_fill SKIP=0 COUNT=1 cap long_fill 1 python tools/prof_long.py
KREGEX="fill_kernel|walk_kernel" SKIP=4 COUNT=2 cap cfg2 0 python tools/prof_batch.py cfg2 250000
KREGEX="fill_kernel|walk_kernel" SKIP=4 COUNT=2 cap cfg3 1 python tools/prof_batch.py cfg3 20000
KREGEX="fill_kernel|walk_kernel" SKIP=4 COUNT=2 cap cfg5 0 python tools/prof_batch.py cfg5 100000
timeout 600 python bench.py --steps 10 --warmup 3 --configs none > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "exit $?" >> $O/bench_cfg2.err
cat $O/latency.txt; tail -n 3 $O/bench_cfg2.err; du -sh $O
