#!/bin/bash
# Round-end measurement pass on one B200: GPU test suite, every bench line, then (only after the
# plain runs exited 0) the ncu launch list and one full capture of the fill and walk kernels.
set -u
out=gpurun_out/final
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/pytest_gpu.txt; cat $out/pytest_gpu.txt
python bench.py > $out/bench_cfg2.json 2> $out/bench_cfg2.err || echo "bench cfg2 failed"
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_cfg2_reference.json 2> $out/ref.err || echo "reference arm failed"
for w in cfg1 cfg3 cfg4 cfg5; do
  python bench.py --workload $w > $out/bench_$w.json 2> $out/bench_$w.err || echo "bench $w failed"
done
python -c "
import json,glob
for f in sorted(glob.glob('$out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d.get('value',0),1), d.get('unit'), 'ms/step', round(d.get('ms_per_step',0),3), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'kernels', d.get('kernels_ms'))
    except Exception as e: print(f, 'unreadable', e)
"
if [ "${1:-}" = "ncu" ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/pre_ncu.json 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_1m.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'fill_kernel|walk_kernel' -c 2 -f -o $out/fill_walk_final \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1
  ls -la $out
fi
