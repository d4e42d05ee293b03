#!/bin/bash
# round 2, final single-GPU run: the whole -m gpu suite, smoke, the default bench line (all configs) and the reference arm
mkdir -p gpurun_out/r2final
O=gpurun_out/r2final
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; echo "exit $?" >> $O/smoke.txt
timeout 1500 python bench.py > $O/bench_all_1gpu.json 2> $O/bench_all_1gpu.err; echo "exit $?" >> $O/bench_all_1gpu.err
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "exit $?" >> $O/bench_reference_arm.err
for f in pytest_gpu.txt smoke.txt bench_all_1gpu.err bench_reference_arm.err; do echo "== $f"; tail -n 3 $O/$f | cut -c1-300; done
