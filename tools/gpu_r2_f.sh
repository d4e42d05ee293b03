#!/bin/bash
# round 2, GPU call F: long fill v5 (identity-gated edge blocks), RLE pipeline fix, full parity
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 300 python -m pytest tests/test_gpu_long.py -x -q -m gpu > $O/pytest_long.txt 2>&1; echo "exit $?" >> $O/pytest_long.txt
timeout 900 python tools/diag_long.py > $O/diag_long.jsonl 2> $O/diag_long.err
timeout 120 python tests/fuzz/fuzz_long.py 60 11 > $O/fuzz_long.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_long.py > $O/pytest_rest.txt 2>&1; echo "exit $?" >> $O/pytest_rest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --configs none > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "exit $?" >> $O/bench_cfg2.err
for f in pytest_long.txt fuzz_long.txt pytest_rest.txt bench_cfg2.err; do echo "== $f"; tail -n 4 $O/$f; done
cat $O/diag_long.jsonl
