#!/bin/bash
# round 2, GPU call A: new long-pair fill -- parity first, then the tuning sweep
mkdir -p gpurun_out/r2a
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/gpu.txt
timeout 900 python -m pytest tests/test_gpu_long.py -x -q -m gpu > gpurun_out/r2a/pytest_long.txt 2>&1
echo "pytest_long exit $?" >> gpurun_out/r2a/pytest_long.txt
timeout 200 python tests/fuzz/fuzz_long.py 90 3 > gpurun_out/r2a/fuzz_long.txt 2>&1
timeout 600 python tools/tune_long.py 100000 > gpurun_out/r2a/tune_long.jsonl 2> gpurun_out/r2a/tune_long.err
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_long.py > gpurun_out/r2a/pytest_rest.txt 2>&1
echo "pytest_rest exit $?" >> gpurun_out/r2a/pytest_rest.txt
tail -3 gpurun_out/r2a/pytest_long.txt gpurun_out/r2a/fuzz_long.txt gpurun_out/r2a/pytest_rest.txt
cat gpurun_out/r2a/tune_long.jsonl
