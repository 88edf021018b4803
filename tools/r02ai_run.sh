#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "three_piece" 2>&1 | grep -E "conv |passed|failed"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02ai_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02ai_tests.log; tail -3 gpurun_out/r02ai_tests.log
