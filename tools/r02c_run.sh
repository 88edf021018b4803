#!/bin/bash
# round 2: multi-sequence decoder -- GPU tests, then the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02c_tests.log
tail -40 gpurun_out/r02c_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
tail -5 gpurun_out/r02c_bench_n1.err; head -c 300 gpurun_out/r02c_bench_n1.json
