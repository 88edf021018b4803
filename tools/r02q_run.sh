#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02q_tests.log
tail -6 gpurun_out/r02q_tests.log
bash tools/r02f_run.sh
bash tools/prof_round.sh r02p
