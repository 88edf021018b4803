#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02i_tests.log
tail -4 gpurun_out/r02i_tests.log
bash tools/r02f_run.sh
PA2S_HOSTLAUNCH=1 python tools/trace_step.py > gpurun_out/r02i_timeline.txt 2>&1
