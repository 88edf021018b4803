#!/bin/bash
mkdir -p gpurun_out
python tools/prof_decm.py > gpurun_out/r02d_prof_decm.txt 2>&1
cat gpurun_out/r02d_prof_decm.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02d_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02d_tests.log
tail -15 gpurun_out/r02d_tests.log
